// wdp.cu -- K3: batched wrap-around DP (fill + argmax + direction bits) and traceback for sm_100a.
//
// Replaces wrap_around_DP_sub (/root/reference/wrap_around_DP.c:222-354), the DP nest inside
// revise_representative_unit_sub (consensus.c:875-962) and pretty_print_alignment (wrap_around_DP.c:68-186).
//
// The recurrence (wrap_around_DP.c:258-285):
//     x_i == u_j :  W[i][j] = W[i-1][j-1] + G                                  (unconditionally)
//     else       :  W[i][j] = max(0, W[i-1][j-1]-MM, W[i-1][j]-IN, (j>1) W[i][j-1]-IN)
//     W[i][0] := W[i][U]  (wrap)  => cell (i,1) depends on (i-1,U): rows are strictly sequential.
// There is no valid anti-diagonal wavefront; the parallelism inside one DP is along a
// row.  A group of G lanes owns one job; lane l holds C consecutive columns in registers.  Each row:
//   phase A   per cell, the left-independent candidate  a = max(0, diag-MM, up-IN)  (or diag+G on a match)
//   phase B   the left chain  W[j] = max(a_j, W[j-1]-IN)  inside the lane, speculating that nothing
//             crosses the lane boundary; the last cell of every lane is handed to its right neighbour by
//             one shuffle and the chain is redone only while some lane's first cell really changes
//             (monotone fixpoint, at most G rounds, usually none or one)
//   epilogue  W[i][U] is shuffled to lane 0 (wrap), direction codes are packed and stored, argmax updated.
//
// Scores are kept multiplied by 4 (paired int16x2 kernels) or 16 (int32 kernels: four free low bits in an untagged score
// hold the column tag of the argmax key) with the low two bits free, so that the candidate that wins the max
// carries its own traceback code: diag-MM is tagged 3, left-IN 2, up-IN 1, the zero floor 0.  A tie between
// candidates is then resolved by the max itself in the reference's traceback priority (mismatch before
// deletion before insertion, wrap_around_DP.c:306-323) and the 2-bit direction code is just (r & 3):
// no compares.  Match cells need no code: the traceback tests x_i == u_j first (:306), and a match cell
// always equals diag+G.  The j == 1 "deletion" test of the traceback reads W[i][0], the wrap copy of the
// same row (:302,314); that code is patched in lane 0 at row end, when W[i][U] is known.
//
// Direction matrix in HBM: 2 bits per slot, row-major, row stride = G*C/4 bytes, written with one
// coalesced 1/2/4-byte store per lane per row.  The traceback kernel walks it with one thread per task,
// tracking the running score exactly as the reference does (max_wrd), so no "stop" code is needed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>
#include "mtr_internal.h"

// ---------------------------------------------------------------- job classes
// U <= G*C.  Classes 0-9 maximise throughput (few lanes per job, long register chains: more jobs per warp, less
// carry traffic), with capacities 16, 32, 48, 64, 96, 128, 192, 256, 384, 512 so that at most a third of the slots
// is padding; classes 10-15 minimise the latency of one row (many lanes, C = 4 where possible) and are used when a
// batch is too small to fill the GPU anyway.  Every class writes the same direction-matrix layout (slot s of a row
// at bit 2*(s%4) of byte s/4).
static const WdpClass kClasses[WDP_NCLASS] = {
    {4, 4, 0},  {4, 8, 0},  {4, 12, 0}, {8, 8, 0},  {8, 12, 0},  {8, 16, 0}, {16, 12, 0}, {16, 16, 0}, {32, 12, 0}, {32, 16, 0},
    {4, 4, 0},  {8, 4, 0},  {16, 4, 0}, {32, 4, 0}, {32, 8, 0},  {32, 16, 0},
};
constexpr int kThroughputClasses = 10, kLatencyClasses = 6;
constexpr int kPairedBase = 16;          // classes 16..25: the throughput classes as paired int16x2 kernels

static int class_of(int ulen, int latency)
{
    const int lo = latency ? kThroughputClasses : 0, n = latency ? kLatencyClasses : kThroughputClasses;
    for (int k = lo; k < lo + n; k++)
        if (ulen <= kClasses[k].G * kClasses[k].C) return k;
    return -1;
}

long long wdp_dir_bytes(int ulen, int rows)
{
    const int k = class_of(ulen, 0);
    if (k < 0) return 0;
    return ((long long)rows * (kClasses[k].G * kClasses[k].C / 4) + 15) & ~15LL;
}

#define NEG_INF (-(1 << 28))

__device__ __forceinline__ int read_base(const uint32_t *__restrict__ packed, long long b)
{
    return (int)((packed[b >> 4] >> ((int)(b & 15) * 2)) & 3u);
}

// ---------------------------------------------------------------- traceback of one (task, penalty set), one thread
// wrap_around_DP.c:288-333 (counts), consensus.c:919-962 (histograms), wrap_around_DP.c:123-186 (path).
// Called by the lane that owns the task right after its fill (fused: the direction rows are still in L1/L2 and the
// other warps of the SM keep filling meanwhile), or by the stand-alone wdp_traceback kernel.
__device__ __forceinline__ void traceback_one(const WdpTask &t, const int p, const int best, const int max_i, const int max_j,
                                              const uint32_t *__restrict__ packed, const uint8_t *__restrict__ units,
                                              const uint8_t *dirs, mtr_wdp_result *res, void *aux)
{
    const int G = t.gain[p], MM = t.mis[p], IN = t.indel[p];
    const int ulen = t.ulen, dstride = t.dir_stride;
    const long long unit_off = t.unit_off;
    const long long base0 = t.base0, aux_cap = t.aux_cap;
    const uint8_t *d = dirs + t.dir_off + (size_t)p * t.dir_bytes;
    int i = max_i, j = max_j, run = best;
    if (j == 0) j = ulen;                                   // wrap_around_DP.c:296
    int nm = 0, nx = 0, ni = 0, nd = 0, steps = 0, flags = 0;
    int *cons = nullptr, *miss = nullptr;
    uint8_t *path = nullptr;
    if (t.mode == MTR_TB_CONSENSUS) {
        cons = (int *)aux + t.aux_off;
        miss = cons + (size_t)(ulen + 1) * 5;
    } else if (t.mode == MTR_TB_PATH) {
        path = (uint8_t *)aux + t.aux_off;
    }
    // The walk is a chain of dependent loads (read base, unit base, direction byte).  ~85 % of the steps of a
    // real alignment are diagonal, so the operands of the next TB_DEPTH cells down the diagonal are fetched
    // together (independent loads in flight) and consumed until the path leaves the diagonal.
    constexpr int TB_DEPTH = 8;
    bool alive = i > 0 && run > 0;
    while (alive) {
        int qx[TB_DEPTH], qu[TB_DEPTH], qd[TB_DEPTH];
        {
            int jt = j;
#pragma unroll
            for (int q = 0; q < TB_DEPTH; q++) {
                const int it = i - q;
                const bool ok = it >= 1;
                qx[q] = ok ? read_base(packed, base0 + it) : 0;
                qu[q] = units[unit_off + jt - 1];
                qd[q] = ok ? d[(size_t)(it - 1) * dstride + ((jt - 1) >> 2)] : 0;
                jt = jt == 1 ? ulen : jt - 1;
            }
        }
#pragma unroll
        for (int q = 0; q < TB_DEPTH; q++) {
            if (!(i > 0 && run > 0)) { alive = false; break; }
            const int xi = qx[q], uj = qu[q];
            int op;
            if (xi == uj) {
                op = 0;
            } else {
                const int code = (qd[q] >> (((j - 1) & 3) * 2)) & 3;
                if (code == 0) { flags |= 2; alive = false; break; }       // cannot happen: run > 0 means W[i][j] > 0
                op = code == 3 ? 1 : (code == 2 ? 2 : 3);
            }
            if (path) { if (steps < aux_cap) path[steps] = (uint8_t)op; else flags |= 1; }
            steps++;
            if (op == 0)      { if (cons) cons[j * 5 + xi]++; run -= G;  i--; j--; nm++; }
            else if (op == 1) { if (cons) cons[j * 5 + xi]++; run += MM; i--; j--; nx++; }
            else if (op == 2) { if (cons) cons[j * 5 + 4]++;  run += IN; j--;      nd++; }
            else              { if (miss) miss[j * 4 + xi]++; run += IN; i--;      ni++; }
            if (j == 0) j = ulen;
            if (op >= 2) break;                             // left the diagonal: refill from the new cell
        }
        if (!(i > 0 && run > 0)) alive = false;
    }
    // 48 bytes as three 16-byte stores (the result array may be pinned host memory: few, wide PCIe writes)
    static_assert(sizeof(mtr_wdp_result) == 48, "mtr_wdp_result layout");
    int4 *o = reinterpret_cast<int4 *>(res);
    o[0] = make_int4(best, max_i, max_j, i);
    o[1] = make_int4(j, nm, nx, ni);
    o[2] = make_int4(nd, nm + nx + nd, steps, flags);
}

// stand-alone traceback (split mode: lets the fill kernels be timed alone), one thread per task
__global__ void __launch_bounds__(128)
wdp_traceback(const WdpTask *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ packed,
              const uint8_t *__restrict__ units, const uint8_t *dirs, const mtr_wdp_result *__restrict__ partial,
              mtr_wdp_result *results, void *aux)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ntasks) return;
    const WdpTask &t = tasks[idx];
    for (int p = 0; p < (t.n_param == 2 ? 2 : 1); p++) {
        const mtr_wdp_result *in = partial + t.result_idx + p;      // argmax left by the fill kernel (device memory)
        traceback_one(t, p, in->best, in->max_i, in->max_j, packed, units, dirs, results + t.result_idx + p, aux);
    }
}

// ---------------------------------------------------------------- fill kernel, int32 scores
// One warp-slot: 32/G tasks of one class, tasks[slot * (32/G) + group].  fused != 0: the group's first lane runs the
// traceback of its task as soon as the fill is done.
template <int G, int C, bool FUSED>
__device__ __forceinline__ void fill_slot_i32(const WdpTask *__restrict__ tasks, const int ntasks, const int slot,
                                              const uint32_t *__restrict__ packed, const uint8_t *__restrict__ units,
                                              uint8_t *dirs, mtr_wdp_result *results, void *aux)
{
    constexpr int JPW = 32 / G;                    // jobs per warp
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;                       // lane within its group
    const int grp = lane / G;
    {
        const int tidx = slot * JPW + grp;
        const bool have = tidx < ntasks;
        const WdpTask *tp = tasks + (have ? tidx : 0);
        const int rows = have ? tp->rows : 0;
        const int ulen = tp->ulen;
        const long long base0 = tp->base0;
        // int32 scores are kept x16 (the paired kernels: x4): the low two bits carry the direction tag as described in the
        // header, and the four low bits of an UNTAGGED score are free for the column tag of the argmax key, which then
        // costs one VIADDMNMX per cell (key = max(key, W + (15 - c)))
        const int g4 = 16 * tp->gain[0];
        const int cD = -16 * tp->mis[0] + 3;       // diagonal (mismatch) candidate, tag 3
        const int cL = -16 * tp->indel[0] + 2;     // left (deletion) candidate, tag 2
        const int cU = -16 * tp->indel[0] + 1;     // up (insertion) candidate, tag 1
        const int in4 = 16 * tp->indel[0];
        uint8_t *drow = dirs + tp->dir_off + (size_t)gl * (C / 4);
        const int dstride = tp->dir_stride;

        // per-lane match masks: bit c of eq[b] <=> unit base of my c-th column == b
        unsigned eq01 = 0, eq23 = 0;
        const int lu = (ulen - 1) / C, cu = (ulen - 1) % C;
        int sel[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = gl * C + c;
            if (j < ulen) {
                const int b = units[tp->unit_off + j];
                const unsigned bit = 1u << (c + ((b & 1) ? 16 : 0));
                if (b & 2) eq23 |= bit; else eq01 |= bit;
            }
            sel[c] = (gl == lu && c == cu) ? -1 : 0;
        }

        int Wp[C];
#pragma unroll
        for (int c = 0; c < C; c++) Wp[c] = 0;
        int dgin = 0;
        int best_v = 0, best_i = 0, best_c = 0;

        int maxrows = rows;
#pragma unroll
        for (int off = 16; off >= G; off >>= 1) maxrows = max(maxrows, __shfl_xor_sync(FULL, maxrows, off));

        unsigned xw = 0;
        // One row.  Ws = row i-1 of my columns (scores x4, untagged), Wd receives row i.
        auto row = [&](const int (&Ws)[C], int (&Wd)[C], const int i) {
            const bool act = i <= rows;
            const long long bi = base0 + i;
            if (act && (i == 1 || (bi & 15) == 0)) xw = packed[bi >> 4];
            const int xi = (int)((xw >> ((int)(bi & 15) * 2)) & 3u);
            const unsigned m = (((xi & 2) ? eq23 : eq01) >> ((xi & 1) * 16)) & ((1u << C) - 1u);

            // pass 1: every cell once, speculating that no left-carry enters the lane
            int R[C];                                       // tagged results (low 2 bits = direction code)
            int left = NEG_INF;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int dg = (c == 0) ? dgin : Ws[c - 1];
                const int a = __viaddmax_s32_relu(dg, cD, Ws[c] + cU);
                const int t = __viaddmax_s32(left, cL, a);
                const int r = (m & (1u << c)) ? dg + g4 : t;
                R[c] = r;
                Wd[c] = r & ~3;
                left = Wd[c];
            }
            // exact lane-boundary carries from the lane summaries: a lane passes its carry-in on (minus C
            // indels) only if it holds no match cell; iterate to the fixpoint (monotone, <= G rounds)
            const int out0 = Wd[C - 1];
            int out = out0, cin;
            for (;;) {
                cin = __shfl_up_sync(FULL, out, 1, G);
                if (gl == 0) cin = NEG_INF;
                const int nout = m ? out0 : max(out0, cin - in4 * C);
                const bool changed = nout != out;
                out = nout;
                if (!__any_sync(FULL, changed)) break;
            }
            // fix-up: cells before the lane's first match see the carry; cell c gets cin - (c+1) indels, tag "left"
            // (the carry decays by one indel per cell, so it usually dies within the first few cells: the cells
            // are visited in groups of four and the warp leaves as soon as no lane's carry is alive any more)
            bool live = !(m & 1u) && cin + cL > R[0];
            if (__any_sync(FULL, live)) {
                int cand = cin + cL;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 4) {
#pragma unroll
                    for (int c = c0; c < c0 + 4; c++) {
                        live = live && !(m & (1u << c)) && cand > R[c];
                        if (live) { R[c] = cand; Wd[c] = cand & ~3; }
                        cand -= in4;
                    }
                    if (c0 + 4 < C && !__any_sync(FULL, live)) break;
                }
            }
            // W[i][U] -> everybody (lane 0 needs it as next row's diagonal and for the j == 1 quirk)
            int mine = 0;
#pragma unroll
            for (int c = 0; c < C; c++) mine |= Wd[c] & sel[c];
            const int wU = __shfl_sync(FULL, mine, lu, G);

            // direction codes = tagged - untagged, packed 2 bits per cell with two multiply-add chains
            // (sum (R_c - W_c) 4^c is exact modulo 2^32); argmax key on the same pipe
            unsigned accR = 0, accW = 0;
            int key = 0;
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                accR = accR * 4u + (unsigned)R[c];
                accW = accW * 4u + (unsigned)Wd[c];
                key = __viaddmax_s32(Wd[c], 15 - c, key);
            }
            unsigned bits = accR - accW;
            // traceback at j == 1 tests "deletion" against W[i][0] == W[i][U] before insertion
            if (gl == 0 && (bits & 3u) == 1u && Wd[0] == wU - in4) bits ^= 3u;
            if (act) {
                uint8_t *p = drow + (size_t)(i - 1) * dstride;
                if (C == 4) *p = (uint8_t)bits;
                else if (C == 8) *(uint16_t *)p = (uint16_t)bits;
                else if (C == 12) { p[0] = (uint8_t)bits; p[1] = (uint8_t)(bits >> 8); p[2] = (uint8_t)(bits >> 16); }
                else *(uint32_t *)p = bits;
            }
            if (act && (key >> 4) > best_v) { best_v = key >> 4; best_i = i; best_c = 15 - (key & 15); }
            dgin = (gl == 0) ? wU : cin;
        };
        int Wq[C];
        for (int i = 1; i <= maxrows; i += 2) {
            row(Wp, Wq, i);
            if (i + 1 <= maxrows) row(Wq, Wp, i + 1);
        }

        // first row-major argmax over the group: larger value, then smaller row, then smaller column
        int best_j = gl * C + best_c + 1;
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            const int ov = __shfl_xor_sync(FULL, best_v, off);
            const int oi = __shfl_xor_sync(FULL, best_i, off);
            const int oj = __shfl_xor_sync(FULL, best_j, off);
            const bool take = (ov > best_v) || (ov == best_v && (oi < best_i || (oi == best_i && oj < best_j)));
            if (take) { best_v = ov; best_i = oi; best_j = oj; }
        }
        if (FUSED) {
            __syncwarp();                                   // the direction rows written by the other lanes of the group
            if (have && gl == 0)
                traceback_one(*tp, 0, best_v, best_v > 0 ? best_i : 0, best_v > 0 ? best_j : 0, packed, units, dirs,
                              results + tp->result_idx, aux);
            __syncwarp();
        } else if (have && gl == 0) {
            mtr_wdp_result *res = results + tp->result_idx;
            res->best = best_v;
            res->max_i = best_v > 0 ? best_i : 0;
            res->max_j = best_v > 0 ? best_j : 0;
        }
    }
}

template <int G, int C, bool FUSED>
__global__ void __launch_bounds__(128)
wdp_fill_i32(const WdpTask *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ packed,
             const uint8_t *__restrict__ units, uint8_t *dirs, mtr_wdp_result *results, int *__restrict__ counter,
             void *aux)
{
    constexpr int JPW = 32 / G;
    const int nslots = (ntasks + JPW - 1) / JPW;
    for (;;) {
        int slot = 0;
        if ((threadIdx.x & 31) == 0) slot = atomicAdd(counter, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= nslots) break;
        fill_slot_i32<G, C, FUSED>(tasks, ntasks, slot, packed, units, dirs, results, aux);
    }
}

// Latency mode: ONE launch for a whole small batch.  The tasks are sorted by latency class (10..15); a warp pulls the
// next slot of the shared queue, finds its class from the slot prefix sums and runs that class's fill (+ traceback).
struct WdpMultiParams { int task_begin[kLatencyClasses + 1]; int slot_begin[kLatencyClasses + 1]; };

template <bool FUSED>
__global__ void __launch_bounds__(128)
wdp_fill_multi(const WdpTask *__restrict__ tasks, const WdpMultiParams mp, const uint32_t *__restrict__ packed,
               const uint8_t *__restrict__ units, uint8_t *dirs, mtr_wdp_result *results, int *__restrict__ counter,
               void *aux)
{
    for (;;) {
        int slot = 0;
        if ((threadIdx.x & 31) == 0) slot = atomicAdd(counter, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= mp.slot_begin[kLatencyClasses]) break;
        int k = 0;
        while (slot >= mp.slot_begin[k + 1]) k++;
        const WdpTask *ct = tasks + mp.task_begin[k];
        const int cn = mp.task_begin[k + 1] - mp.task_begin[k];
        const int ls = slot - mp.slot_begin[k];
        switch (k) {
        case 0: fill_slot_i32<4, 4, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        case 1: fill_slot_i32<8, 4, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        case 2: fill_slot_i32<16, 4, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        case 3: fill_slot_i32<32, 4, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        case 4: fill_slot_i32<32, 8, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        default: fill_slot_i32<32, 16, FUSED>(ct, cn, ls, packed, units, dirs, results, aux); break;
        }
    }
}

// ---------------------------------------------------------------- fill kernel, paired int16x2 scores
// wrap_around_DP (wrap_around_DP.c:357-429) always runs the same window and unit under two penalty sets.  When
// 4 * gain * rows fits int16 both DPs share one register per cell (low half: set 0, high half: set 1) and every
// score operation is one VIADDMNMX.S16x2: the match mask, the carries' control flow and the wrap shuffle are
// common, only the constants differ per half.  Same tagged-score scheme and direction layout as wdp_fill_i32;
// the two direction matrices of a task lie dir_bytes apart.
#define P16_NEG 0x8ad08ad0u            /* (-30000, -30000): multiple of 4, far from wrapping */
#define P16_MIN 0x80008000u            /* (-32768, -32768): max(x + b, MIN) == x + b */

__device__ __forceinline__ unsigned pack2(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }

template <int G, int C, bool FUSED>
__device__ __forceinline__ void fill_slot_p16(const WdpTask *__restrict__ tasks, const int ntasks, const int slot,
                                              const uint32_t *__restrict__ packed, const uint8_t *__restrict__ units,
                                              uint8_t *dirs, mtr_wdp_result *results, void *aux)
{
    constexpr int JPW = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned UNTAG = 0xfffcfffcu;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const int grp = lane / G;
    {
        const int tidx = slot * JPW + grp;
        const bool have = tidx < ntasks;
        const WdpTask *tp = tasks + (have ? tidx : 0);
        const int rows = have ? tp->rows : 0;
        const int ulen = tp->ulen;
        const long long base0 = tp->base0;
        const unsigned g42 = pack2(4 * tp->gain[0], 4 * tp->gain[1]);
        const unsigned cD = pack2(-4 * tp->mis[0] + 3, -4 * tp->mis[1] + 3);
        const unsigned cL = pack2(-4 * tp->indel[0] + 2, -4 * tp->indel[1] + 2);
        const unsigned cU = pack2(-4 * tp->indel[0] + 1, -4 * tp->indel[1] + 1);
        const unsigned nin4 = pack2(-4 * tp->indel[0], -4 * tp->indel[1]);
        const unsigned nspan = pack2(-4 * tp->indel[0] * C, -4 * tp->indel[1] * C);
        uint8_t *drow = dirs + tp->dir_off + (size_t)gl * (C / 4);
        const long long dbytes = tp->dir_bytes;
        const int dstride = tp->dir_stride;

        unsigned eq01 = 0, eq23 = 0;
        const int lu = (ulen - 1) / C, cu = (ulen - 1) % C;
        unsigned sel[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = gl * C + c;
            if (j < ulen) {
                const int b = units[tp->unit_off + j];
                const unsigned bit = 1u << (c + ((b & 1) ? 16 : 0));
                if (b & 2) eq23 |= bit; else eq01 |= bit;
            }
            sel[c] = (gl == lu && c == cu) ? 0xffffffffu : 0u;
        }

        unsigned Wp[C], Wq[C];
#pragma unroll
        for (int c = 0; c < C; c++) Wp[c] = 0;
        unsigned dgin = 0;
        int best_v[2] = {0, 0}, best_i[2] = {0, 0}, best_c[2] = {0, 0};

        int maxrows = rows;
#pragma unroll
        for (int off = 16; off >= G; off >>= 1) maxrows = max(maxrows, __shfl_xor_sync(FULL, maxrows, off));

        unsigned xw = 0;
        auto row = [&](const unsigned (&Ws)[C], unsigned (&Wd)[C], const int i) {
            const bool act = i <= rows;
            const long long bi = base0 + i;
            if (act && (i == 1 || (bi & 15) == 0)) xw = packed[bi >> 4];
            const int xi = (int)((xw >> ((int)(bi & 15) * 2)) & 3u);
            const unsigned m = (((xi & 2) ? eq23 : eq01) >> ((xi & 1) * 16)) & ((1u << C) - 1u);

            unsigned R[C];
            unsigned left = P16_NEG;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const unsigned dg = (c == 0) ? dgin : Ws[c - 1];
                const unsigned up = __viaddmax_s16x2(Ws[c], cU, P16_MIN);
                const unsigned a = __viaddmax_s16x2_relu(dg, cD, up);
                const unsigned t = __viaddmax_s16x2(left, cL, a);
                const unsigned vm = __viaddmax_s16x2(dg, g42, P16_MIN);
                const unsigned r = (m & (1u << c)) ? vm : t;
                R[c] = r;
                Wd[c] = r & UNTAG;
                left = Wd[c];
            }
            const unsigned out0 = Wd[C - 1];
            unsigned out = out0, cin;
            for (;;) {
                cin = __shfl_up_sync(FULL, out, 1, G);
                if (gl == 0) cin = P16_NEG;
                const unsigned nout = m ? out0 : __viaddmax_s16x2(cin, nspan, out0);
                const bool changed = nout != out;
                out = nout;
                if (!__any_sync(FULL, changed)) break;
            }
            bool live = !(m & 1u) && __viaddmax_s16x2(cin, cL, R[0]) != R[0];
            if (__any_sync(FULL, live)) {
                unsigned cand = __viaddmax_s16x2(cin, cL, P16_MIN);
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 4) {
#pragma unroll
                    for (int c = c0; c < c0 + 4; c++) {
                        live = live && !(m & (1u << c));
                        const unsigned rn = __viaddmax_s16x2(live ? cand : P16_NEG, 0u, R[c]);
                        live = live && rn != R[c];
                        R[c] = rn;
                        Wd[c] = rn & UNTAG;
                        cand = __viaddmax_s16x2(cand, nin4, P16_MIN);
                    }
                    if (c0 + 4 < C && !__any_sync(FULL, live)) break;
                }
            }
            unsigned mine = 0;
#pragma unroll
            for (int c = 0; c < C; c++) mine |= Wd[c] & sel[c];
            const unsigned wU = __shfl_sync(FULL, mine, lu, G);

            // direction codes: (R - W) holds one 2-bit code per half; eight cells fit one accumulator
            unsigned acc0 = 0, acc1 = 0;
            int key0 = 0, key1 = 0;
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const unsigned code2 = R[c] - Wd[c];
                if (c >= 8) acc1 = acc1 * 4u + code2; else acc0 = acc0 * 4u + code2;
            }
            // argmax keys, one 32-bit key per half: the score (x4) in bits 16..30, 15 - c below; two cells per 3-input max
#pragma unroll
            for (int c = C - 1; c >= 1; c -= 2) {
                key0 = __vimax3_s32(key0, (int)(Wd[c] << 16) + (15 - c), (int)(Wd[c - 1] << 16) + (16 - c));
                key1 = __vimax3_s32(key1, (int)((Wd[c] & 0xffff0000u) | (unsigned)(15 - c)), (int)((Wd[c - 1] & 0xffff0000u) | (unsigned)(16 - c)));
            }
            unsigned bits0 = (acc0 & 0xffffu) | (acc1 << 16);
            unsigned bits1 = (acc0 >> 16) | (acc1 & 0xffff0000u);
            if (gl == 0) {
                // traceback at j == 1 tests "deletion" against W[i][0] == W[i][U] before insertion
                const unsigned wl = __viaddmax_s16x2(wU, nin4, P16_MIN);          // W[i][U] - IN per half
                if ((bits0 & 3u) == 1u && (Wd[0] & 0xffffu) == (wl & 0xffffu)) bits0 ^= 3u;
                if ((bits1 & 3u) == 1u && (Wd[0] >> 16) == (wl >> 16)) bits1 ^= 3u;
            }
            if (act) {
                uint8_t *p = drow + (size_t)(i - 1) * dstride;
                uint8_t *q = p + dbytes;
                if (C == 4) { *p = (uint8_t)bits0; *q = (uint8_t)bits1; }
                else if (C == 8) { *(uint16_t *)p = (uint16_t)bits0; *(uint16_t *)q = (uint16_t)bits1; }
                else if (C == 12) {
                    p[0] = (uint8_t)bits0; p[1] = (uint8_t)(bits0 >> 8); p[2] = (uint8_t)(bits0 >> 16);
                    q[0] = (uint8_t)bits1; q[1] = (uint8_t)(bits1 >> 8); q[2] = (uint8_t)(bits1 >> 16);
                } else { *(uint32_t *)p = bits0; *(uint32_t *)q = bits1; }
                if ((key0 >> 18) > best_v[0]) { best_v[0] = key0 >> 18; best_i[0] = i; best_c[0] = 15 - (key0 & 15); }
                if ((key1 >> 18) > best_v[1]) { best_v[1] = key1 >> 18; best_i[1] = i; best_c[1] = 15 - (key1 & 15); }
            }
            dgin = (gl == 0) ? wU : cin;
        };
        for (int i = 1; i <= maxrows; i += 2) {
            row(Wp, Wq, i);
            if (i + 1 <= maxrows) row(Wq, Wp, i + 1);
        }
        int my_v = 0, my_i = 0, my_j = 0;                   // lane gl == h keeps the argmax of penalty set h
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int bv = best_v[h], bi = best_i[h], bj = gl * C + best_c[h] + 1;
#pragma unroll
            for (int off = G / 2; off >= 1; off >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bv, off);
                const int oi = __shfl_xor_sync(FULL, bi, off);
                const int oj = __shfl_xor_sync(FULL, bj, off);
                const bool take = (ov > bv) || (ov == bv && (oi < bi || (oi == bi && oj < bj)));
                if (take) { bv = ov; bi = oi; bj = oj; }
            }
            if (gl == h) { my_v = bv; my_i = bv > 0 ? bi : 0; my_j = bv > 0 ? bj : 0; }
        }
        if (FUSED) {
            __syncwarp();                                   // the direction rows written by the other lanes of the group
            if (have && gl < 2) traceback_one(*tp, gl, my_v, my_i, my_j, packed, units, dirs, results + tp->result_idx + gl, aux);
            __syncwarp();
        } else if (have && gl < 2) {
            mtr_wdp_result *res = results + tp->result_idx + gl;
            res->best = my_v; res->max_i = my_i; res->max_j = my_j;
        }
    }
}

template <int G, int C, bool FUSED>
__global__ void __launch_bounds__(128)
wdp_fill_p16(const WdpTask *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ packed,
             const uint8_t *__restrict__ units, uint8_t *dirs, mtr_wdp_result *results, int *__restrict__ counter,
             void *aux)
{
    constexpr int JPW = 32 / G;
    const int nslots = (ntasks + JPW - 1) / JPW;
    for (;;) {
        int slot = 0;
        if ((threadIdx.x & 31) == 0) slot = atomicAdd(counter, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= nslots) break;
        fill_slot_p16<G, C, FUSED>(tasks, ntasks, slot, packed, units, dirs, results, aux);
    }
}

// Warp-cooperative traceback of one (task, penalty set): wrap_around_DP.c:288-333 (counts) and consensus.c:919-962
// (histograms).  The walk itself is sequential, but one thread chasing it pays an L2 round trip per step (read base,
// direction byte): here the 32 lanes fetch the next 32 rows at once -- lane q the read base of row i - q and 64 direction
// slots of that row around the column the diagonal will reach (the whole row when it has at most 64 slots) -- and the
// warp then walks through registers (one shuffle per step) until the path leaves the fetched rows or columns.
__device__ __forceinline__ void traceback_warp(const WdpTask &t, const int p, const int best, const int max_i, const int max_j,
                                               const uint32_t *__restrict__ packed, const uint8_t *__restrict__ units,
                                               const uint8_t *dirs, mtr_wdp_result *res, void *aux)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int G = t.gain[p], MM = t.mis[p], IN = t.indel[p];
    const int ulen = t.ulen, dstride = t.dir_stride, nslots = t.dir_stride * 4;
    const uint8_t *un = units + t.unit_off;
    const long long base0 = t.base0;
    const uint8_t *d = dirs + t.dir_off + (size_t)p * t.dir_bytes;
    int i = max_i, j = max_j, run = best;
    if (j == 0) j = ulen;                                   // wrap_around_DP.c:296
    int nm = 0, nx = 0, ni = 0, nd = 0, steps = 0, flags = 0;
    int *cons = nullptr, *miss = nullptr;
    if (t.mode == MTR_TB_CONSENSUS) {
        cons = (int *)aux + t.aux_off;
        miss = cons + (size_t)(ulen + 1) * 5;
    }
    while (i > 0 && run > 0) {
        // fetch rows i, i-1, ..., i-31
        const int top = i;
        const int rq = top - lane;
        int jq = j - lane;                                  // column the diagonal reaches in my row
        if (jq < 1) { jq = ulen - ((-jq) % ulen); }
        int c0 = 0;                                         // first slot of my 64-slot window
        if (nslots > 64) { c0 = ((jq - 1) - 32) & ~15; c0 = max(0, min(c0, nslots - 64)); }
        int xq = 0;
        unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        if (rq >= 1) {
            xq = read_base(packed, base0 + rq);
            const unsigned *row = reinterpret_cast<const unsigned *>(d + (size_t)(rq - 1) * dstride + (c0 >> 2));
            w0 = row[0];
            if (nslots > 16) w1 = row[1];
            if (nslots > 32) w2 = row[2];
            if (nslots > 48) w3 = row[3];
        }
        while (i > 0 && run > 0 && i > top - 32) {
            const int q = top - i;
            // A run of matches is a run of diagonal steps that ask nothing of the direction matrix (:306 tests
            // x_i == u_j first): lane q + k looks at the k-th cell down the diagonal from (i, j), one ballot finds
            // the length of the run, and the whole run is taken at once -- every lane adds its own histogram entry.
            // The running score loses G per step and the walk stops when it reaches zero: ceil(run / G) steps at most.
            int jl = j - (lane - q);
            if (jl < 1) jl = ulen - ((-jl) % ulen);
            bool mt = false;
            if (lane >= q && rq >= 1) mt = xq == (int)un[jl - 1];
            const unsigned bal = __ballot_sync(FULL, mt) >> q;
            int n = bal == FULL ? 32 : __ffs(~bal) - 1;
            if (n > 0) {
                if (G > 0) { const int kmax = (run + G - 1) / G; n = n < kmax ? n : kmax; }
                if (cons && lane >= q && lane < q + n) atomicAdd(&cons[jl * 5 + xq], 1);
                run -= n * G; i -= n; nm += n; steps += n;
                j -= n;
                if (j < 1) j = ulen - ((-j) % ulen);
                continue;
            }
            // one step that is not a match: the direction code of (i, j)
            const int xi = __shfl_sync(FULL, xq, q);
            const int s = j - 1 - __shfl_sync(FULL, c0, q);
            if (s < 0 || s >= 64) break;                       // the path left the fetched columns: fetch again from here
            const int wi = s >> 4;
            const unsigned mine = wi == 0 ? w0 : (wi == 1 ? w1 : (wi == 2 ? w2 : w3));
            const unsigned v = __shfl_sync(FULL, mine, q);
            const int code = (int)((v >> ((s & 15) * 2)) & 3u);
            if (code == 0) { flags |= 2; run = 0; break; }      // cannot happen: run > 0 means W[i][j] > 0
            steps++;
            if (code == 3)      { if (cons && lane == 0) atomicAdd(&cons[j * 5 + xi], 1); run += MM; i--; j--; nx++; }
            else if (code == 2) { if (cons && lane == 0) atomicAdd(&cons[j * 5 + 4], 1);  run += IN; j--;      nd++; }
            else                { if (miss && lane == 0) atomicAdd(&miss[j * 4 + xi], 1); run += IN; i--;      ni++; }
            if (j == 0) j = ulen;
        }
    }
    if (lane == 0) {
        int4 *o = reinterpret_cast<int4 *>(res);
        o[0] = make_int4(best, max_i, max_j, i);
        o[1] = make_int4(j, nm, nx, ni);
        o[2] = make_int4(nd, nm + nx + nd, steps, flags);
    }
}

// ---------------------------------------------------------------- resident engine: task lists built on the device
// The engine (eng_core.h) writes the sorted task array and the class boundaries in device memory; the host cannot know
// the counts, so these kernels are launched with a fixed persistent grid and read their range from class_begin[].
// Split traceback: the fill stores the argmax into results[], wdp_traceback_dev finishes the record in place.
// One kernel per score family (int32 | paired int16x2) covers all ten fill classes: the task list is sorted by segment =
// (family, rows bucket descending, class), a warp pulls the next slot of the family's queue, finds its segment by binary
// search in the slot prefix sums and runs that class's fill.  Longest tasks first across ALL classes, two launches per
// wave instead of twenty.
struct WdpSegs { const int *task; const int *slot; const unsigned long long *class_work; int nseg_family; };   // prefix sums over the 2 * 10 * kWdpRowBuckets (family, class, rows bucket) segments (+1); estimated work per (family, class)

// tb.fused: the warp that filled a slot also walks the tracebacks of its tasks (all 32 lanes on one task at a time) and
// counts the owners' outstanding results down, so that every task is complete on its own -- nothing waits for the
// longest fill of the launch.
struct WdpTb { void *aux; char *pending0; int pending_stride; int *pending_total; int fused; unsigned long long *prof; };   // prof: [fill ticks, traceback ticks, slot rows, slots, traceback rows] per family (5 + 5), nullptr = off

template <bool P16>
__global__ void __launch_bounds__(64, 8)
wdp_fill_family(const WdpTask *__restrict__ tasks, const WdpSegs sg, const uint32_t *__restrict__ packed,
                const uint8_t *__restrict__ units, uint8_t *dirs, mtr_wdp_result *results, int *__restrict__ counters, const WdpTb tb)
{
    constexpr int RB = kWdpRowBuckets;
    const int fam = P16 ? 1 : 0;
    const int seg0 = fam * 10 * RB;
    const int lane = threadIdx.x & 31;
    // The class this SM starts with: the classes share the SMs in proportion to their estimated work (class_work, written
    // by the engine's plan pass), so that the warps of one SM run one instantiation at a time.
    int cls = 0;
    {
        unsigned smid, nsm;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        long long total = 0;
        for (int c = 0; c < 10; c++) total += (long long)sg.class_work[fam * 10 + c];
        const long long target = (total * (2 * (long long)smid + 1)) / (2 * (long long)(nsm ? nsm : 1));
        long long acc = 0;
        for (int c = 0; c < 10; c++) {
            acc += (long long)sg.class_work[fam * 10 + c];
            cls = c;
            if (target < acc) break;
        }
    }
    for (int tried = 0; tried < 10; tried++, cls = (cls + 1) % 10) {
        const int sega = seg0 + cls * RB;
        const int slot0 = sg.slot[sega], slot1 = sg.slot[sega + RB];
        if (slot0 >= slot1) continue;
        const int jpw = cls < 3 ? 8 : (cls < 6 ? 4 : (cls < 8 ? 2 : 1));
        for (;;) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(counters + fam * 10 + cls, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0) + slot0;
            if (slot >= slot1) break;
            int lo = sega, hi = sega + RB;                      // largest segment with slot[seg] <= slot (empty segments share a value)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sg.slot[mid] <= slot) lo = mid; else hi = mid;
            }
            const WdpTask *ct = tasks + sg.task[lo];
            const int cn = sg.task[lo + 1] - sg.task[lo];
            const int ls = slot - sg.slot[lo];
            const long long pt0 = tb.prof ? clock64() : 0;
            if (P16) {
                switch (cls) {
                case 0: fill_slot_p16<4, 4, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 1: fill_slot_p16<4, 8, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 2: fill_slot_p16<4, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 3: fill_slot_p16<8, 8, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 4: fill_slot_p16<8, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 5: fill_slot_p16<8, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 6: fill_slot_p16<16, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 7: fill_slot_p16<16, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 8: fill_slot_p16<32, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                default: fill_slot_p16<32, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                }
            } else {
                switch (cls) {
                case 0: fill_slot_i32<4, 4, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 1: fill_slot_i32<4, 8, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 2: fill_slot_i32<4, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 3: fill_slot_i32<8, 8, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 4: fill_slot_i32<8, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 5: fill_slot_i32<8, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 6: fill_slot_i32<16, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 7: fill_slot_i32<16, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                case 8: fill_slot_i32<32, 12, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                default: fill_slot_i32<32, 16, false>(ct, cn, ls, packed, units, dirs, results, nullptr); break;
                }
            }
            const long long pt1 = tb.prof ? clock64() : 0;
            long long tb_rows = 0;
            if (tb.fused) {
                __syncwarp();                               // argmax records and direction rows written by the lanes of this warp
                if (jpw >= 4) {
                    // many short tasks in the slot: one lane per (task, penalty set), all of them at once
                    constexpr int NP = P16 ? 2 : 1;
                    const int g = lane / NP, p = lane % NP;
                    const int tidx = ls * jpw + g;
                    if (g < jpw && tidx < cn) {
                        const WdpTask &t = ct[tidx];
                        if (p < (int)t.n_param) {
                            const mtr_wdp_result in = results[t.result_idx + p];
                            traceback_one(t, p, in.best, in.max_i, in.max_j, packed, units, dirs, results + t.result_idx + p, tb.aux);
                            tb_rows = t.rows;
                            if (tb.pending0) {
                                __threadfence();
                                atomicSub(reinterpret_cast<int *>(tb.pending0 + (size_t)(t.result_idx >> kWdpOwnerShift) * tb.pending_stride), 1);
                                atomicSub(tb.pending_total, 1);
                            }
                        }
                    }
                    __syncwarp();
                    if (tb.prof) for (int off = 16; off >= 1; off >>= 1) tb_rows += __shfl_xor_sync(0xffffffffu, tb_rows, off);
                } else {
                    // one or two long tasks: the whole warp walks one traceback at a time
                    for (int g = 0; g < jpw; g++) {
                        const int tidx = ls * jpw + g;
                        if (tidx >= cn) break;
                        const WdpTask &t = ct[tidx];
                        const int np = (int)t.n_param;
                        tb_rows += (long long)t.rows * np;
                        for (int p = 0; p < np; p++) {
                            const mtr_wdp_result in = results[t.result_idx + p];
                            __syncwarp();
                            traceback_warp(t, p, in.best, in.max_i, in.max_j, packed, units, dirs, results + t.result_idx + p, tb.aux);
                            __syncwarp();
                        }
                        // the owner of the results (an engine chain: four result slots each) counts its outstanding results down
                        if (tb.pending0 && lane == 0) {
                            __threadfence();
                            atomicSub(reinterpret_cast<int *>(tb.pending0 + (size_t)(t.result_idx >> kWdpOwnerShift) * tb.pending_stride), np);
                            atomicSub(tb.pending_total, np);
                        }
                    }
                }
            }
            if (tb.prof && lane == 0) {
                unsigned long long *pr = tb.prof + (P16 ? 5 : 0);
                const long long pt2 = clock64();
                atomicAdd(pr + 0, (unsigned long long)(pt1 - pt0)); atomicAdd(pr + 1, (unsigned long long)(pt2 - pt1));
                atomicAdd(pr + 2, (unsigned long long)ct[ls * jpw].rows); atomicAdd(pr + 3, 1ull); atomicAdd(pr + 4, (unsigned long long)tb_rows);
            }
        }
    }
}

// one warp per (task, penalty set), pulled from a queue in task order (longest tasks of every class first)
__global__ void __launch_bounds__(128)
wdp_traceback_dev(const WdpTask *__restrict__ tasks, const int *__restrict__ class_begin, const uint32_t *__restrict__ packed,
                  const uint8_t *__restrict__ units, const uint8_t *dirs, mtr_wdp_result *results, void *aux, int *__restrict__ head,
                  char *pending0, int pending_stride, int *pending_total)
{
    const int total = 2 * class_begin[WDP_NCLASS];
    for (;;) {
        int idx = 0;
        if ((threadIdx.x & 31) == 0) idx = atomicAdd(head, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= total) break;
        const WdpTask &t = tasks[idx >> 1];
        const int p = idx & 1;
        if (p >= (int)t.n_param) continue;
        const mtr_wdp_result in = results[t.result_idx + p];
        __syncwarp();
        traceback_warp(t, p, in.best, in.max_i, in.max_j, packed, units, dirs, results + t.result_idx + p, aux);
        __syncwarp();
        // the owner of the result (an engine chain: four result slots each) counts its outstanding results down
        if (pending0 && (threadIdx.x & 31) == 0) {
            __threadfence();
            atomicSub(reinterpret_cast<int *>(pending0 + (size_t)(t.result_idx >> kWdpOwnerShift) * pending_stride), 1);
            atomicSub(pending_total, 1);
        }
    }
}

// enqueues the two family fill kernels (the paired one on the side stream); the traceback runs inside them (L.fused) or
// as a kernel of its own behind both
cudaError_t wdp_launch_dev(const WdpDevLaunch &L, cudaStream_t s)
{
    cudaError_t e;
    WdpSegs sg;
    sg.task = L.seg_task; sg.slot = L.seg_slot; sg.class_work = L.class_share; sg.nseg_family = L.nseg_family;
    WdpTb tb;
    tb.aux = L.aux; tb.pending0 = L.pending0; tb.pending_stride = L.pending_stride; tb.pending_total = L.pending_total; tb.fused = L.fused; tb.prof = L.prof;
    const int b0 = L.blocks_i32 >= 0 ? L.blocks_i32 : L.blocks, b1 = L.blocks_p16 >= 0 ? L.blocks_p16 : L.blocks;
    const bool fork = L.n_side > 0 && b0 > 0 && b1 > 0;
    cudaStream_t ps = fork ? L.side[0] : s;
    if (fork) {
        if ((e = cudaEventRecord(L.fork, s)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(ps, L.fork, 0)) != cudaSuccess) return e;
    }
    if (b0 > 0) wdp_fill_family<false><<<b0, 64, L.fill_smem, s>>>(L.tasks, sg, L.packed, L.units, L.dirs, L.results, L.counters, tb);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (b1 > 0) wdp_fill_family<true><<<b1, 64, L.fill_smem, ps>>>(L.tasks, sg, L.packed, L.units, L.dirs, L.results, L.counters, tb);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (fork) {
        if ((e = cudaEventRecord(L.join[0], ps)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(s, L.join[0], 0)) != cudaSuccess) return e;
    }
    if (!L.fused)
        wdp_traceback_dev<<<L.blocks, 128, 0, s>>>(L.tasks, L.class_begin, L.packed, L.units, L.dirs, L.results, L.aux, L.counters + 20, L.pending0, L.pending_stride, L.pending_total);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- host side
constexpr int kCounterGens = 128;      // ring of counter sets: one memset per kCounterGens launches instead of one per launch

static inline const uint8_t *d_units_of(const WdpState &w) { return (const uint8_t *)w.d_tasks.p + w.units_dev_off; }
// The thread that finishes a task's traceback stores the 48-byte result straight into the pinned host buffer (mapped
// into the device address space): no device->host copy is queued behind the kernels, where it would sit at the head
// of a copy-engine queue and hold up the copies of every other lane until this lane's kernels are done.
// Small (latency-class) batches run the traceback inside the fill kernel: one launch per batch.  Large batches keep
// a separate traceback kernel -- a lane that walks a 10 k-step traceback would hold its whole warp back from filling.
static inline bool fused_now(const WdpState &w) { return w.fused_tb && w.latency; }
static inline mtr_wdp_result *results_of(const WdpState &w) { return (mtr_wdp_result *)(fused_now(w) ? w.h_results.p : w.d_results.p); }

template <int G, int C>
static void launch_fill(mtr_ctx *ctx, int *counter, const WdpTask *d_tasks, int ntasks, cudaStream_t s)
{
    const int jpw = 32 / G;
    const int nslots = (ntasks + jpw - 1) / jpw;
    int blocks = (nslots + 3) / 4;
    blocks = std::min(blocks, ctx->n_sm * 8);
    if (fused_now(ctx->wdp))
        wdp_fill_i32<G, C, true><<<blocks, 128, 0, s>>>(d_tasks, ntasks, (const uint32_t *)ctx->d_packed.p, d_units_of(ctx->wdp),
                                                        (uint8_t *)ctx->wdp.d_dirs.p, results_of(ctx->wdp), counter, ctx->wdp.d_aux.p);
    else
        wdp_fill_i32<G, C, false><<<blocks, 128, 0, s>>>(d_tasks, ntasks, (const uint32_t *)ctx->d_packed.p, d_units_of(ctx->wdp),
                                                         (uint8_t *)ctx->wdp.d_dirs.p, results_of(ctx->wdp), counter, ctx->wdp.d_aux.p);
}

template <int G, int C>
static void launch_fill_p16(mtr_ctx *ctx, int *counter, const WdpTask *d_tasks, int ntasks, cudaStream_t s)
{
    const int jpw = 32 / G;
    const int nslots = (ntasks + jpw - 1) / jpw;
    int blocks = (nslots + 3) / 4;
    blocks = std::min(blocks, ctx->n_sm * 8);
    if (fused_now(ctx->wdp))
        wdp_fill_p16<G, C, true><<<blocks, 128, 0, s>>>(d_tasks, ntasks, (const uint32_t *)ctx->d_packed.p, d_units_of(ctx->wdp),
                                                        (uint8_t *)ctx->wdp.d_dirs.p, results_of(ctx->wdp), counter, ctx->wdp.d_aux.p);
    else
        wdp_fill_p16<G, C, false><<<blocks, 128, 0, s>>>(d_tasks, ntasks, (const uint32_t *)ctx->d_packed.p, d_units_of(ctx->wdp),
                                                         (uint8_t *)ctx->wdp.d_dirs.p, results_of(ctx->wdp), counter, ctx->wdp.d_aux.p);
}

int wdp_upload_impl(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                    int64_t aux_bytes, bool sync)
{
    WdpState &w = ctx->wdp;
    w.uploaded = false;
    static const bool prof = getenv("MTR_PROFILE") != nullptr;
    auto wall = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
    const double tp0 = prof ? wall() : 0;
    if (n_jobs < 0 || (n_jobs > 0 && (!jobs || !units))) { mtr_set_error(ctx, "wdp_upload: null argument"); return MTR_EINVAL; }
    if (n_jobs > 0 && ctx->n_reads == 0) { mtr_set_error(ctx, "wdp_upload: no resident read batch"); return MTR_EINVAL; }
    w.tasks.clear();
    w.tasks.reserve((size_t)n_jobs * 2);
    w.n_jobs = n_jobs;
    w.n_results = n_jobs * 2;
    w.cells = 0; w.slot_cells = 0;
    std::vector<int> cls;
    cls.reserve((size_t)n_jobs * 2);
    // latency classes if the whole batch fits the resident warps of the GPU even with one warp-row per job
    long long lat_warps = 0;
    for (int jn = 0; jn < n_jobs; jn++) {
        const int k = class_of(std::max(1, std::min(jobs[jn].ulen, 499)), 1);
        lat_warps += (long long)jobs[jn].n_param * kClasses[k].G;
    }
    int latency = (lat_warps / 32) <= (long long)ctx->n_sm * 24 ? 1 : 0;
    w.unused_results.clear();
    if (const char *e = getenv("MTR_WDP_MODE")) latency = strcmp(e, "latency") == 0 ? 1 : (strcmp(e, "throughput") == 0 ? 0 : latency);
    const bool pair_ok = !getenv("MTR_NO_PAIRED");
    for (int jn = 0; jn < n_jobs; jn++) {
        const mtr_wdp_job &j = jobs[jn];
        if (j.read < 0 || j.read >= ctx->n_reads || j.ulen < 1 || j.ulen >= 500 || j.rows < 0 ||
            (j.n_param != 1 && j.n_param != 2) || j.mode > MTR_TB_PATH || (j.mode != MTR_TB_COUNTS && j.n_param != 1) ||
            j.unit_off < 0 || (int64_t)j.unit_off + j.ulen > units_len || j.first < -1 ||
            (int64_t)j.first + j.rows > (int64_t)ctx->len[j.read] + 1) {
            mtr_set_error(ctx, "wdp_upload: job %d is malformed (read %d first %d rows %d ulen %d n_param %d mode %d)",
                          jn, j.read, j.first, j.rows, j.ulen, j.n_param, j.mode);
            return MTR_EINVAL;
        }
        if ((int64_t)j.ulen * j.rows > 200000000LL) {       // WrapDPsize, mTR.h:51 / handle_one_read.c:89
            mtr_set_error(ctx, "wdp_upload: job %d exceeds WrapDPsize", jn);
            return MTR_ERANGE;
        }
        if (j.mode == MTR_TB_CONSENSUS && (j.aux_off < 0 || (j.aux_off + (int64_t)(j.ulen + 1) * 9) * 4 > aux_bytes)) {
            mtr_set_error(ctx, "wdp_upload: job %d consensus block outside aux", jn);
            return MTR_EINVAL;
        }
        if (j.mode == MTR_TB_PATH && (j.aux_off < 0 || j.aux_cap < 0 || j.aux_off + j.aux_cap > aux_bytes)) {
            mtr_set_error(ctx, "wdp_upload: job %d path block outside aux", jn);
            return MTR_EINVAL;
        }
        // both penalty sets in one int16x2 task when every score (x4, plus tag) fits 15 bits
        const int gmax = std::max((int)j.gain[0], (int)j.gain[1]);
        if (!latency && pair_ok && j.n_param == 2 && j.mode == MTR_TB_COUNTS && 4LL * gmax * j.rows <= 32760 &&
            j.indel[0] * 64 < 8000 && j.indel[1] * 64 < 8000 && j.mis[0] < 64 && j.mis[1] < 64) {
            WdpTask t;
            memset(&t, 0, sizeof t);
            t.base0 = ctx->word_off[j.read] * 16 + j.first;
            t.rows = j.rows; t.ulen = j.ulen; t.unit_off = j.unit_off;
            for (int p = 0; p < 2; p++) { t.gain[p] = j.gain[p]; t.mis[p] = j.mis[p]; t.indel[p] = j.indel[p]; }
            t.n_param = 2; t.mode = j.mode;
            t.result_idx = jn * 2;
            const int k = class_of(j.ulen, 0);
            t.dir_stride = kClasses[k].G * kClasses[k].C / 4;
            t.dir_bytes = (long long)t.rows * t.dir_stride;
            cls.push_back(kPairedBase + k);
            w.tasks.push_back(t);
            w.cells += 2LL * j.rows * j.ulen;
            w.slot_cells += 2LL * j.rows * kClasses[k].G * kClasses[k].C;
            continue;
        }
        if (j.n_param == 1) w.unused_results.push_back(jn * 2 + 1);
        for (int p = 0; p < j.n_param; p++) {
            WdpTask t;
            memset(&t, 0, sizeof t);
            t.base0 = ctx->word_off[j.read] * 16 + j.first;
            t.rows = j.rows; t.ulen = j.ulen; t.unit_off = j.unit_off;
            t.gain[0] = j.gain[p]; t.mis[0] = j.mis[p]; t.indel[0] = j.indel[p];
            t.n_param = 1; t.mode = j.mode;
            t.aux_off = j.aux_off; t.aux_cap = j.aux_cap;
            t.result_idx = jn * 2 + p;
            const int k = class_of(j.ulen, latency);
            t.dir_stride = kClasses[k].G * kClasses[k].C / 4;
            t.dir_bytes = (long long)t.rows * t.dir_stride;
            cls.push_back(k);
            w.tasks.push_back(t);
            w.cells += (long long)j.rows * j.ulen;
            w.slot_cells += (long long)j.rows * kClasses[k].G * kClasses[k].C;
        }
    }
    // sort by (class, rows descending): similar jobs share a warp, long jobs start first
    const int nt = (int)w.tasks.size();
    std::vector<int> order(nt);
    for (int i = 0; i < nt; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (cls[a] != cls[b]) return cls[a] < cls[b];
        if (w.tasks[a].rows != w.tasks[b].rows) return w.tasks[a].rows > w.tasks[b].rows;
        return a < b;
    });
    std::vector<WdpTask> sorted(nt);
    for (int k = 0; k <= WDP_NCLASS; k++) w.class_begin[k] = 0;
    long long off = 0;
    for (int i = 0; i < nt; i++) {
        sorted[i] = w.tasks[order[i]];
        sorted[i].dir_off = off;
        off += (sorted[i].dir_bytes * sorted[i].n_param + 15) & ~15LL;
        w.class_begin[cls[order[i]] + 1]++;
    }
    for (int k = 0; k < WDP_NCLASS; k++) w.class_begin[k + 1] += w.class_begin[k];
    w.tasks.swap(sorted);
    w.dir_total = off;
    w.aux_bytes = aux_bytes;
    w.latency = latency;

    const double tp1 = prof ? wall() : 0;
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    // tasks and units travel as one block: [WdpTask x nt | units]
    w.units_dev_off = sizeof(WdpTask) * (size_t)std::max(nt, 1);
    const size_t in_bytes = w.units_dev_off + (size_t)std::max<int64_t>(units_len, 1);
    MTR_CUDA(ctx, w.d_tasks.reserve(in_bytes));
    // the direction matrices of a batch run to gigabytes and every regrowth is a device-wide cudaFree + cudaMalloc:
    // grow in big steps so that a context reallocates a handful of times in its life
    if ((size_t)w.dir_total > w.d_dirs.cap)
        MTR_CUDA(ctx, w.d_dirs.reserve_exact((size_t)std::max<long long>(w.dir_total + std::min<long long>(w.dir_total / 2, 2LL << 30), 64 << 20)));
    MTR_CUDA(ctx, w.d_results.reserve(sizeof(mtr_wdp_result) * (size_t)std::max(w.n_results, 1)));
    MTR_CUDA(ctx, w.h_results.reserve(sizeof(mtr_wdp_result) * (size_t)std::max(w.n_results, 1)));
    MTR_CUDA(ctx, w.d_aux.reserve((size_t)std::max<int64_t>(aux_bytes, 16)));
    MTR_CUDA(ctx, w.d_counters.reserve(sizeof(int) * WDP_NCLASS * kCounterGens));
    const double tp2 = prof ? wall() : 0;
    if (nt > 0) {
        MTR_CUDA(ctx, w.h_tasks.reserve(in_bytes));
        memcpy(w.h_tasks.p, w.tasks.data(), sizeof(WdpTask) * (size_t)nt);
        memcpy((char *)w.h_tasks.p + w.units_dev_off, units, (size_t)units_len);
        MTR_CUDA(ctx, cudaMemcpyAsync(w.d_tasks.p, w.h_tasks.p, w.units_dev_off + (size_t)units_len, cudaMemcpyHostToDevice, ctx->main_stream));
    }
    if (sync) MTR_CUDA(ctx, mtr_sync(ctx));
    if (prof) { const double tp3 = wall(); ctx->prof_up[0] += tp1 - tp0; ctx->prof_up[1] += tp2 - tp1; ctx->prof_up[2] += tp3 - tp2; }
    w.uploaded = true;
    return MTR_OK;
}

// reads the CUDA-event times of the last launch (after a sync)
static int wdp_read_times(mtr_ctx *ctx)
{
    float f0 = 0, f1 = 0;
    MTR_CUDA(ctx, cudaEventElapsedTime(&f0, ctx->ev[0], ctx->ev[1]));
    MTR_CUDA(ctx, cudaEventElapsedTime(&f1, ctx->ev[1], ctx->ev[2]));
    ctx->stats.wdp_fill_ms = f0;
    ctx->stats.wdp_tb_ms = f1;
    return MTR_OK;
}

int wdp_launch_impl(mtr_ctx *ctx, bool sync)
{
    WdpState &w = ctx->wdp;
    if (!w.uploaded) { mtr_set_error(ctx, "wdp_launch: nothing uploaded"); return MTR_EINVAL; }
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nt = (int)w.tasks.size();
    ctx->stats.launches = 0;
    ctx->stats.wdp_cells = w.cells;
    ctx->stats.wdp_slot_cells = w.slot_cells;
    ctx->stats.wdp_dir_bytes = w.dir_total;
    cudaStream_t ms = ctx->main_stream;
    // slot counters: a fresh, already-zero set per launch
    if (w.counter_gen % kCounterGens == 0)
        MTR_CUDA(ctx, cudaMemsetAsync(w.d_counters.p, 0, sizeof(int) * WDP_NCLASS * kCounterGens, ms));
    int *counters = (int *)w.d_counters.p + (size_t)(w.counter_gen % kCounterGens) * WDP_NCLASS;
    w.counter_gen++;
    if (w.aux_bytes > 0) MTR_CUDA(ctx, cudaMemsetAsync(w.d_aux.p, 0, (size_t)w.aux_bytes, ms));
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[0], ms));
    if (w.latency && nt > 0) {
        // one launch for the whole batch
        WdpMultiParams mp;
        int slots = 0;
        for (int q = 0; q <= kLatencyClasses; q++) {
            mp.task_begin[q] = w.class_begin[kThroughputClasses + q] - w.class_begin[kThroughputClasses];
            mp.slot_begin[q] = slots;
            if (q < kLatencyClasses) {
                const int n = w.class_begin[kThroughputClasses + q + 1] - w.class_begin[kThroughputClasses + q];
                const int jpw = 32 / kClasses[kThroughputClasses + q].G;
                slots += (n + jpw - 1) / jpw;
            }
        }
        const int blocks = std::min((slots + 3) / 4, ctx->n_sm * 4);
        if (fused_now(w))
            wdp_fill_multi<true><<<blocks, 128, 0, ms>>>((const WdpTask *)w.d_tasks.p + w.class_begin[kThroughputClasses], mp,
                                                         (const uint32_t *)ctx->d_packed.p, d_units_of(w), (uint8_t *)w.d_dirs.p,
                                                         results_of(w), counters, w.d_aux.p);
        else
            wdp_fill_multi<false><<<blocks, 128, 0, ms>>>((const WdpTask *)w.d_tasks.p + w.class_begin[kThroughputClasses], mp,
                                                          (const uint32_t *)ctx->d_packed.p, d_units_of(w), (uint8_t *)w.d_dirs.p,
                                                          results_of(w), counters, w.d_aux.p);
        MTR_CUDA(ctx, cudaGetLastError());
        ctx->stats.launches++;
    } else {
        for (int k = 0; k < WDP_NCLASS; k++) {
            const int n = w.class_begin[k + 1] - w.class_begin[k];
            if (n == 0) continue;
            cudaStream_t s = ctx->stream[k];
            MTR_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev[0], 0));
            const WdpTask *dt = (const WdpTask *)w.d_tasks.p + w.class_begin[k];
            int *cnt = counters + k;
            switch (k) {
            case 0: launch_fill<4, 4>(ctx, cnt, dt, n, s); break;
            case 1: launch_fill<4, 8>(ctx, cnt, dt, n, s); break;
            case 2: launch_fill<4, 12>(ctx, cnt, dt, n, s); break;
            case 3: launch_fill<8, 8>(ctx, cnt, dt, n, s); break;
            case 4: launch_fill<8, 12>(ctx, cnt, dt, n, s); break;
            case 5: launch_fill<8, 16>(ctx, cnt, dt, n, s); break;
            case 6: launch_fill<16, 12>(ctx, cnt, dt, n, s); break;
            case 7: launch_fill<16, 16>(ctx, cnt, dt, n, s); break;
            case 8: launch_fill<32, 12>(ctx, cnt, dt, n, s); break;
            case 9: launch_fill<32, 16>(ctx, cnt, dt, n, s); break;
            case 10: launch_fill<4, 4>(ctx, cnt, dt, n, s); break;
            case 11: launch_fill<8, 4>(ctx, cnt, dt, n, s); break;
            case 12: launch_fill<16, 4>(ctx, cnt, dt, n, s); break;
            case 13: launch_fill<32, 4>(ctx, cnt, dt, n, s); break;
            case 14: launch_fill<32, 8>(ctx, cnt, dt, n, s); break;
            case 15: launch_fill<32, 16>(ctx, cnt, dt, n, s); break;
            case 16: launch_fill_p16<4, 4>(ctx, cnt, dt, n, s); break;
            case 17: launch_fill_p16<4, 8>(ctx, cnt, dt, n, s); break;
            case 18: launch_fill_p16<4, 12>(ctx, cnt, dt, n, s); break;
            case 19: launch_fill_p16<8, 8>(ctx, cnt, dt, n, s); break;
            case 20: launch_fill_p16<8, 12>(ctx, cnt, dt, n, s); break;
            case 21: launch_fill_p16<8, 16>(ctx, cnt, dt, n, s); break;
            case 22: launch_fill_p16<16, 12>(ctx, cnt, dt, n, s); break;
            case 23: launch_fill_p16<16, 16>(ctx, cnt, dt, n, s); break;
            case 24: launch_fill_p16<32, 12>(ctx, cnt, dt, n, s); break;
            case 25: launch_fill_p16<32, 16>(ctx, cnt, dt, n, s); break;
            default: mtr_set_error(ctx, "wdp_launch: class %d has no kernel", k); return MTR_EINVAL;
            }
            MTR_CUDA(ctx, cudaGetLastError());
            ctx->stats.launches++;
            MTR_CUDA(ctx, cudaEventRecord(ctx->class_done[k], s));
            MTR_CUDA(ctx, cudaStreamWaitEvent(ms, ctx->class_done[k], 0));
        }
    }
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[1], ms));
    if (nt > 0 && !fused_now(w)) {
        wdp_traceback<<<(nt + 127) / 128, 128, 0, ms>>>((const WdpTask *)w.d_tasks.p, nt, (const uint32_t *)ctx->d_packed.p,
                                                        d_units_of(w), (const uint8_t *)w.d_dirs.p,
                                                        (const mtr_wdp_result *)w.d_results.p, (mtr_wdp_result *)w.h_results.p, w.d_aux.p);
        MTR_CUDA(ctx, cudaGetLastError());
        ctx->stats.launches++;
    }
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[2], ms));
    if (sync) {
        MTR_CUDA(ctx, mtr_sync(ctx));
        return wdp_read_times(ctx);
    }
    return MTR_OK;
}

int wdp_download_impl(mtr_ctx *ctx, mtr_wdp_result *results, void *aux, int64_t aux_bytes, bool read_times)
{
    WdpState &w = ctx->wdp;
    if (!w.uploaded) { mtr_set_error(ctx, "wdp_download: nothing uploaded"); return MTR_EINVAL; }
    if (aux_bytes > w.aux_bytes) { mtr_set_error(ctx, "wdp_download: aux_bytes larger than uploaded"); return MTR_EINVAL; }
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    // the results are already on their way into h_results; a copy is only issued once the kernels are done
    if (aux && aux_bytes > 0) {
        MTR_CUDA(ctx, mtr_sync(ctx));
        MTR_CUDA(ctx, cudaMemcpyAsync(aux, w.d_aux.p, (size_t)aux_bytes, cudaMemcpyDeviceToHost, ctx->main_stream));
    }
    MTR_CUDA(ctx, mtr_sync(ctx));
    if (w.n_results > 0) memcpy(results, w.h_results.p, sizeof(mtr_wdp_result) * (size_t)w.n_results);
    for (int idx : w.unused_results) memset(results + idx, 0, sizeof(mtr_wdp_result));     // second slot of a one-set job
    if (read_times) return wdp_read_times(ctx);
    return MTR_OK;
}
