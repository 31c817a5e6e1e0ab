// di.cu -- K1/K2: directional index of every read of the resident batch, for sm_100a.
//
// Replaces fill_directional_index_with_end (/root/reference/fill_directional_index.c:549-602):
//   K1  di_codes   : 2-bit packed read + MT19937 flanks (+ stale tail) -> per-position k-mer codes for
//                    k = 1, 3, 5  (init_inputString_surrounded_by_random_seq, :137-169)
//   K2  di_slide   : per (read, k, w, chunk): slides two adjacent w-wide k-mer histograms and emits one
//                    distance stream  D(q) = |H(q) - H(q+w)|_1  (Manhattan, :171-295) or the Pearson
//                    correlation P(q) of the two histograms (:298-450).  The reference's three windows are
//                    two consecutive samples of this stream (v1(p) = v0(p+w)):
//                        Manhattan  tmp[p] = (D(p-w) - D(p)) / (2w)        Pearson  tmp[p] = P(p) - P(p-w)
//                    Every operand is an exact integer, so the fp64 results are bit-identical to the CPU's.
//   K2b di_merge   : per read, the sequential local-max / local-min merge of the passes (:467-503), the
//                    unshift (:587-597) and remove_redundant_ranges (:505-546).
#include <algorithm>
#include <cstring>
#include "mtr_internal.h"

#define MT_TABLE_LEN 1300000
#define FULLMASK 0xffffffffu

struct DiRead {
    long long word_off;     // first word of the packed read
    long long code_off;     // offset of this read in the S1/S3/S5 arrays
    long long stale_off;    // offset into the stale array (uint16)
    long long pos_off;      // offset into the di/end/w output arrays
    long long work_off;     // offset into the merge work arrays (N entries)
    int len, r, N, M;       // read length, flank length, N = L + 2r, M = coded positions kept
    int nstale;
    int pass_begin;         // first entry of this read in the pass table
    int npass;
};

struct DiPass {
    long long stream_off;   // offset into the distance stream (int32 or double)
    int k, w, steps, nq;    // steps = L + r - w - k + 1, nq = steps + w samples of the stream
};

struct DiTask {             // one chunk of one pass
    int read, pass, q0, nq;
};

struct DiState {
    DevBuf d_mt, d_reads, d_passes, d_tasks[3], d_s1, d_s3, d_s5, d_stream, d_stale;
    DevBuf d_di, d_end, d_w, d_work_di, d_work_end, d_work_w, d_work_tmp;
    bool mt_ready = false;
};

void di_state_free(mtr_ctx *ctx)
{
    if (!ctx->di) return;
    DiState *d = ctx->di;
    d->d_mt.release(); d->d_reads.release(); d->d_passes.release();
    for (int i = 0; i < 3; i++) d->d_tasks[i].release();
    d->d_s1.release(); d->d_s3.release(); d->d_s5.release(); d->d_stream.release(); d->d_stale.release();
    d->d_di.release(); d->d_end.release(); d->d_w.release();
    d->d_work_di.release(); d->d_work_end.release(); d->d_work_w.release(); d->d_work_tmp.release();
    delete d;
    ctx->di = nullptr;
}

// MT19937 with seed 0, reduced mod 4: the flank bases (MT.h:65-78,110-145; fill_directional_index.c:129-141).
static void mt_base_table(std::vector<uint8_t> &out)
{
    out.resize(MT_TABLE_LEN);
    uint32_t st[624];
    st[0] = 0;
    for (int i = 1; i < 624; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    int pos = 624;
    for (size_t t = 0; t < out.size(); t++) {
        if (pos == 624) {
            for (int i = 0; i < 624; i++) {
                const uint32_t y = (st[i] & 0x80000000u) | (st[(i + 1) % 624] & 0x7fffffffu);
                st[i] = st[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = st[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        out[t] = (uint8_t)(y & 3u);
    }
}

// ---------------------------------------------------------------- K1: k-mer codes
__device__ __forceinline__ int padded_base(const DiRead &rd, const uint32_t *__restrict__ packed,
                                           const uint8_t *__restrict__ mt, int i)
{
    // rand(r) | read | rand(r) | first-loop randoms, fill_directional_index.c:143-156
    const int L = rd.len, r = rd.r;
    const int c = min(L + 4 * r, 1000000);
    if (i < r) return mt[c + i];
    if (i < r + L) {
        const int b = i - r;
        return (int)((packed[rd.word_off + (b >> 4)] >> ((b & 15) * 2)) & 3u);
    }
    if (i < rd.N) return mt[c + r + (i - r - L)];
    return mt[i];
}

__global__ void __launch_bounds__(256)
di_codes(const DiRead *__restrict__ reads, int n_reads, const uint32_t *__restrict__ packed,
         const uint8_t *__restrict__ mt, const uint16_t *__restrict__ stale,
         uint8_t *__restrict__ s1, uint8_t *__restrict__ s3, uint16_t *__restrict__ s5)
{
    const int rdi = blockIdx.x;
    const DiRead rd = reads[rdi];
    const int written = min(rd.len + 4 * rd.r, 1000000);       // area re-initialised for this read (:143)
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < rd.M; i += gridDim.y * blockDim.x) {
        int c1, c3, c5;
        if (i >= written) {
            // beyond the area this read re-initialises: what an earlier, longer read left there (H3)
            const int t = i - written;
            c5 = t < rd.nstale ? stale[rd.stale_off + t] : 0;
            c1 = c5 & 3; c3 = c5 & 63;                         // never read by the k = 1, 3 passes
        } else {
            int b[5];
#pragma unroll
            for (int t = 0; t < 5; t++) b[t] = (i + t < written) ? padded_base(rd, packed, mt, i + t) : 0;
            c1 = b[0];
            c3 = (i < rd.N - 2) ? (b[0] * 16 + b[1] * 4 + b[2]) : b[0];
            c5 = (i < rd.N - 4) ? ((((b[0] * 4 + b[1]) * 4 + b[2]) * 4 + b[3]) * 4 + b[4]) : b[0];
        }
        s1[rd.code_off + i] = (uint8_t)c1;
        s3[rd.code_off + i] = (uint8_t)c3;
        s5[rd.code_off + i] = (uint16_t)c5;
    }
}

// ---------------------------------------------------------------- K2: sliding two-window distance stream
// One thread per chunk; the histograms live in shared memory, bin-major so that a warp touching the same bin is
// conflict-free.  Pearson needs both windows' histograms (sum HA^2, sum HB^2, sum HA*HB); the Manhattan distance
// sum |HA - HB| only needs their difference, so that mode keeps ONE table HA - HB per thread: half the shared-memory
// traffic per position and twice the threads per SM (the k = 5 passes are bound by shared-memory capacity).
template <int K, bool MANHATTAN, int T, typename CodeT>
__global__ void __launch_bounds__(T)
di_slide(const DiTask *__restrict__ tasks, int ntasks, const DiRead *__restrict__ reads,
         const DiPass *__restrict__ passes, const CodeT *__restrict__ codes,
         int *__restrict__ stream_i, double *__restrict__ stream_d)
{
    constexpr int BINS = K == 1 ? 4 : (K == 3 ? 64 : 1024);
    extern __shared__ short sh[];
    short *HA = sh + threadIdx.x;                   // HA[bin * T]   (Manhattan: HA - HB)
    short *HB = sh + (MANHATTAN ? 0 : BINS * T) + threadIdx.x;
    const int tid = blockIdx.x * T + threadIdx.x;
    if (tid >= ntasks) return;
    const DiTask tk = tasks[tid];
    const DiRead rd = reads[tk.read];
    const DiPass ps = passes[rd.pass_begin + tk.pass];
    const CodeT *S = codes + rd.code_off;
    const int w = ps.w;
    for (int b = 0; b < BINS; b++) { HA[b * T] = 0; if (!MANHATTAN) HB[b * T] = 0; }
    // exact integer sums
    int D = 0;                                       // sum |HA - HB|
    long long SA = 0, SB = 0, IP = 0;                // sum HA^2, sum HB^2, sum HA*HB
    auto bumpA = [&](int bin, int delta) {
        if (MANHATTAN) {
            const int d = HA[bin * T];
            D += abs(d + delta) - abs(d);
            HA[bin * T] = (short)(d + delta);
        } else {
            const int a = HA[bin * T], b = HB[bin * T];
            SA += 2 * a * delta + 1; IP += (long long)delta * b;
            HA[bin * T] = (short)(a + delta);
        }
    };
    auto bumpB = [&](int bin, int delta) {
        if (MANHATTAN) {
            const int d = HA[bin * T];
            D += abs(d - delta) - abs(d);
            HA[bin * T] = (short)(d - delta);
        } else {
            const int a = HA[bin * T], b = HB[bin * T];
            SB += 2 * b * delta + 1; IP += (long long)delta * a;
            HB[bin * T] = (short)(b + delta);
        }
    };
    const int q0 = tk.q0;
    for (int t = 0; t < w; t++) { bumpA(S[q0 + t], +1); bumpB(S[q0 + w + t], +1); }
    const double n = (double)BINS, sw = (double)w;
    for (int q = q0; q < q0 + tk.nq; q++) {
        if (MANHATTAN) {
            stream_i[ps.stream_off + q] = D;
        } else {
            // fill_directional_index.c:340-352 (all operands are exact integers below 2^53)
            const double sda = sqrt((double)SA * n - sw * sw);
            const double sdb = sqrt((double)SB * n - sw * sw);
            double P = 0;
            if (sda * sdb > 0) P = ((double)IP * n - sw * sw) / (sda * sdb);
            stream_d[ps.stream_off + q] = P;
        }
        const int a = S[q], b = S[q + w], c = S[q + 2 * w];
        bumpA(a, -1); bumpA(b, +1); bumpB(b, -1); bumpB(c, +1);
    }
}

// ---------------------------------------------------------------- K2b: merge, unshift, prune (one thread per read)
template <bool MANHATTAN>
__device__ __forceinline__ double di_value(const DiPass &ps, const int *__restrict__ si,
                                           const double *__restrict__ sd, int p)
{
    // directional_index_tmp[p]: -1 outside [w, w + steps)
    const int i = p - ps.w;
    if (i < 0 || i >= ps.steps) return -1.0;
    if (MANHATTAN) {
        const int d01 = si[ps.stream_off + i], d12 = si[ps.stream_off + i + ps.w];
        return ((double)d01 - (double)d12) / (2 * (double)ps.w);          // :211
    }
    return sd[ps.stream_off + i + ps.w] - sd[ps.stream_off + i];          // P_12 - P_01, :355
}

// One warp per read.  Per pass the 32 lanes first materialise directional_index_tmp[0..N) as fp64 (two loads and
// one divide per position, in parallel), then the warp runs the reference's sequential state machine over it,
// 32 positions per step.
template <bool MANHATTAN>
__global__ void __launch_bounds__(128)
di_merge(const DiRead *__restrict__ reads, int n_reads, const DiPass *__restrict__ passes,
         const int *__restrict__ si, const double *__restrict__ sd,
         double *__restrict__ wdi, int *__restrict__ wend, int *__restrict__ ww, double *__restrict__ wtmp,
         double *__restrict__ odi, int *__restrict__ oend, int *__restrict__ ow)
{
    const int rdi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (rdi >= n_reads) return;
    const DiRead rd = reads[rdi];
    const int N = rd.N, L = rd.len, r = rd.r;
    double *DI = wdi + rd.work_off, *tmp = wtmp + rd.work_off;
    int *EN = wend + rd.work_off, *WW = ww + rd.work_off;
    for (int i = lane; i < N; i += 32) { DI[i] = -1; EN[i] = -1; WW[i] = -1; }
    for (int pi = 0; pi < rd.npass; pi++) {
        const DiPass ps = passes[rd.pass_begin + pi];
        const int w = ps.w;
        __syncwarp();
        for (int i = lane; i < N; i += 32) tmp[i] = di_value<MANHATTAN>(ps, si, sd, i);
        __syncwarp();
        // put_local_maximum_into_directional_index (:467-503), 32 positions per step.  The sequential state machine
        // is a running (max, first argmax) with a trigger test per position, then a running (min, first argmin) with
        // another trigger: both are inclusive prefix scans with a first-wins combiner, and the first lane whose
        // trigger fires is the event the sequential loop would have reached.
        double local_max = -1;
        int lmi = -1, i = 0;
        while (i < N) {
            const int p = i + lane;
            double v = p < N ? tmp[p] : -2.0;
            int vi = p;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double ov = __shfl_up_sync(FULLMASK, v, o);
                const int oi = __shfl_up_sync(FULLMASK, vi, o);
                if (lane >= o && !(v > ov)) { v = ov; vi = oi; }          // the earlier position wins ties
            }
            const bool beats = local_max < v;
            const double pm = beats ? v : local_max;
            const int pmi = beats ? vi : lmi;
            bool trig = false;
            if (p < N && pmi >= 0 && pmi + w < p && 0 < pm) trig = DI[pmi] < pm;
            const unsigned tb = __ballot_sync(FULLMASK, trig);
            if (!tb) {
                local_max = __shfl_sync(FULLMASK, pm, 31);
                lmi = __shfl_sync(FULLMASK, pmi, 31);
                i += 32;
                continue;
            }
            const int e = __ffs(tb) - 1;
            local_max = __shfl_sync(FULLMASK, pm, e);
            lmi = __shfl_sync(FULLMASK, pmi, e);
            int next_i = i + e + 1;                                       // if no end is found the loop just goes on
            double local_min = 1;
            int lmj = lmi;
            for (int j0 = lmi; j0 < N; j0 += 32) {
                const int q = j0 + lane;
                double u = q < N ? tmp[q] : 2.0;
                int ui = q;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double ou = __shfl_up_sync(FULLMASK, u, o);
                    const int oi = __shfl_up_sync(FULLMASK, ui, o);
                    if (lane >= o && !(u < ou)) { u = ou; ui = oi; }
                }
                const bool lower = local_min > u;
                const double qm = lower ? u : local_min;
                const int qmi = lower ? ui : lmj;
                const unsigned fb = __ballot_sync(FULLMASK, q < N && qmi + w < q);
                if (fb) {
                    const int f = __ffs(fb) - 1;
                    const int end_j = __shfl_sync(FULLMASK, qmi, f) + w;
                    if (lane == 0) { DI[lmi] = local_max; WW[lmi] = w; EN[lmi] = end_j; }
                    next_i = end_j + 1;                                   // i = local_min_j + w, then the for's i++
                    break;
                }
                local_min = __shfl_sync(FULLMASK, qm, 31);
                lmj = __shfl_sync(FULLMASK, qmi, 31);
            }
            __syncwarp();
            local_max = -1;                                               // local_max_i keeps its value (Q2)
            i = next_i;
        }
    }
    __syncwarp();
    // unshift (:587-597) into the output arrays (only [0, L) is ever read again)
    double *O = odi + rd.pos_off;
    int *OE = oend + rd.pos_off, *OW = ow + rd.pos_off;
    for (int i = lane; i < L; i += 32) { O[i] = DI[i + r]; OE[i] = EN[i + r] - r; OW[i] = WW[i + r]; }
}

// remove_redundant_ranges (fill_directional_index.c:505-546), one warp per read.  The outer loop over range
// starts i is sequential (an earlier iteration may have removed i or j); for a fixed i every later range j is
// judged against i's cached values only, so the j loop is data-parallel: the warp tests 32 live ranges at a
// time, applies "remove j" up to the first lane that says "remove i", and stops there like the reference's break.
__global__ void __launch_bounds__(128)
di_prune(const DiRead *__restrict__ reads, int n_reads, double *__restrict__ wdi, int *__restrict__ wend, int *__restrict__ wpos,
         double *__restrict__ odi, int *__restrict__ oend)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    const DiRead rd = reads[warp];
    const int L = rd.len;
    double *O = odi + rd.pos_off;
    int *OE = oend + rd.pos_off;
    double *LD = wdi + rd.work_off;                 // compacted live ranges: DI, end, start
    int *LE = wend + rd.work_off, *LP = wpos + rd.work_off;
    int n = 0;
    for (int base = 0; base < L; base += 32) {
        const int i = base + lane;
        const double v = i < L ? O[i] : -1.0;
        const bool live = 0 < v;
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int at = n + __popc(m & ((1u << lane) - 1u));
            LD[at] = v; LE[at] = OE[i]; LP[at] = i;
        }
        n += __popc(m);
    }
    __syncwarp();
    for (int a = 0; a < n; a++) {
        const double idi = LD[a];
        if (!(0 < idi)) continue;
        const int i = LP[a], ie = LE[a];
        bool i_dead = false;
        for (int b0 = a + 1; b0 < n && !i_dead; b0 += 32) {
            const int b = b0 + lane;
            int outcome = 0;                        // 1: remove j, 2: remove i
            bool in_range = false;
            if (b < n) {
                const int j = LP[b];
                in_range = j <= ie;
                const double jdi = LD[b];
                if (in_range && 0 < jdi) {
                    const int je = LE[b];
                    const double jac = (double)(min(ie, je) - j) / (double)(max(ie, je) - i);
                    if (0.98 < jac) outcome = idi < jdi ? 2 : 1;
                    else if (ie >= je && idi > jdi) outcome = 1;
                }
            }
            const unsigned kill_i = __ballot_sync(0xffffffffu, outcome == 2);
            const unsigned beyond = __ballot_sync(0xffffffffu, b < n && !in_range);
            const int first = kill_i ? __ffs(kill_i) - 1 : 32;
            if (outcome == 1 && lane < first) { LD[b] = -1; O[LP[b]] = -1; OE[LP[b]] = -1; }
            if (kill_i) { i_dead = true; if (lane == 0) { LD[a] = -1; O[i] = -1; OE[i] = -1; } }
            __syncwarp();
            if (beyond) break;                      // starts are sorted: nothing further can be <= ie
        }
    }
}

// ---------------------------------------------------------------- host side
static int wmax_for(int k, int L)       // largest w of the pass list, fill_directional_index.c:559-574
{
    const int cap = k == 1 ? 20 : (k == 3 ? 80 : 10240);
    int best = 0;
    for (int w = 5; w <= cap && w < L / 2; w *= 2) best = w;
    return best;
}

extern "C" int mtr_di_run(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off,
                          const int64_t *pos_off, double *di, int32_t *end, int32_t *w_out)
{
    if (!ctx) return MTR_EINVAL;
    return mtr_di_run_range(ctx, manhattan, stale, stale_off, pos_off, di, end, w_out, 0, ctx->n_reads);
}

// Reads [first, first + count) of the resident batch only; stale_off / pos_off and the output arrays are indexed as
// for the whole batch, so a caller can sweep the batch in slices and start working on a slice while the next one is
// still on the GPU.
// Runs the kernels for reads [first, first + count) and leaves directional_index / _end / _w in device memory
// (di_device_outputs; read r of the range starts at pos_off[first + r] - pos_off[first]).  The resident engine consumes
// them there; mtr_di_run_range copies them out.
int di_compute(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off, int first, int count)
{
    if (!ctx) return MTR_EINVAL;
    if (first < 0 || count < 0 || first + count > ctx->n_reads) { mtr_set_error(ctx, "di_run: read range outside the resident batch"); return MTR_EINVAL; }
    if (count == 0) return MTR_OK;
    if (!pos_off) { mtr_set_error(ctx, "di_run: null argument"); return MTR_EINVAL; }
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->di) ctx->di = new DiState();
    DiState &d = *ctx->di;
    cudaStream_t s = ctx->main_stream;
    if (!d.mt_ready) {
        std::vector<uint8_t> mt;
        mt_base_table(mt);
        MTR_CUDA(ctx, d.d_mt.reserve(mt.size()));
        MTR_CUDA(ctx, cudaMemcpy(d.d_mt.p, mt.data(), mt.size(), cudaMemcpyHostToDevice));
        d.mt_ready = true;
    }
    const int n = count;
    const int64_t pos0 = pos_off[first], stale0 = (stale && stale_off) ? stale_off[first] : 0;
    std::vector<DiRead> reads(n);
    std::vector<DiPass> passes;
    std::vector<DiTask> tasks[3];
    long long code_total = 0, work_total = 0, stream_total = 0, pp = 0;
    int max_M = 0;
    for (int r = 0; r < n; r++) {
        DiRead &rd = reads[r];
        const int L = ctx->len[first + r];
        rd.word_off = ctx->word_off[first + r];
        rd.len = L;
        rd.r = L < 1000 ? 100 : L / 10;                       // handle_one_read.c:194-202
        rd.N = L + 2 * rd.r;
        rd.M = L + rd.r + 2 * std::max(wmax_for(5, L), 80) + 8;
        rd.M = std::max(rd.M, rd.N);
        rd.code_off = code_total; code_total += (rd.M + 15) & ~15;
        rd.work_off = work_total; work_total += rd.N;
        rd.pos_off = pos_off[first + r] - pos0;
        rd.stale_off = (stale && stale_off) ? stale_off[first + r] - stale0 : 0;
        rd.nstale = (stale && stale_off) ? (int)(stale_off[first + r + 1] - stale_off[first + r]) : 0;
        rd.pass_begin = (int)passes.size();
        max_M = std::max(max_M, rd.M);
        for (int k = 1; k <= 5; k += 2) {
            const int cap = k == 1 ? 20 : (k == 3 ? 80 : 10240);
            for (int w = 5; w <= cap && w < L / 2; w *= 2) {
                DiPass ps;
                ps.k = k; ps.w = w;
                ps.steps = L + rd.r - w - k + 1;
                ps.nq = ps.steps + w;
                ps.stream_off = stream_total; stream_total += ps.nq;
                pp += ps.steps;
                const int chunk = std::max(1024, 2 * w);
                const int pidx = (int)passes.size() - rd.pass_begin;
                for (int q0 = 0; q0 < ps.nq; q0 += chunk)
                    tasks[k / 2].push_back(DiTask{r, pidx, q0, std::min(chunk, ps.nq - q0)});
                passes.push_back(ps);
            }
        }
        rd.npass = (int)passes.size() - rd.pass_begin;
    }
    const long long total_pos = pos_off[first + n] - pos0;
    const long long nstale_total = (stale && stale_off) ? stale_off[first + n] - stale0 : 0;
    MTR_CUDA(ctx, d.d_reads.reserve(sizeof(DiRead) * (size_t)n));
    MTR_CUDA(ctx, d.d_passes.reserve(sizeof(DiPass) * std::max<size_t>(passes.size(), 1)));
    MTR_CUDA(ctx, d.d_s1.reserve((size_t)code_total + 64));
    MTR_CUDA(ctx, d.d_s3.reserve((size_t)code_total + 64));
    MTR_CUDA(ctx, d.d_s5.reserve((size_t)code_total * 2 + 64));
    MTR_CUDA(ctx, d.d_stream.reserve((size_t)std::max<long long>(stream_total, 1) * (manhattan ? 4 : 8)));
    MTR_CUDA(ctx, d.d_stale.reserve((size_t)std::max<long long>(nstale_total, 1) * 2));
    MTR_CUDA(ctx, d.d_di.reserve((size_t)std::max<long long>(total_pos, 1) * 8));
    MTR_CUDA(ctx, d.d_end.reserve((size_t)std::max<long long>(total_pos, 1) * 4));
    MTR_CUDA(ctx, d.d_w.reserve((size_t)std::max<long long>(total_pos, 1) * 4));
    MTR_CUDA(ctx, d.d_work_di.reserve((size_t)work_total * 8));
    MTR_CUDA(ctx, d.d_work_tmp.reserve((size_t)work_total * 8));
    MTR_CUDA(ctx, d.d_work_end.reserve((size_t)work_total * 4));
    MTR_CUDA(ctx, d.d_work_w.reserve((size_t)work_total * 4));
    MTR_CUDA(ctx, cudaMemcpyAsync(d.d_reads.p, reads.data(), sizeof(DiRead) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (!passes.empty())
        MTR_CUDA(ctx, cudaMemcpyAsync(d.d_passes.p, passes.data(), sizeof(DiPass) * passes.size(), cudaMemcpyHostToDevice, s));
    if (nstale_total > 0)
        MTR_CUDA(ctx, cudaMemcpyAsync(d.d_stale.p, stale + stale0, (size_t)nstale_total * 2, cudaMemcpyHostToDevice, s));
    for (int t = 0; t < 3; t++) {
        MTR_CUDA(ctx, d.d_tasks[t].reserve(sizeof(DiTask) * std::max<size_t>(tasks[t].size(), 1)));
        if (!tasks[t].empty())
            MTR_CUDA(ctx, cudaMemcpyAsync(d.d_tasks[t].p, tasks[t].data(), sizeof(DiTask) * tasks[t].size(), cudaMemcpyHostToDevice, s));
    }
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    int launches = 0;
    {
        dim3 grid((unsigned)n, (unsigned)std::min((max_M + 255) / 256, 64));
        di_codes<<<grid, 256, 0, s>>>((const DiRead *)d.d_reads.p, n, (const uint32_t *)ctx->d_packed.p, (const uint8_t *)d.d_mt.p,
                                      (const uint16_t *)d.d_stale.p, (uint8_t *)d.d_s1.p, (uint8_t *)d.d_s3.p, (uint16_t *)d.d_s5.p);
        MTR_CUDA(ctx, cudaGetLastError());
        launches++;
    }
    const DiRead *dr = (const DiRead *)d.d_reads.p;
    const DiPass *dp = (const DiPass *)d.d_passes.p;
    int *si = (int *)d.d_stream.p;
    double *sd = (double *)d.d_stream.p;
#define SLIDE(K, TM, TP, CODET, CODES, IDX)                                                                    \
    if (!tasks[IDX].empty()) {                                                                                 \
        const int nt = (int)tasks[IDX].size();                                                                 \
        constexpr int BINS_ = (K == 1 ? 4 : (K == 3 ? 64 : 1024));                                             \
        if (manhattan) {                                                                                       \
            const size_t smem = (size_t)BINS_ * TM * sizeof(short);                                            \
            MTR_CUDA(ctx, cudaFuncSetAttribute(di_slide<K, true, TM, CODET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            di_slide<K, true, TM, CODET><<<(nt + TM - 1) / TM, TM, smem, s>>>((const DiTask *)d.d_tasks[IDX].p, nt, dr, dp, (const CODET *)CODES, si, sd); \
        } else {                                                                                               \
            const size_t smem = (size_t)2 * BINS_ * TP * sizeof(short);                                        \
            MTR_CUDA(ctx, cudaFuncSetAttribute(di_slide<K, false, TP, CODET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            di_slide<K, false, TP, CODET><<<(nt + TP - 1) / TP, TP, smem, s>>>((const DiTask *)d.d_tasks[IDX].p, nt, dr, dp, (const CODET *)CODES, si, sd); \
        }                                                                                                      \
        MTR_CUDA(ctx, cudaGetLastError());                                                                     \
        launches++;                                                                                            \
    }
    SLIDE(1, 128, 128, uint8_t, d.d_s1.p, 0)
    SLIDE(3, 128, 128, uint8_t, d.d_s3.p, 1)
    SLIDE(5, 96, 48, uint16_t, d.d_s5.p, 2)
#undef SLIDE
    if (manhattan)
        di_merge<true><<<(n * 32 + 127) / 128, 128, 0, s>>>(dr, n, dp, si, sd, (double *)d.d_work_di.p, (int *)d.d_work_end.p, (int *)d.d_work_w.p,
                                                           (double *)d.d_work_tmp.p, (double *)d.d_di.p, (int *)d.d_end.p, (int *)d.d_w.p);
    else
        di_merge<false><<<(n * 32 + 127) / 128, 128, 0, s>>>(dr, n, dp, si, sd, (double *)d.d_work_di.p, (int *)d.d_work_end.p, (int *)d.d_work_w.p,
                                                            (double *)d.d_work_tmp.p, (double *)d.d_di.p, (int *)d.d_end.p, (int *)d.d_w.p);
    MTR_CUDA(ctx, cudaGetLastError());
    launches++;
    di_prune<<<(n * 32 + 127) / 128, 128, 0, s>>>(dr, n, (double *)d.d_work_di.p, (int *)d.d_work_end.p, (int *)d.d_work_w.p,
                                                 (double *)d.d_di.p, (int *)d.d_end.p);
    MTR_CUDA(ctx, cudaGetLastError());
    launches++;
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
    MTR_CUDA(ctx, mtr_sync(ctx));
    float ms = 0;
    MTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]));
    ctx->stats.di_ms = ms;
    ctx->stats.di_position_passes = pp;
    ctx->stats.di_bytes_in = (ctx->word_off[first + n] - ctx->word_off[first]) * 4 + nstale_total * 2;
    ctx->stats.di_bytes_out = 0;
    ctx->stats.launches = launches;
    return MTR_OK;
}

void di_device_outputs(mtr_ctx *ctx, double **di, int **end, int **w)
{
    DiState &d = *ctx->di;
    if (di) *di = (double *)d.d_di.p;
    if (end) *end = (int *)d.d_end.p;
    if (w) *w = (int *)d.d_w.p;
}

extern "C" int mtr_di_run_range(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off,
                                const int64_t *pos_off, double *di, int32_t *end, int32_t *w_out, int first, int count)
{
    if (!ctx) return MTR_EINVAL;
    if (count > 0 && (!pos_off || !end || !w_out)) { mtr_set_error(ctx, "di_run: null argument"); return MTR_EINVAL; }
    const int rc = di_compute(ctx, manhattan, stale, stale_off, pos_off, first, count);
    if (rc || count == 0) return rc;
    DiState &d = *ctx->di;
    cudaStream_t s = ctx->main_stream;
    const int64_t pos0 = pos_off[first];
    const long long total_pos = pos_off[first + count] - pos0;
    // (di_compute has drained the stream: a copy waiting for its stream at the head of a copy-engine queue would hold
    // up the copies of every other context on the GPU)
    if (di) MTR_CUDA(ctx, cudaMemcpyAsync(di + pos0, d.d_di.p, (size_t)total_pos * 8, cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, cudaMemcpyAsync(end + pos0, d.d_end.p, (size_t)total_pos * 4, cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, cudaMemcpyAsync(w_out + pos0, d.d_w.p, (size_t)total_pos * 4, cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, mtr_sync(ctx));
    ctx->stats.di_bytes_out = total_pos * (di ? 16 : 8);
    return MTR_OK;
}
