// uf.cu -- K4: the unit finder of one (candidate range, k) on the GPU, one warp per task.
//
// Replaces search_De_Bruijn_graph up to the point where it calls wrap_around_DP
// (/root/reference/consensus.c:507-549): init_inputString (:37-60), generate_freqNode_return_list_maxNodes
// (:132-229) and the greedy walks search_De_Bruijn_graph_forward / _backward (:269-505).
//
//   codes   k-mer code of every window position straight from the 2-bit packed read (64-bit window, 2-bit
//           groups reversed with __brev); positions >= min(qe, L-k+1) keep the raw base, like the reference (Q7)
//   counts  exact multiset counts in an open-addressing table in HBM (keys by atomicCAS, counts by atomicAdd;
//           the reference's table layout is unobservable); the maximum frequency falls out of the atomicAdd
//           return values
//   list    nodes whose current count equals the maximum, in order of first occurrence, at most 100, each
//           decremented when listed (:157-163): 32 positions per step, duplicates inside a step resolved with
//           __match_any_sync so that the order is exactly the sequential one
//   walks   the greedy cyclic walk is sequential in its steps, but every step's look-ahead evaluates 4*|ties|
//           extensions: lanes probe them in parallel, the maximum is a warp reduction and the tie list is an
//           ordered ballot compaction (first-wins, capped at 1024 like MAX_tiebreaks)
// All table reads use ld.global.cg: the counts are modified by atomics during the task.
#include <algorithm>
#include <cstring>
#include <vector>
#include "mtr_internal.h"

#define UF_EMPTY 0xffffffffu
#define UF_MAXP 500
#define UF_MAXTIES 1024
#define UF_WARPS 4

struct UfTask {
    long long word_off;     // first packed word of the read
    long long table_off;    // first slot of this task's table
    int L, qs, qe, k;
    int cap;                // table slots (power of two)
    int result_idx;
};

struct UfState {
    DevBuf d_tasks, d_keys, d_cnts, d_results, d_units, d_scores, d_used;
    PinBuf h_results;
};

void uf_state_free(mtr_ctx *ctx)
{
    if (!ctx->uf) return;
    UfState *u = ctx->uf;
    u->d_tasks.release(); u->d_keys.release(); u->d_cnts.release(); u->d_results.release();
    u->d_units.release(); u->d_scores.release(); u->d_used.release(); u->h_results.release();
    delete u;
    ctx->uf = nullptr;
}

__device__ __forceinline__ int uf_base(const uint32_t *__restrict__ rd, int i)
{
    return (int)((rd[i >> 4] >> ((i & 15) * 2)) & 3u);
}

// code of the k bases starting at i, first base most significant
__device__ __forceinline__ unsigned uf_kmer(const uint32_t *__restrict__ rd, int i, int k)
{
    const unsigned long long w = (unsigned long long)rd[i >> 4] | ((unsigned long long)rd[(i >> 4) + 1] << 32);
    unsigned v = (unsigned)(w >> ((i & 15) * 2));
    if (k < 16) v &= (1u << (2 * k)) - 1u;
    v = __brev(v);
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);     // bit-reversed -> 2-bit groups reversed
    return v >> (32 - 2 * k);
}

struct UfTable {
    unsigned *keys;
    int *cnts;
    unsigned mask;
    __device__ __forceinline__ unsigned slot0(unsigned code) const { return (code * 2654435761u) & mask; }
    __device__ __forceinline__ int insert(unsigned code)          // returns the count after the insert
    {
        unsigned h = slot0(code);
        for (;;) {
            const unsigned old = atomicCAS(&keys[h], UF_EMPTY, code);
            if (old == UF_EMPTY || old == code) return atomicAdd(&cnts[h], 1) + 1;
            h = (h + 1) & mask;
        }
    }
    __device__ __forceinline__ int find(unsigned code) const       // slot or -1
    {
        unsigned h = slot0(code);
        for (;;) {
            const unsigned key = __ldcg(&keys[h]);
            if (key == code) return (int)h;
            if (key == UF_EMPTY) return -1;
            h = (h + 1) & mask;
        }
    }
    __device__ __forceinline__ int count(unsigned code) const      // freq_node, consensus.c:231-253
    {
        const int h = find(code);
        return h < 0 ? 0 : __ldcg(&cnts[h]);
    }
};

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One greedy walk (consensus.c:269-359 forward, :361-505 backward).  Returns the period (0 = no loop).
// ustr / uscore / ties / fresh live in shared memory and belong to this warp.
__device__ int uf_walk(const UfTable &tb, int qs, int qe, unsigned start, int k, bool backward,
                       unsigned char *ustr, int *uscore, int *ties, int *fresh, int lane)
{
    const unsigned FULL = 0xffffffffu;
    unsigned node = start;
    const int limit = min(UF_MAXP, (qe - qs) / 5);
    int period = 0;
    for (int l = 0; l < limit; l++) {
        if (!backward) {
            const int sc = tb.count(node);
            if (lane == 0) { ustr[l] = (unsigned char)(node >> (2 * (k - 1))); uscore[l] = sc; }
        }
        const int depth = l < 10 ? 1 : k;
        int nties = 1, pick = 0, m;
        if (lane == 0) ties[0] = 0;
        __syncwarp();
        int *cur = ties, *nxt = fresh;
        for (m = 1; m <= depth; m++) {
            const int ncand = 4 * nties;
            const unsigned keep = (k - m) >= 16 ? 0xffffffffu : ((1u << (2 * (k - m))) - 1u);
            // pass 1: maximum count over all extensions
            int best = -1;
            int c_first = -1, d_first = 0;                  // this lane's value in the first chunk (re-used in pass 2)
            for (int ci = lane; ci < ncand; ci += 32) {
                const int t = cur[ci >> 2], b = ci & 3;
                const int digits = backward ? (b << (2 * (m - 1))) + t : 4 * t + b;
                const unsigned cand = backward ? ((unsigned)digits << (2 * (k - m))) + (node >> (2 * m))
                                               : ((node & keep) << (2 * m)) + (unsigned)digits;
                const int c = tb.count(cand);
                if (ci < 32) { c_first = c; d_first = digits; }
                best = max(best, c);
            }
            best = warp_max(best);
            // pass 2: ordered list of the extensions that reach the maximum (first wins, at most 1024)
            int nf = 0;
            bool have_pick = false;
            for (int c0 = 0; c0 < ncand; c0 += 32) {
                const int ci = c0 + lane;
                int c = -2, digits = 0;
                if (ci < ncand) {
                    if (c0 == 0) { c = c_first; digits = d_first; }
                    else {
                        const int t = cur[ci >> 2], b = ci & 3;
                        digits = backward ? (b << (2 * (m - 1))) + t : 4 * t + b;
                        const unsigned cand = backward ? ((unsigned)digits << (2 * (k - m))) + (node >> (2 * m))
                                                       : ((node & keep) << (2 * m)) + (unsigned)digits;
                        c = tb.count(cand);
                    }
                }
                const unsigned eq = __ballot_sync(FULL, c == best);
                if (eq) {
                    if (!have_pick) { pick = __shfl_sync(FULL, digits, __ffs(eq) - 1); have_pick = true; }
                    const int at = nf + __popc(eq & ((1u << lane) - 1u));
                    if (c == best && at < UF_MAXTIES) nxt[at] = digits;
                    nf = min(UF_MAXTIES, nf + __popc(eq));
                }
            }
            __syncwarp();
            if (backward ? nf <= 1 : nf == 1) break;
            int *sw = cur; cur = nxt; nxt = sw;
            nties = nf;
        }
        // m == depth + 1 when the ties were never resolved: the appended base is then pick / 4^depth == 0 ('A', :336)
        if (!backward) {
            node = ((node & ((1u << (2 * (k - 1))) - 1u)) << 2) + ((unsigned)pick >> (2 * (m - 1)));
        } else {
            node = ((unsigned)(pick & 3) << (2 * (k - 1))) + (node >> 2);
            const int sc = tb.count(node);
            if (lane == 0) { ustr[l] = (unsigned char)(node >> (2 * (k - 1))); uscore[l] = sc; }
        }
        if (node == start) { period = l + 1; if (UF_MAXP <= period) period = 0; break; }
    }
    __syncwarp();
    return period;
}

__global__ void __launch_bounds__(UF_WARPS * 32)
uf_kernel(const UfTask *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ packed,
          unsigned *__restrict__ keys, int *__restrict__ cnts, mtr_uf_result *__restrict__ results,
          unsigned char *__restrict__ out_units, int *__restrict__ out_scores, unsigned long long *__restrict__ used)
{
    const unsigned FULL = 0xffffffffu;
    __shared__ int s_ties[UF_WARPS][UF_MAXTIES];
    __shared__ int s_fresh[UF_WARPS][UF_MAXTIES];
    __shared__ int s_score[UF_WARPS][UF_MAXP];
    __shared__ unsigned s_nodes[UF_WARPS][100];
    __shared__ unsigned char s_unit[UF_WARPS][UF_MAXP + 12];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tid = blockIdx.x * UF_WARPS + wib;
    if (tid >= ntasks) return;
    const UfTask t = tasks[tid];
    const uint32_t *rd = packed + t.word_off;
    const int k = t.k, n = t.qe - t.qs + 1;
    const int coded_end = min(t.qe, t.L - k + 1);
    UfTable tb;
    tb.keys = keys + t.table_off; tb.cnts = cnts + t.table_off; tb.mask = (unsigned)t.cap - 1u;
    auto code_at = [&](int p) -> unsigned {                 // init_inputString, consensus.c:37-60
        const int i = t.qs + p;
        if (i < coded_end) return uf_kmer(rd, i, k);
        return i < t.L ? (unsigned)uf_base(rd, i) : 0u;     // raw base; index L is stale in the reference (H4b): 0
    };
    // counts + maximum frequency
    int maxf = -1;
    for (int p = lane; p < n; p += 32) maxf = max(maxf, tb.insert(code_at(p)));
    maxf = warp_max(maxf);
    __syncwarp();
    mtr_uf_result r;
    r.max_freq = maxf; r.found_last = 0;
    r.found[0] = r.found[1] = 0; r.period[0] = r.period[1] = 0; r.unit_off[0] = r.unit_off[1] = -1;
    if (5 < maxf) {
        // list of maximum-frequency nodes (each listed node loses one count)
        int nn = 0;
        for (int p0 = 0; p0 < n && nn < 100; p0 += 32) {
            const int p = p0 + lane;
            unsigned code = 0x80000000u | (unsigned)lane;   // unique dummy for idle lanes
            int h = -1;
            bool ismax = false;
            if (p < n) {
                const unsigned c = code_at(p);
                h = tb.find(c);
                ismax = h >= 0 && __ldcg(&tb.cnts[h]) == maxf;
                if (ismax) code = c;
            }
            const unsigned grp = __match_any_sync(FULL, code);
            const bool lead = ismax && (__ffs(grp) - 1) == lane;
            const unsigned lst = __ballot_sync(FULL, lead);
            const int at = nn + __popc(lst & ((1u << lane) - 1u));
            if (lead && at < 100) { s_nodes[wib][at] = code; atomicSub(&tb.cnts[h], 1); }
            nn = min(100, nn + __popc(lst));
            __syncwarp();
        }
        for (int d = 0; d < 2; d++) {
            for (int i = 0; i < nn; i++) {
                const int period = uf_walk(tb, t.qs, t.qe, s_nodes[wib][i], k, d == 1, s_unit[wib], s_score[wib], s_ties[wib],
                                           s_fresh[wib], lane);
                r.found_last = period > 0;
                if (period == 0) continue;
                unsigned long long off = 0;
                if (lane == 0) off = atomicAdd(used, (unsigned long long)period);
                off = __shfl_sync(FULL, off, 0);
                for (int x = lane; x < period; x += 32) {
                    const int s = d == 1 ? period - 1 - x : x;            // the backward walk is reversed (:458-470)
                    out_units[off + x] = s_unit[wib][s];
                    out_scores[off + x] = s_score[wib][s];
                }
                r.found[d] = 1; r.period[d] = period; r.unit_off[d] = (long long)off;
                __syncwarp();
                break;
            }
        }
    }
    if (lane == 0) results[t.result_idx] = r;
}

// ---------------------------------------------------------------- host side
extern "C" int mtr_uf_run(mtr_ctx *ctx, const mtr_uf_task *tasks, int n_tasks, mtr_uf_result *results,
                          uint8_t *units, int32_t *scores, int64_t out_cap, int64_t *out_used)
{
    if (!ctx) return MTR_EINVAL;
    if (out_used) *out_used = 0;
    if (n_tasks == 0) return MTR_OK;
    if (n_tasks < 0 || !tasks || !results || !units || !scores) { mtr_set_error(ctx, "uf_run: null argument"); return MTR_EINVAL; }
    if (ctx->n_reads == 0) { mtr_set_error(ctx, "uf_run: no resident read batch"); return MTR_EINVAL; }
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->uf) ctx->uf = new UfState();
    UfState &u = *ctx->uf;
    std::vector<UfTask> dt(n_tasks);
    long long slots = 0, need_out = 0;
    for (int i = 0; i < n_tasks; i++) {
        const mtr_uf_task &q = tasks[i];
        if (q.read < 0 || q.read >= ctx->n_reads || q.k < 2 || q.k > 15 || q.qs < 0 || q.qe < q.qs || q.qe > ctx->len[q.read]) {
            mtr_set_error(ctx, "uf_run: task %d is malformed (read %d qs %d qe %d k %d)", i, q.read, q.qs, q.qe, q.k);
            return MTR_EINVAL;
        }
        UfTask &t = dt[i];
        t.word_off = ctx->word_off[q.read]; t.L = ctx->len[q.read]; t.qs = q.qs; t.qe = q.qe; t.k = q.k;
        int cap = 64;
        while (cap < 2 * (q.qe - q.qs + 2)) cap <<= 1;
        t.cap = cap; t.table_off = slots; slots += cap;
        t.result_idx = i;
        need_out += 2LL * std::min(UF_MAXP, (q.qe - q.qs) / 5);
    }
    if (need_out > out_cap) { mtr_set_error(ctx, "uf_run: output capacity %lld < %lld", (long long)out_cap, need_out); return MTR_EINVAL; }
    // longest windows first: their walks are the critical path of the launch
    std::sort(dt.begin(), dt.end(), [](const UfTask &a, const UfTask &b) {
        const int na = a.qe - a.qs, nb = b.qe - b.qs;
        return na != nb ? na > nb : a.result_idx < b.result_idx;
    });
    cudaStream_t s = ctx->main_stream;
    MTR_CUDA(ctx, u.d_tasks.reserve(sizeof(UfTask) * (size_t)n_tasks));
    MTR_CUDA(ctx, u.d_keys.reserve((size_t)slots * 4));
    MTR_CUDA(ctx, u.d_cnts.reserve((size_t)slots * 4));
    MTR_CUDA(ctx, u.d_results.reserve(sizeof(mtr_uf_result) * (size_t)n_tasks));
    MTR_CUDA(ctx, u.d_units.reserve((size_t)std::max<long long>(need_out, 16)));
    MTR_CUDA(ctx, u.d_scores.reserve((size_t)std::max<long long>(need_out, 16) * 4));
    MTR_CUDA(ctx, u.d_used.reserve(8));
    MTR_CUDA(ctx, cudaMemcpyAsync(u.d_tasks.p, dt.data(), sizeof(UfTask) * (size_t)n_tasks, cudaMemcpyHostToDevice, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_keys.p, 0xff, (size_t)slots * 4, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_cnts.p, 0, (size_t)slots * 4, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_used.p, 0, 8, s));
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[5], s));
    uf_kernel<<<(n_tasks + UF_WARPS - 1) / UF_WARPS, UF_WARPS * 32, 0, s>>>(
        (const UfTask *)u.d_tasks.p, n_tasks, (const uint32_t *)ctx->d_packed.p, (unsigned *)u.d_keys.p, (int *)u.d_cnts.p,
        (mtr_uf_result *)u.d_results.p, (unsigned char *)u.d_units.p, (int *)u.d_scores.p, (unsigned long long *)u.d_used.p);
    MTR_CUDA(ctx, cudaGetLastError());
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[6], s));
    unsigned long long used = 0;
    MTR_CUDA(ctx, cudaMemcpyAsync(results, u.d_results.p, sizeof(mtr_uf_result) * (size_t)n_tasks, cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, cudaMemcpyAsync(&used, u.d_used.p, 8, cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, mtr_sync(ctx));
    if ((long long)used > need_out) { mtr_set_error(ctx, "uf_run: device wrote %llu > %lld unit bytes", used, need_out); return MTR_ECUDA; }
    if (used > 0) {
        MTR_CUDA(ctx, cudaMemcpyAsync(units, u.d_units.p, (size_t)used, cudaMemcpyDeviceToHost, s));
        MTR_CUDA(ctx, cudaMemcpyAsync(scores, u.d_scores.p, (size_t)used * 4, cudaMemcpyDeviceToHost, s));
        MTR_CUDA(ctx, mtr_sync(ctx));
    }
    float ms = 0;
    MTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
    ctx->stats.uf_ms = ms;
    ctx->stats.uf_tasks = n_tasks;
    ctx->stats.uf_table_bytes = slots * 8;
    ctx->stats.launches = 1;
    if (out_used) *out_used = (int64_t)used;
    return MTR_OK;
}
