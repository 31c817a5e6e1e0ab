// uf.cu -- K4 at kernel level: the unit finder of one (candidate range, k) per warp, as a stand-alone call.
//
// mtr_uf_run = search_De_Bruijn_graph up to the point where it calls wrap_around_DP
// (/root/reference/consensus.c:507-549): init_inputString (:37-60), generate_freqNode_return_list_maxNodes (:132-229)
// and the greedy walks search_De_Bruijn_graph_forward / _backward (:269-505).  The work itself is the resident engine's
// code (eng_core.h: Table, table_list_max, walk with its memo and cycle cut) -- this entry point exists so that exactly
// that device code can be compared task by task with the oracle (tests/test_uf_gpu.py); the pipeline reaches the same
// functions through eng::walk_chain.
#include <algorithm>
#include <cstring>
#include <vector>
#include "eng_core.h"

using namespace eng;

struct UfTaskDev { long long word_off; int L, qs, qe, k, result_idx; };

struct UfState {
    DevBuf d_tasks, d_scratch, d_wide, d_results, d_units, d_scores, d_used, d_head;
};

void uf_state_free(mtr_ctx *ctx)
{
    if (!ctx->uf) return;
    UfState *u = ctx->uf;
    u->d_tasks.release(); u->d_scratch.release(); u->d_results.release(); u->d_units.release(); u->d_scores.release();
    u->d_used.release(); u->d_head.release(); u->d_wide.release();
    delete u;
    ctx->uf = nullptr;
}

// One block per task, exactly as in the engine's walk kernel.
__global__ void __launch_bounds__(128)
uf_kernel(const UfTaskDev *__restrict__ tasks, int ntasks, const uint32_t *__restrict__ packed, Ptrs P,
          mtr_uf_result *__restrict__ results, unsigned char *__restrict__ out_units, int *__restrict__ out_scores,
          unsigned long long *__restrict__ used, int *__restrict__ head)
{
    __shared__ int sh[16];
    extern __shared__ __align__(16) unsigned char uf_dyn[];
    unsigned *tab = (unsigned *)uf_dyn;
    int *near = (int *)(uf_dyn + kUfSmemWords * 4);
    MemoEntry *memos = (MemoEntry *)(uf_dyn + kUfSmemWords * 4 + 4 * kTiesNear * 4);
    for (int i = threadIdx.x; i < 2 * kMemoSlots; i += blockDim.x) memos[i].epoch = 0u;
    __syncthreads();
    const Cta c = cta_of_block(sh);
    const Scratch S = scratch_of(P, blockIdx.x);
    for (;;) {
        int ti = 0;
        if (c.tid == 0) ti = atomicAdd(head, 1);
        ti = cta_bcast(c, ti, 7);
        if (ti >= ntasks) break;
        const UfTaskDev t = tasks[ti];
        const Window win = window_make(packed + t.word_off, t.L, t.k, t.qs, t.qe);
        const Table tb = table_for(S, P, tab, t.qe - t.qs + 1, t.k);
        unit_walks(tb, win, S, c, tab, near, memos, nullptr);
        for (int d = 0; d < 2; d++) {
            if (c.warp != d || !sh[2 + d]) continue;
            const int p = sh[4 + d];
            unsigned long long off = 0;
            if (lane() == 0) off = atomicAdd(used, (unsigned long long)p);
            off = (unsigned long long)bcast((long long)off, 0);
            for (int x = lane(); x < p; x += NL) {
                const int sx = d == 1 ? p - 1 - x : x;
                out_units[off + x] = S.ustr[d][sx];
                out_scores[off + x] = S.uscore[d][sx];
            }
            if (lane() == 0) { sh[9 + d] = (int)(off & 0xffffffffull); sh[11 + d] = (int)(off >> 32); }
        }
        cta_sync(c);
        if (c.tid == 0) {
            mtr_uf_result r;
            r.max_freq = sh[8]; r.found_last = sh[3];
            for (int d = 0; d < 2; d++) {
                r.found[d] = sh[2 + d]; r.period[d] = sh[4 + d];
                r.unit_off[d] = sh[2 + d] ? (long long)(((unsigned long long)(unsigned)sh[11 + d] << 32) | (unsigned)sh[9 + d]) : -1;
            }
            results[t.result_idx] = r;
        }
        cta_sync(c);
    }
}

// Node counts come back clamped to 255 (the engine keeps them in bytes: only == 1 and < 2 are ever tested, consensus.c:600-608).
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
    std::vector<UfTaskDev> dt(n_tasks);
    long long need_out = 0;
    int max_win = 0;
    for (int i = 0; i < n_tasks; i++) {
        const mtr_uf_task &q = tasks[i];
        if (q.read < 0 || q.read >= ctx->n_reads || q.k < 2 || q.k > 15 || q.qs < 0 || q.qe < q.qs || q.qe > ctx->len[q.read]) {
            mtr_set_error(ctx, "uf_run: task %d is malformed (read %d qs %d qe %d k %d)", i, q.read, q.qs, q.qe, q.k);
            return MTR_EINVAL;
        }
        UfTaskDev &t = dt[i];
        t.word_off = ctx->word_off[q.read]; t.L = ctx->len[q.read]; t.qs = q.qs; t.qe = q.qe; t.k = q.k; t.result_idx = i;
        max_win = std::max(max_win, q.qe - q.qs + 1);
        need_out += 2LL * std::min(kMaxPeriod, (q.qe - q.qs) / 5);
    }
    if (need_out > out_cap) { mtr_set_error(ctx, "uf_run: output capacity %lld < %lld", (long long)out_cap, need_out); return MTR_EINVAL; }
    // longest windows first: their walks are the critical path of the launch
    std::sort(dt.begin(), dt.end(), [](const UfTaskDev &a, const UfTaskDev &b) {
        const int na = a.qe - a.qs, nb = b.qe - b.qs;
        return na != nb ? na > nb : a.result_idx < b.result_idx;
    });
    unsigned cap = 64;
    while (cap < 2u * (unsigned)(max_win + 8)) cap <<= 1;
    const long long stride = (kScratchFixed + 255) & ~255LL;
    // MTR_UF_COMPACT_CAP / MTR_UF_DIRECT_K shrink the shared-memory layouts (tests: reach the COMPACT and WIDE paths with
    // small windows)
    unsigned compact_cap = kCompactCap;
    int direct_max_k = 7;
    if (const char *e = getenv("MTR_UF_COMPACT_CAP")) compact_cap = (unsigned)std::max(64, std::min(4096, atoi(e)));
    if (const char *e = getenv("MTR_UF_DIRECT_K")) direct_max_k = std::max(0, std::min(7, atoi(e)));
    const int blocks = (int)std::max<long long>(1, std::min<long long>(std::min(ctx->n_sm * 4, n_tasks), (1LL << 30) / ((long long)cap * 8)));
    cudaStream_t s = ctx->main_stream;
    MTR_CUDA(ctx, u.d_tasks.reserve(sizeof(UfTaskDev) * (size_t)n_tasks));
    MTR_CUDA(ctx, u.d_scratch.reserve((size_t)stride * (size_t)blocks));
    MTR_CUDA(ctx, u.d_wide.reserve((size_t)cap * 8 * (size_t)blocks));
    MTR_CUDA(ctx, u.d_results.reserve(sizeof(mtr_uf_result) * (size_t)n_tasks));
    MTR_CUDA(ctx, u.d_units.reserve((size_t)std::max<long long>(need_out, 16)));
    MTR_CUDA(ctx, u.d_scores.reserve((size_t)std::max<long long>(need_out, 16) * 4));
    MTR_CUDA(ctx, u.d_used.reserve(8));
    MTR_CUDA(ctx, u.d_head.reserve(8));
    MTR_CUDA(ctx, cudaMemcpyAsync(u.d_tasks.p, dt.data(), sizeof(UfTaskDev) * (size_t)n_tasks, cudaMemcpyHostToDevice, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_scratch.p, 0, (size_t)stride * (size_t)blocks, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_used.p, 0, 8, s));
    MTR_CUDA(ctx, cudaMemsetAsync(u.d_head.p, 0, 8, s));
    MTR_CUDA(ctx, cudaEventRecord(ctx->ev[5], s));
    Ptrs P;
    memset(&P, 0, sizeof P);
    P.uf_scratch = (unsigned char *)u.d_scratch.p; P.uf_stride = stride; P.table_cap = cap;
    P.uf_wide = (unsigned char *)u.d_wide.p; P.compact_cap = compact_cap; P.direct_max_k = direct_max_k;
    MTR_CUDA(ctx, cudaFuncSetAttribute(uf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUfDynSmem));
    uf_kernel<<<blocks, 128, kUfDynSmem, s>>>((const UfTaskDev *)u.d_tasks.p, n_tasks, (const uint32_t *)ctx->d_packed.p, P,
                                       (mtr_uf_result *)u.d_results.p, (unsigned char *)u.d_units.p, (int *)u.d_scores.p,
                                       (unsigned long long *)u.d_used.p, (int *)u.d_head.p);
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
    ctx->stats.uf_table_bytes = (long long)kUfSmemWords * 4;
    ctx->stats.launches = 1;
    if (out_used) *out_used = (int64_t)used;
    return MTR_OK;
}
