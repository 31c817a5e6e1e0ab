// eng.cu -- the resident engine on the GPU: kernel wrappers around eng_core.h and the host driver that launches waves.
//
// mtr_engine_run = handle_one_TR (/root/reference/handle_one_read.c:190-261) for every read of the resident batch.
// One wave (see eng_core.h) is ~30 kernel launches on one stream (the K3 class kernels fan out over side streams);
// nothing is copied between host and device inside the loop except a 64-byte status snapshot that the last kernel of
// a wave writes into mapped host memory.  The host launches a few waves ahead and looks at the snapshot in between.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "eng_host.h"

using namespace eng;

// ---------------------------------------------------------------- kernels
// MTR_TIMELINE: every kernel of a wave stamps %globaltimer when its first thread starts (and publish when it ends), so the
// waves of concurrent groups can be laid side by side afterwards
__device__ __forceinline__ void stamp(const Ptrs &P, int id)
{
    if (!P.stamps) return;
    const int w = P.ctr->waves;
    if (w >= P.stamp_waves) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.stamps[(size_t)w * 16 + id] = t;
}
#define STAMP0(id) do { if (blockIdx.x == 0 && threadIdx.x == 0) stamp(P, id); } while (0)

__global__ void eng_begin(Ptrs P, DpQueue QL) { if (threadIdx.x == 0 && blockIdx.x == 0) { wave_begin(P, QL); stamp(P, 0); } }

// every chain in WAIT_* whose results have all arrived; 32 chain slots per warp step, the ready ones one after the other
__global__ void __launch_bounds__(128) eng_advance(Ptrs P, int n_chains)
{
    STAMP0(1);
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_chains; base += nw * 32) {
        const int mine = base + lane();
        unsigned m = wballot(mine < n_chains && chain_ready(P, mine));
        while (m) {
            const int c = base + ffs(m) - 1;
            m &= m - 1u;
            advance_chain(P, c);
        }
    }
}

// walk / polish kernels: one block (four warps) per task, persistent blocks pulling from a list; shared memory = 16 ints
// of cta scratch + 32 KB for the count table (DIRECT / COMPACT) or for the probe cache of a WIDE table
template <int WHICH>      // 0: walks (queue entries below the tail seen at the start), 1: polish
__global__ void __launch_bounds__(128) eng_unitfinder(Ptrs P, int slice0)
{
    __shared__ int sh[16];
    extern __shared__ __align__(16) unsigned char uf_dyn[];     // kUfDynSmem bytes: count table | near tie lists | walk memos
    unsigned *tab = (unsigned *)uf_dyn;
    int *near = (int *)(uf_dyn + kUfSmemWords * 4);
    MemoEntry *memos = (MemoEntry *)(uf_dyn + kUfSmemWords * 4 + 4 * kTiesNear * 4);   // zeroed once: epoch 0 is never current
    if (WHICH == 0) { for (int i = threadIdx.x; i < 2 * kMemoSlots; i += blockDim.x) memos[i].epoch = 0u; __syncthreads(); }
    const Cta c = cta_of_block(sh);
    const Scratch S = scratch_of(P, slice0 + blockIdx.x);
    STAMP0(WHICH == 1 ? 2 : 4);
    if (WHICH == 1) {
        const int n = P.ctr->n_polish;
        for (;;) {
            int i = 0;
            if (c.tid == 0) i = atomicAdd(&P.ctr->polish_head, 1);
            i = cta_bcast(c, i, 7);
            if (i >= n) break;
            polish_chain(P, P.polish_list[i], S, c, tab);
        }
        return;
    }
    // every entry below `limit` was written before this instance was launched (later ones belong to later instances)
    const unsigned limit = *(volatile unsigned *)&P.ctr->walk_tail;
    for (;;) {
        int chain = -1;
        if (c.tid == 0) {
            unsigned h = *(volatile unsigned *)&P.ctr->walk_head;
            while ((int)(limit - h) > 0) {
                const unsigned seen = atomicCAS(&P.ctr->walk_head, h, h + 1u);
                if (seen == h) { chain = P.walk_ring[h & P.walk_ring_mask]; break; }
                h = seen;
            }
        }
        chain = cta_bcast(c, chain, 7);
        if (chain < 0) break;
        walk_chain(P, chain, S, c, tab, near, memos);
    }
}

// one warp per read, one warp per block: 16.5 KB of shared memory fit beside the unit-finder ctas of other groups
__global__ void __launch_bounds__(32) eng_sched(Ptrs P)
{
    __shared__ unsigned long long tab[kInlineSlots];
    __shared__ int sh[16];
    const int read = blockIdx.x;
    STAMP0(3);
    if (read >= P.n_reads) return;
    sched_read(P, read, tab, kInlineSlots, sh);
}

__global__ void __launch_bounds__(256) eng_emit(Ptrs P, DpQueue QL, int n_chains)
{
    STAMP0(5);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chains) emit_chain(P, QL, c);
}

__global__ void eng_plan(Ptrs P, DpQueue Q) { if (Q.id == 0) STAMP0(6); if (blockIdx.x == 0 && threadIdx.x < 32) plan_tasks(P, Q); }

__global__ void __launch_bounds__(256) eng_scatter(Ptrs P, DpQueue Q)
{
    if (Q.id == 0) STAMP0(7);
    const int n = min(Q.qc->n_tasks, Q.task_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) scatter_task(Q, i);
}

__global__ void __launch_bounds__(256) eng_zero_aux(Ptrs P, DpQueue Q)
{
    if (Q.id == 0) STAMP0(8);
    const long long n = min((long long)Q.qc->aux_used, Q.aux_cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) Q.aux[i] = 0;
}

struct EngSnapshot { int unfinished, error, error_read, deferred, waves, n_accepted, n_tasks, pad; unsigned long long tasks_total, candidates_started; };

__global__ void eng_publish(Ptrs P, EngSnapshot *snap)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    stamp(P, 10);
    const Counters &c = *P.ctr;
    snap->error = c.error; snap->error_read = c.error_read; snap->deferred = c.deferred; snap->waves = c.waves;
    snap->n_accepted = c.n_accepted; snap->n_tasks = P.q.qc->n_tasks; snap->pad = (int)(c.walk_tail - c.walk_head) + c.walks_running + c.dp_pending; snap->tasks_total = c.tasks_total; snap->candidates_started = c.tables + (unsigned long long)(unsigned)c.progress + ((unsigned long long)(unsigned)c.walks_done << 32);
    __threadfence_system();
    snap->unfinished = c.unfinished;
}

// ---------------------------------------------------------------- per-context engine state
struct EngState {
    DevBuf d_main, d_dirs, d_scratch, d_wide, d_stamps, d_dirs_long[kLongInst];
    cudaStream_t long_stream[kLongInst] = {};  // long DP queues run here, beside the waves
    cudaEvent_t long_done[kLongInst] = {}, long_ev0[kLongInst] = {}, long_ev1[kLongInst] = {}, emit_done = nullptr;
    bool long_busy[kLongInst] = {};
    PinBuf h_snap, h_acc, h_ctr;
    Config cfg;
    Layout lay;
    Ptrs P;
    bool bound = false;
    int speculate = 8;
    cudaStream_t side[8] = {};
    cudaEvent_t fork = nullptr, join[8] = {};
    cudaStream_t walk_stream[4] = {};      // walk kernel instances run here, beside everything else
    cudaEvent_t sched_done[4] = {};
    int n_side = 0;
    static constexpr int kEv = 32;
    cudaEvent_t ev_dp0[kEv] = {}, ev_dp1[kEv] = {}, ev_w0[kEv] = {};
    std::vector<mtr_repeat> reps;
    std::vector<uint8_t> units;
};

void eng_state_free(mtr_ctx *ctx)
{
    if (!ctx->eng) return;
    EngState *e = ctx->eng;
    e->d_main.release(); e->d_dirs.release(); e->d_scratch.release(); e->d_wide.release(); e->d_stamps.release();
    for (int i = 0; i < kLongInst; i++) {
        e->d_dirs_long[i].release();
        if (e->long_stream[i]) cudaStreamDestroy(e->long_stream[i]);
        if (e->long_done[i]) cudaEventDestroy(e->long_done[i]);
        if (e->long_ev0[i]) cudaEventDestroy(e->long_ev0[i]);
        if (e->long_ev1[i]) cudaEventDestroy(e->long_ev1[i]);
    }
    if (e->emit_done) cudaEventDestroy(e->emit_done);
    e->h_snap.release(); e->h_acc.release(); e->h_ctr.release();
    for (int i = 0; i < 8; i++) { if (e->side[i]) cudaStreamDestroy(e->side[i]); if (e->join[i]) cudaEventDestroy(e->join[i]); }
    if (e->fork) cudaEventDestroy(e->fork);
    for (int i = 0; i < 4; i++) { if (e->walk_stream[i]) cudaStreamDestroy(e->walk_stream[i]); if (e->sched_done[i]) cudaEventDestroy(e->sched_done[i]); }
    for (int i = 0; i < EngState::kEv; i++) {
        if (e->ev_dp0[i]) cudaEventDestroy(e->ev_dp0[i]);
        if (e->ev_dp1[i]) cudaEventDestroy(e->ev_dp1[i]);
        if (e->ev_w0[i]) cudaEventDestroy(e->ev_w0[i]);
    }
    delete e;
    ctx->eng = nullptr;
}

// ---------------------------------------------------------------- K3 busy intervals of the whole process (roofline)
namespace {
struct BusyLog {
    std::mutex mu;
    cudaEvent_t base[16] = {};
    std::vector<std::pair<double, double>> iv[16];
} g_busy;

cudaEvent_t busy_base(int device)
{
    std::lock_guard<std::mutex> g(g_busy.mu);
    if (device < 0 || device >= 16) return nullptr;
    if (!g_busy.base[device]) {
        cudaEventCreate(&g_busy.base[device]);
        cudaEventRecord(g_busy.base[device], 0);
        cudaEventSynchronize(g_busy.base[device]);
    }
    return g_busy.base[device];
}
}  // namespace

extern "C" double mtr_engine_dp_busy_ms(int device, int reset)
{
    if (device < 0 || device >= 16) return 0;
    std::lock_guard<std::mutex> g(g_busy.mu);
    std::vector<std::pair<double, double>> &v = g_busy.iv[device];
    std::sort(v.begin(), v.end());
    double total = 0, hi = -1e300;
    for (const auto &p : v) {
        if (p.first > hi) { total += p.second - p.first; hi = p.second; }
        else if (p.second > hi) { total += p.second - hi; hi = p.second; }
    }
    if (reset) v.clear();
    return total;
}

extern "C" int mtr_engine_set_speculate(mtr_ctx *ctx, int depth)
{
    if (!ctx || depth < 0) return MTR_EINVAL;
    if (!ctx->eng) ctx->eng = new EngState();
    ctx->eng->speculate = depth;
    return MTR_OK;
}

static double wall_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- the driver
extern "C" int mtr_engine_run(mtr_ctx *ctx, int manhattan, float min_match_ratio, const uint16_t *stale, const int64_t *stale_off,
                              const mtr_repeat **repeats, int64_t *n_repeats, const uint8_t **units, mtr_engine_stats *stats)
{
    if (!ctx) return MTR_EINVAL;
    return mtr_engine_run_range(ctx, 0, ctx->n_reads, manhattan, min_match_ratio, stale, stale_off, repeats, n_repeats, units, stats);
}

extern "C" int mtr_engine_run_range(mtr_ctx *ctx, int first, int count, int manhattan, float min_match_ratio, const uint16_t *stale,
                                    const int64_t *stale_off, const mtr_repeat **repeats, int64_t *n_repeats, const uint8_t **units,
                                    mtr_engine_stats *stats)
{
    if (!ctx) return MTR_EINVAL;
    if (repeats) *repeats = nullptr;
    if (n_repeats) *n_repeats = 0;
    if (units) *units = nullptr;
    if (stats) memset(stats, 0, sizeof *stats);
    if (first < 0 || count < 0 || first + count > ctx->n_reads) { mtr_set_error(ctx, "engine_run: read range outside the resident batch"); return MTR_EINVAL; }
    const int n = count;
    if (n == 0) return MTR_OK;
    const double t_wall0 = wall_ms();
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->eng) ctx->eng = new EngState();
    EngState &E = *ctx->eng;
    cudaStream_t s = ctx->main_stream;
    static const bool prof = getenv("MTR_PROFILE") != nullptr;
    if (E.n_side == 0) {
        int want = 1;
        if (const char *e = getenv("MTR_ENGINE_SIDE_STREAMS")) want = std::max(0, std::min(1, atoi(e)));
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int i = 0; i < want; i++) {
            MTR_CUDA(ctx, cudaStreamCreateWithFlags(&E.side[i], cudaStreamNonBlocking));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.join[i], cudaEventDisableTiming));
        }
        MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.fork, cudaEventDisableTiming));
        for (int i = 0; i < kLongInst; i++) {
            MTR_CUDA(ctx, cudaStreamCreateWithFlags(&E.long_stream[i], cudaStreamNonBlocking));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.long_done[i], cudaEventDisableTiming));
            MTR_CUDA(ctx, cudaEventCreate(&E.long_ev0[i]));
            MTR_CUDA(ctx, cudaEventCreate(&E.long_ev1[i]));
        }
        MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.emit_done, cudaEventDisableTiming));
        for (int i = 0; i < 4; i++) {
            MTR_CUDA(ctx, cudaStreamCreateWithFlags(&E.walk_stream[i], cudaStreamNonBlocking));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.sched_done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < EngState::kEv; i++) {
            MTR_CUDA(ctx, cudaEventCreate(&E.ev_dp0[i]));
            MTR_CUDA(ctx, cudaEventCreate(&E.ev_dp1[i]));
            MTR_CUDA(ctx, cudaEventCreate(&E.ev_w0[i]));
        }
        E.n_side = want;
        if (want == 0) E.n_side = -1;                        // "initialised, no side streams"
    }
    const int n_side = std::max(E.n_side, 0);

    // ---- directional index, left in device memory
    // (pos_off is indexed like the resident batch; only the entries of the range matter)
    std::vector<int64_t> pos_off((size_t)first + n + 1, 0);
    int max_len = 0;
    for (int r = 0; r < n; r++) { pos_off[first + r + 1] = pos_off[first + r] + ctx->len[first + r]; max_len = std::max(max_len, (int)ctx->len[first + r]); }
    int rc = di_compute(ctx, manhattan, stale, stale_off, pos_off.data(), first, n);
    if (rc) return rc;
    const double di_ms = ctx->stats.di_ms;
    const int di_launches = ctx->stats.launches;
    const int64_t di_h2d = ctx->stats.di_bytes_in - (ctx->word_off[first + n] - ctx->word_off[first]) * 4;

    // ---- buffers
    Config cfg = default_config(n, pos_off[first + n], max_len, ctx->n_sm);
    if (const char *e = getenv("MTR_ENGINE_WALK_STREAMS")) cfg.walk_streams = std::max(1, std::min(4, atoi(e)));
    if (const char *e = getenv("MTR_ENGINE_WALK_CTAS")) cfg.uf_ctas = cfg.polish_ctas = std::max(1, atoi(e));
    if (const char *e = getenv("MTR_ENGINE_DIR_MB")) cfg.dir_cap = std::max(64LL, atoll(e)) << 20;
    Layout lay = make_layout(cfg);
    MTR_CUDA(ctx, E.d_main.reserve(lay.total));
    const size_t n_slices = (size_t)cfg.uf_ctas * (size_t)cfg.walk_streams + (size_t)cfg.polish_ctas;
    MTR_CUDA(ctx, E.d_scratch.reserve((size_t)lay.uf_stride * n_slices));
    MTR_CUDA(ctx, E.d_wide.reserve((size_t)lay.table_cap * 8 * n_slices));
    if ((size_t)cfg.dir_cap > E.d_dirs.cap) MTR_CUDA(ctx, E.d_dirs.reserve_exact((size_t)cfg.dir_cap));
    cfg.dir_cap = (long long)E.d_dirs.cap;
    if (const char *e = getenv("MTR_ENGINE_LONG_ROWS")) cfg.long_rows = std::max(1, atoi(e));
    for (int i = 0; i < kLongInst; i++) {
        if ((size_t)cfg.long_dir_cap > E.d_dirs_long[i].cap) MTR_CUDA(ctx, E.d_dirs_long[i].reserve_exact((size_t)cfg.long_dir_cap));
        E.long_busy[i] = false;
    }
    cfg.long_dir_cap = (long long)E.d_dirs_long[0].cap;
    for (int i = 1; i < kLongInst; i++) cfg.long_dir_cap = std::min<long long>(cfg.long_dir_cap, (long long)E.d_dirs_long[i].cap);
    MTR_CUDA(ctx, E.h_snap.reserve(sizeof(EngSnapshot)));
    MTR_CUDA(ctx, E.h_ctr.reserve(sizeof(Counters)));
    E.cfg = cfg; E.lay = lay;
    Ptrs P = bind(E.d_main.p, lay, cfg);
    P.packed = (const uint32_t *)ctx->d_packed.p;
    int *d_end = nullptr, *d_w = nullptr;
    di_device_outputs(ctx, nullptr, &d_end, &d_w);
    P.end = d_end; P.w = d_w;
    P.uf_scratch = (unsigned char *)E.d_scratch.p;
    P.uf_wide = (unsigned char *)E.d_wide.p;
    P.min_match_ratio = min_match_ratio;
    P.speculate = E.speculate;
    if (const char *e = getenv("MTR_SPECULATE")) P.speculate = std::max(0, atoi(e));
    static const bool timeline = getenv("MTR_TIMELINE") != nullptr;
    if (timeline) {
        P.stamp_waves = 1024;
        MTR_CUDA(ctx, E.d_stamps.reserve((size_t)P.stamp_waves * 16 * 8));
        MTR_CUDA(ctx, cudaMemsetAsync(E.d_stamps.p, 0, (size_t)P.stamp_waves * 16 * 8, s));
        P.stamps = (unsigned long long *)E.d_stamps.p;
    }
    E.P = P;

    // zero: chain stages (ST_FREE), lists, counters, histograms, the scratch epochs; then the per-read state
    MTR_CUDA(ctx, cudaMemsetAsync((char *)E.d_main.p + lay.chains, 0, sizeof(Chain) * (size_t)lay.n_chains, s));
    MTR_CUDA(ctx, cudaMemsetAsync((char *)E.d_main.p + lay.zero_begin, 0, lay.total - lay.zero_begin, s));
    MTR_CUDA(ctx, cudaMemsetAsync(E.d_scratch.p, 0, (size_t)lay.uf_stride * n_slices, s));
    std::vector<Read> reads;
    init_reads(reads, ctx->word_off.data() + first, ctx->len.data() + first, n);
    MTR_CUDA(ctx, cudaMemcpyAsync(P.reads, reads.data(), sizeof(Read) * (size_t)n, cudaMemcpyHostToDevice, s));
    {
        Counters c0;
        memset(&c0, 0, sizeof c0);
        c0.unfinished = n;
        MTR_CUDA(ctx, cudaMemcpyAsync(P.ctr, &c0, sizeof c0, cudaMemcpyHostToDevice, s));
    }
    EngSnapshot *snap = (EngSnapshot *)E.h_snap.p;
    memset(snap, 0, sizeof *snap);
    snap->unfinished = n;

    // launch descriptions of K3 over the wave's queue and over the long queues
    auto describe = [&](const DpQueue &Q, uint8_t *dirs) {
        WdpDevLaunch L;
        memset(&L, 0, sizeof L);
        L.tasks = Q.tasks; L.class_begin = Q.class_begin; L.seg_task = Q.seg_task; L.seg_slot = Q.seg_slot; L.nseg_family = kSegs / 2;
        L.counters = Q.slot_counter; L.packed = P.packed; L.units = P.units; L.dirs = dirs; L.results = P.results; L.aux = Q.aux;
        L.pending0 = (char *)&P.chains[0].pending; L.pending_stride = (int)sizeof(Chain); L.pending_total = &P.ctr->dp_pending;
        L.blocks = ctx->n_sm * 2;
        return L;
    };
    WdpDevLaunch L = describe(P.q, (uint8_t *)E.d_dirs.p);
    for (int i = 0; i < n_side; i++) { L.side[i] = E.side[i]; L.join[i] = E.join[i]; }
    L.fork = E.fork; L.n_side = n_side;
    DpQueue QLs[kLongInst];
    WdpDevLaunch LL[kLongInst];
    for (int i = 0; i < kLongInst; i++) { QLs[i] = bind_queue(E.d_main.p, lay, cfg, 1 + i); LL[i] = describe(QLs[i], (uint8_t *)E.d_dirs_long[i].p); }
    const DpQueue none = no_queue();

    // (the attribute belongs to the function, not to the launch: every context sets the same constant)
    MTR_CUDA(ctx, cudaFuncSetAttribute(eng_unitfinder<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUfDynSmem));
    MTR_CUDA(ctx, cudaFuncSetAttribute(eng_unitfinder<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUfDynSmem));
    const int launches_per_wave = 10 + 2 + 1, launches_per_long = 3 + 2 + 1;
    cudaEvent_t base = busy_base(ctx->device);
    double dp_ms = 0, uf_ms = 0;
    long long launches = di_launches;
    int burst = 4, wave_no = 0, long_next = 0;
    double dp_long_ms = 0;
    if (const char *e = getenv("MTR_ENGINE_BURST")) burst = std::max(1, std::min(EngState::kEv, atoi(e)));
    unsigned long long last_tasks = ~0ull, last_started = ~0ull;
    int last_accepted = -1, last_unfinished = -1, stalled = 0, quiet = 0;
    double t_launch = 0, t_wait = 0;
    for (;;) {
        const double tl0 = wall_ms();
        for (int b = 0; b < burst; b++) {
            MTR_CUDA(ctx, cudaEventRecord(E.ev_w0[b], s));
            // the long queue of this wave: the next one whose previous use has ended (none: long tasks wait a wave)
            int li = -1;
            for (int t = 0; t < kLongInst && li < 0; t++) {
                const int i = (long_next + t) % kLongInst;
                if (E.long_busy[i] && cudaEventQuery(E.long_done[i]) == cudaSuccess) {
                    E.long_busy[i] = false;
                    float t0 = 0, t1 = 0;
                    if (base && cudaEventElapsedTime(&t0, base, E.long_ev0[i]) == cudaSuccess && cudaEventElapsedTime(&t1, base, E.long_ev1[i]) == cudaSuccess) {
                        std::lock_guard<std::mutex> gl(g_busy.mu);
                        g_busy.iv[ctx->device].push_back(std::make_pair((double)t0, (double)t1));
                        dp_long_ms += t1 - t0;
                    }
                }
                if (!E.long_busy[i]) li = i;
            }
            cudaGetLastError();                                 // (cudaErrorNotReady of the queries)
            if (li >= 0) long_next = (li + 1) % kLongInst;
            const DpQueue &QL = li >= 0 ? QLs[li] : none;
            eng_begin<<<1, 32, 0, s>>>(P, QL);
            eng_advance<<<ctx->n_sm * 4, 128, 0, s>>>(P, lay.n_chains);
            eng_unitfinder<1><<<cfg.polish_ctas, 128, kUfDynSmem, s>>>(P, cfg.uf_ctas * cfg.walk_streams);
            eng_sched<<<n, 32, 0, s>>>(P);
            {
                // the walk kernel of this wave: on the next walk stream, behind the scheduler pass, beside everything else
                const int ws = wave_no++ % cfg.walk_streams;
                MTR_CUDA(ctx, cudaEventRecord(E.sched_done[ws], s));
                MTR_CUDA(ctx, cudaStreamWaitEvent(E.walk_stream[ws], E.sched_done[ws], 0));
                eng_unitfinder<0><<<cfg.uf_ctas, 128, kUfDynSmem, E.walk_stream[ws]>>>(P, ws * cfg.uf_ctas);
            }
            eng_emit<<<(lay.n_chains + 255) / 256, 256, 0, s>>>(P, QL, lay.n_chains);
            if (li >= 0) {
                // K3 over the long queue, on its own stream behind this wave's emission
                cudaStream_t ls = E.long_stream[li];
                MTR_CUDA(ctx, cudaEventRecord(E.emit_done, s));
                MTR_CUDA(ctx, cudaStreamWaitEvent(ls, E.emit_done, 0));
                eng_plan<<<1, 32, 0, ls>>>(P, QL);
                eng_scatter<<<ctx->n_sm, 256, 0, ls>>>(P, QL);
                eng_zero_aux<<<ctx->n_sm, 256, 0, ls>>>(P, QL);
                MTR_CUDA(ctx, cudaEventRecord(E.long_ev0[li], ls));
                MTR_CUDA(ctx, wdp_launch_dev(LL[li], ls));
                MTR_CUDA(ctx, cudaEventRecord(E.long_ev1[li], ls));
                MTR_CUDA(ctx, cudaEventRecord(E.long_done[li], ls));
                E.long_busy[li] = true;
                launches += launches_per_long;
            }
            eng_plan<<<1, 32, 0, s>>>(P, P.q);
            eng_scatter<<<ctx->n_sm * 2, 256, 0, s>>>(P, P.q);
            eng_zero_aux<<<ctx->n_sm * 2, 256, 0, s>>>(P, P.q);
            MTR_CUDA(ctx, cudaGetLastError());
            MTR_CUDA(ctx, cudaEventRecord(E.ev_dp0[b], s));
            MTR_CUDA(ctx, wdp_launch_dev(L, s));
            MTR_CUDA(ctx, cudaEventRecord(E.ev_dp1[b], s));
            eng_publish<<<1, 1, 0, s>>>(P, snap);
            MTR_CUDA(ctx, cudaGetLastError());
            launches += launches_per_wave;
        }
        const double tl1 = wall_ms();
        MTR_CUDA(ctx, mtr_sync(ctx));
        t_launch += tl1 - tl0; t_wait += wall_ms() - tl1;
        for (int b = 0; b < burst; b++) {
            float f = 0, g = 0, t0 = 0, t1 = 0;
            MTR_CUDA(ctx, cudaEventElapsedTime(&f, E.ev_dp0[b], E.ev_dp1[b]));
            MTR_CUDA(ctx, cudaEventElapsedTime(&g, E.ev_w0[b], E.ev_dp0[b]));
            dp_ms += f; uf_ms += g;
            if (base && cudaEventElapsedTime(&t0, base, E.ev_dp0[b]) == cudaSuccess && cudaEventElapsedTime(&t1, base, E.ev_dp1[b]) == cudaSuccess) {
                std::lock_guard<std::mutex> gl(g_busy.mu);
                g_busy.iv[ctx->device].push_back(std::make_pair((double)t0, (double)t1));
            }
        }
        if (prof) fprintf(stderr, "[mtr engine] ctx %p wave %d: unfinished %d accepted %d tasks(last wave) %d deferred %d walks queued or running %d\n", (void *)ctx, snap->waves, snap->unfinished, snap->n_accepted, snap->n_tasks, snap->deferred, snap->pad);
        if (snap->error) break;
        if (snap->unfinished <= 0) break;
        // no progress at all during a whole burst: either the per-wave budgets are too small for a single task (grow
        // them) or the engine is stuck (a bug: fail loudly instead of spinning)
        if (snap->pad == 0 && snap->tasks_total == last_tasks && snap->candidates_started == last_started && snap->n_accepted == last_accepted && snap->unfinished == last_unfinished) {
            if (snap->deferred == 0 && ++quiet < 3) {
                // (a walk may have published its result after the last emission pass of the burst: look again)
            } else if (snap->deferred > 0 && stalled < 6) {
                quiet = 0;
                const size_t want = E.d_dirs.cap * 2;
                E.d_dirs.release();
                MTR_CUDA(ctx, E.d_dirs.reserve_exact(want));
                P.q.dir_cap = (long long)E.d_dirs.cap; E.P = P;
                L.dirs = (uint8_t *)E.d_dirs.p;
                stalled++;
            } else {
                mtr_set_error(ctx, "engine_run: no progress after wave %d (%d reads unfinished, %d tasks deferred)", snap->waves, snap->unfinished, snap->deferred);
                return MTR_ECUDA;
            }
        }
        else quiet = 0;
        last_tasks = snap->tasks_total; last_started = snap->candidates_started; last_accepted = snap->n_accepted; last_unfinished = snap->unfinished;
    }
    for (int i = 0; i < cfg.walk_streams; i++) MTR_CUDA(ctx, cudaStreamSynchronize(E.walk_stream[i]));   // walks nobody waits for any more
    for (int i = 0; i < kLongInst; i++) {
        MTR_CUDA(ctx, cudaStreamSynchronize(E.long_stream[i]));
        float t0 = 0, t1 = 0;
        if (E.long_busy[i] && base && cudaEventElapsedTime(&t0, base, E.long_ev0[i]) == cudaSuccess && cudaEventElapsedTime(&t1, base, E.long_ev1[i]) == cudaSuccess) {
            std::lock_guard<std::mutex> gl(g_busy.mu);
            g_busy.iv[ctx->device].push_back(std::make_pair((double)t0, (double)t1));
            dp_long_ms += t1 - t0;
        }
        E.long_busy[i] = false;
    }
    Counters *hc = (Counters *)E.h_ctr.p;
    MTR_CUDA(ctx, cudaMemcpyAsync(hc, P.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, mtr_sync(ctx));
    if (timeline) {
        std::vector<unsigned long long> st((size_t)P.stamp_waves * 16);
        MTR_CUDA(ctx, cudaMemcpy(st.data(), E.d_stamps.p, st.size() * 8, cudaMemcpyDeviceToHost));
        static std::mutex tm;
        std::lock_guard<std::mutex> g(tm);
        for (int w = 1; w < P.stamp_waves && w <= hc->waves; w++) {
            fprintf(stderr, "[timeline] %p %d", (void *)ctx, w);
            for (int k = 0; k < 11; k++) fprintf(stderr, " %llu", st[(size_t)w * 16 + k]);
            fprintf(stderr, "\n");
        }
    }
    if (prof)
        fprintf(stderr, "[mtr engine] ctx %p: host ms: launching %.1f, waiting %.1f, whole call so far %.1f (di %.1f)\n", (void *)ctx, t_launch, t_wait, wall_ms() - t_wall0, di_ms);
    if (prof)
        fprintf(stderr, "[mtr engine] ctx %p: %d reads, %d waves | unit finder: %llu tables (direct %llu compact %llu wide %llu), %llu with walks, %llu walks | Mticks: build %.1f list %.1f walk fwd %.1f bwd %.1f, slowest task %.3f, slowest walk %.3f | walk steps %llu (memo hits %llu, steps reaching level 4 %llu), probe rounds %llu, failed walks %llu | dp %.1f ms uf %.1f ms\n", (void *)ctx, n, hc->waves,
                hc->tables, hc->prof_kind[2], hc->prof_kind[1], hc->prof_kind[0], hc->prof_walk_tasks, hc->walks, hc->prof_build / 1e6, hc->prof_list / 1e6, hc->prof_walk[0] / 1e6, hc->prof_walk[1] / 1e6,
                hc->prof_max_task / 1e6, hc->prof_max_walk / 1e6, hc->prof_steps, hc->prof_memo_hits, hc->prof_deep_steps, hc->prof_probe_rounds, hc->prof_fail_walks, dp_ms, uf_ms);
    if (hc->error) {
        switch (hc->error) {
        case ERR_WRAPCAP: mtr_set_error(ctx, "You need to increse the value of WrapDPsize."); return MTR_ERANGE;
        case ERR_EMPTY_UNIT: mtr_set_error(ctx, "the revised repeat unit is empty (read %d)", hc->error_read); return MTR_ERANGE;
        case ERR_ACCEPTED_FULL: mtr_set_error(ctx, "engine_run: more than %d accepted repeats in one group", cfg.acc_cap); return MTR_ENOMEM;
        default: mtr_set_error(ctx, "engine_run: device error %d", hc->error); return MTR_ECUDA;
        }
    }
    const int na = hc->n_accepted;
    if (na > 0) {
        MTR_CUDA(ctx, E.h_acc.reserve(sizeof(Accepted) * (size_t)na));
        MTR_CUDA(ctx, cudaMemcpyAsync(E.h_acc.p, P.acc, sizeof(Accepted) * (size_t)na, cudaMemcpyDeviceToHost, s));
        MTR_CUDA(ctx, mtr_sync(ctx));
    }
    export_repeats((const Accepted *)E.h_acc.p, na, E.reps, E.units, first);
    if (repeats) *repeats = E.reps.data();
    if (n_repeats) *n_repeats = na;
    if (units) *units = E.units.data();
    if (stats) {
        export_stats(*hc, stats);
        stats->launches = launches;
        stats->di_ms = di_ms; stats->dp_ms = dp_ms + dp_long_ms; stats->uf_ms = uf_ms;
        stats->h2d_bytes = di_h2d + (int64_t)sizeof(Read) * n + (int64_t)sizeof(Counters);
        stats->d2h_bytes = (int64_t)sizeof(Accepted) * na + (int64_t)sizeof(Counters);
        stats->wall_ms = wall_ms() - t_wall0;
    }
    ctx->stats.launches = (int32_t)std::min<long long>(launches, 0x7fffffff);
    return MTR_OK;
}
