// eng.cu -- the resident engine on the GPU: kernel wrappers around eng_core.h and the host driver that launches waves.
//
// mtr_engine_run = handle_one_TR (/root/reference/handle_one_read.c:190-261) for every read of the resident batch.
// One wave (see eng_core.h) is a dozen small kernels on a high-priority stream; its last kernel writes a status snapshot
// (reads left, tasks emitted into this wave's two DP queues) into mapped host memory.  The host waits for that, launches
// the K3 kernels of the two queues with grids that fit their task counts on the queues' own streams -- no wave ever waits
// for a DP -- and starts the next wave at once.  Nothing else is copied between host and device inside the loop.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
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

__global__ void eng_begin(Ptrs P, DpQueue QS, DpQueue QL) { if (threadIdx.x == 0 && blockIdx.x == 0) { wave_begin(P, QS, QL); stamp(P, 0); } }

// every chain in WAIT_* whose results have all arrived; 32 chain slots per warp step, the ready ones one after the other
__global__ void __launch_bounds__(128) eng_advance(Ptrs P, int n_chains)
{
    STAMP0(1);
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_chains; base += nw * 32) {
        const int mine = base + lane();
        if (mine < n_chains) take_shared(P, mine);
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
__global__ void __launch_bounds__(128) eng_unitfinder(Ptrs P, int slice0, int table_words, int big)
{
    __shared__ int sh[16];
    extern __shared__ __align__(16) unsigned char uf_dyn[];     // count table (table_words words) | near tie lists | walk memos
    unsigned *tab = (unsigned *)uf_dyn;
    int *near = (int *)(uf_dyn + (size_t)table_words * 4);
    MemoEntry *memos = (MemoEntry *)(uf_dyn + (size_t)table_words * 4 + 4 * kTiesNear * 4);   // zeroed once: epoch 0 is never current
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
    unsigned *tail = big ? &P.ctr->walk_tail_big : &P.ctr->walk_tail, *head = big ? &P.ctr->walk_head_big : &P.ctr->walk_head;
    int *ring = big ? P.walk_ring_big : P.walk_ring;
    const unsigned limit = *(volatile unsigned *)tail;
    for (;;) {
        int chain = -1;
        if (c.tid == 0) {
            unsigned h = *(volatile unsigned *)head;
            while ((int)(limit - h) > 0) {
                const unsigned seen = atomicCAS(head, h, h + 1u);
                if (seen == h) {
                    // A pusher takes its place in the queue first (tail) and writes the entry second; this instance may
                    // have started while a later scheduler pass was between the two.  Entries start as -1 and are set
                    // back to -1 when taken, so "not written yet" is visible: wait the few cycles.
                    volatile int *slot = (volatile int *)&ring[h & P.walk_ring_mask];
                    int v;
                    while ((v = *slot) < 0) { }
                    *slot = -1;
                    chain = v;
                    break;
                }
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

__global__ void __launch_bounds__(256) eng_emit(Ptrs P, DpQueue QS, DpQueue QL, int n_chains)
{
    STAMP0(5);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chains) emit_chain(P, QS, QL, c);
}

__global__ void eng_plan(Ptrs P, DpQueue Q) { if (Q.id < kShortInst) STAMP0(6); if (blockIdx.x == 0 && threadIdx.x < 32) plan_tasks(P, Q); }

__global__ void __launch_bounds__(256) eng_scatter(Ptrs P, DpQueue Q)
{
    if (Q.id < kShortInst) STAMP0(7);
    const int n = min(Q.qc->n_tasks, Q.task_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) scatter_task(Q, i);
}

__global__ void __launch_bounds__(256) eng_zero_aux(Ptrs P, DpQueue Q)
{
    if (Q.id < kShortInst) STAMP0(8);
    const long long n = min((long long)Q.qc->aux_used, Q.aux_cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) Q.aux[i] = 0;
}

struct EngSnapshot {
    int unfinished, error, error_read, deferred, waves, n_accepted, in_flight, pad;
    int q_tasks[2], q_slots[2][2];         // this wave's short / long queue: tasks, warp slots of the int32 / int16x2 family
    unsigned long long tasks_total, progress_sig;
};

__global__ void eng_publish(Ptrs P, DpQueue QS, DpQueue QL, EngSnapshot *snap)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    stamp(P, 10);
    const Counters &c = *P.ctr;
    snap->error = c.error; snap->error_read = c.error_read; snap->deferred = c.deferred; snap->waves = c.waves;
    snap->n_accepted = c.n_accepted;
    snap->in_flight = (int)(c.walk_tail - c.walk_head) + (int)(c.walk_tail_big - c.walk_head_big) + c.walks_running + c.dp_pending;
    for (int i = 0; i < 2; i++) {
        const DpQueue &Q = i == 0 ? QS : QL;
        const bool on = Q.tasks_in != nullptr;
        snap->q_tasks[i] = on ? Q.class_begin[WDP_NCLASS] : 0;
        snap->q_slots[i][0] = on ? Q.seg_slot[kSegs / 2] - Q.seg_slot[0] : 0;
        snap->q_slots[i][1] = on ? Q.seg_slot[kSegs] - Q.seg_slot[kSegs / 2] : 0;
    }
    snap->tasks_total = c.tasks_total;
    snap->progress_sig = c.tables + (unsigned long long)(unsigned)c.progress + ((unsigned long long)(unsigned)c.walks_done << 32) + ((unsigned long long)(unsigned)c.next_read << 16);
    __threadfence_system();
    snap->unfinished = c.unfinished;
}

// expected work of a read: the summed widths of its candidate ranges (one warp per read)
__global__ void __launch_bounds__(128) eng_read_weight(const int *__restrict__ end, const long long *__restrict__ pos_off, const int *__restrict__ len, int n, long long *out)
{
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const int L = len[r];
    const int *e = end + pos_off[r];
    long long sum = 0;
    for (int i = lane(); i < L; i += 32) { const int q = e[i]; if (q > -1 && q < L) sum += q - i + 1; }
    sum = wsum(sum);
    if (lane() == 0) out[r] = sum;
}

__global__ void __launch_bounds__(256) eng_unshare(Ptrs P, int n_chains)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chains) unshare_chain(P, c);
}

// diagnostic for the "no progress" failure: the chains that are neither free nor done, and the candidates of their reads
__global__ void eng_dump_stuck(Ptrs P, int n_chains)
{
    int shown = 0;
    for (int c = 0; c < n_chains && shown < 48; c++) {
        const Chain &ch = P.chains[c];
        if (ch.stage == ST_FREE || ch.stage == ST_DONE) continue;
        const Read &rs = P.reads[ch.read];
        if (rs.phase != 0) continue;
        printf("[stuck] chain %d (set %d k-slot %d) read slot %d id %d: stage %d k %d pass %d qs %d qe %d found %d %d period %d %d pending %d wait_lead %d lead %d %d search_done %d ring %d | read: head %d n_ring %d set_mask %x zombie %x\n",
               c, (c / kMaxK) % kSets, c % kMaxK, ch.read, rs.id, ch.stage, ch.k, ch.pass, ch.qs, ch.qe, ch.dir_found[0], ch.dir_found[1], ch.dir_period[0], ch.dir_period[1],
               ch.pending, ch.wait_lead, ch.lead[0], ch.lead[1], ch.search_done, ch.ring, rs.head, rs.n_ring, rs.set_mask, rs.zombie_mask);
        shown++;
    }
}

// ---------------------------------------------------------------- per-context engine state
struct EngState {
    DevBuf d_main, d_scratch, d_wide, d_stamps, d_dirs_q[kQueues], d_end, d_w;
    cudaStream_t tick = nullptr;               // the waves run here (highest priority)
    cudaEvent_t tick_done = nullptr;
    cudaStream_t q_stream[kQueues] = {}, q_side[kQueues] = {};   // K3 of a DP queue runs here, beside the waves
    cudaEvent_t q_done[kQueues] = {}, q_ev0[kQueues] = {}, q_ev1[kQueues] = {}, q_fork[kQueues] = {}, q_join[kQueues] = {};
    bool q_busy[kQueues] = {};
    PinBuf h_snap, h_acc, h_ctr;
    Config cfg;
    Layout lay;
    Ptrs P;
    int speculate = 8;
    cudaStream_t walk_stream[4] = {}, walk_stream_big[4] = {};      // walk kernel instances run here, beside everything else
    cudaEvent_t sched_done[4] = {};
    bool ready = false;
    std::vector<mtr_repeat> reps;
    std::vector<uint8_t> units;
};

void eng_state_free(mtr_ctx *ctx)
{
    if (!ctx->eng) return;
    EngState *e = ctx->eng;
    e->d_end.release(); e->d_w.release(); e->d_main.release(); e->d_scratch.release(); e->d_wide.release(); e->d_stamps.release();
    for (int i = 0; i < kQueues; i++) {
        e->d_dirs_q[i].release();
        if (e->q_stream[i]) cudaStreamDestroy(e->q_stream[i]);
        if (e->q_side[i]) cudaStreamDestroy(e->q_side[i]);
        for (cudaEvent_t ev : {e->q_done[i], e->q_ev0[i], e->q_ev1[i], e->q_fork[i], e->q_join[i]}) if (ev) cudaEventDestroy(ev);
    }
    if (e->tick) cudaStreamDestroy(e->tick);
    if (e->tick_done) cudaEventDestroy(e->tick_done);
    e->h_snap.release(); e->h_acc.release(); e->h_ctr.release();
    for (int i = 0; i < 4; i++) { if (e->walk_stream[i]) cudaStreamDestroy(e->walk_stream[i]); if (e->walk_stream_big[i]) cudaStreamDestroy(e->walk_stream_big[i]); if (e->sched_done[i]) cudaEventDestroy(e->sched_done[i]); }
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
    static const bool prof = getenv("MTR_PROFILE") != nullptr;
    if (!E.ready) {
        // priorities: waves > walks > DP queues (pending blocks of a higher-priority stream are placed first)
        int least = 0, greatest = 0;
        MTR_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const int mid = greatest < least ? least - 1 : least;
        MTR_CUDA(ctx, cudaStreamCreateWithPriority(&E.tick, cudaStreamNonBlocking, greatest));
        MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.tick_done, cudaEventDisableTiming));
        for (int i = 0; i < kQueues; i++) {
            MTR_CUDA(ctx, cudaStreamCreateWithPriority(&E.q_stream[i], cudaStreamNonBlocking, least));
            MTR_CUDA(ctx, cudaStreamCreateWithPriority(&E.q_side[i], cudaStreamNonBlocking, least));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.q_done[i], cudaEventDisableTiming));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.q_fork[i], cudaEventDisableTiming));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.q_join[i], cudaEventDisableTiming));
            MTR_CUDA(ctx, cudaEventCreate(&E.q_ev0[i]));
            MTR_CUDA(ctx, cudaEventCreate(&E.q_ev1[i]));
        }
        for (int i = 0; i < 4; i++) {
            MTR_CUDA(ctx, cudaStreamCreateWithPriority(&E.walk_stream[i], cudaStreamNonBlocking, mid));
            MTR_CUDA(ctx, cudaStreamCreateWithPriority(&E.walk_stream_big[i], cudaStreamNonBlocking, mid));
            MTR_CUDA(ctx, cudaEventCreateWithFlags(&E.sched_done[i], cudaEventDisableTiming));
        }
        E.ready = true;
    }
    cudaStream_t s = ctx->main_stream;                       // directional index and set-up
    cudaStream_t t = E.tick;

    // ---- directional index, computed in slices (the two-window streams of a slice take ~1.5 MB per read) and gathered
    // into the group's own END / W arrays in device memory
    std::vector<int64_t> pos_off((size_t)first + n + 1, 0);
    int max_len = 0;
    for (int r = 0; r < n; r++) { pos_off[first + r + 1] = pos_off[first + r] + ctx->len[first + r]; max_len = std::max(max_len, (int)ctx->len[first + r]); }
    const long long total_pos = pos_off[first + n];
    MTR_CUDA(ctx, E.d_end.reserve((size_t)(total_pos + 1) * 4));
    MTR_CUDA(ctx, E.d_w.reserve((size_t)(total_pos + 1) * 4));
    double di_ms = 0;
    int di_launches = 0;
    int64_t di_h2d = 0;
    {
        int slice_reads = 2048;
        if (const char *e = getenv("MTR_DI_SLICE_READS")) slice_reads = std::max(1, atoi(e));
        for (int a = 0; a < n;) {
            int b = a;
            long long bases = 0;
            while (b < n && b - a < slice_reads && bases < (32LL << 20)) bases += ctx->len[first + b++];
            int rc = di_compute(ctx, manhattan, stale, stale_off, pos_off.data(), first + a, b - a);
            if (rc) return rc;
            int *se = nullptr, *sw = nullptr;
            di_device_outputs(ctx, nullptr, &se, &sw);
            const size_t cnt = (size_t)(pos_off[first + b] - pos_off[first + a]) * 4;
            if (cnt) {
                MTR_CUDA(ctx, cudaMemcpyAsync((int *)E.d_end.p + pos_off[first + a], se, cnt, cudaMemcpyDeviceToDevice, s));
                MTR_CUDA(ctx, cudaMemcpyAsync((int *)E.d_w.p + pos_off[first + a], sw, cnt, cudaMemcpyDeviceToDevice, s));
                MTR_CUDA(ctx, mtr_sync(ctx));
            }
            di_ms += ctx->stats.di_ms;
            di_launches += ctx->stats.launches + 2;
            di_h2d += ctx->stats.di_bytes_in - (ctx->word_off[first + b] - ctx->word_off[first + a]) * 4;
            a = b;
        }
    }

    // ---- buffers
    // read slots: a slot takes the next read of the group as soon as its read has finished (eng_core.h, sched_read)
    int n_slots = 16384;                                     // [B200] 8192: 5190-5340 reads/s, 16384 and 20480: 5450 (the waves scan every slot: more only costs)
    if (const char *e = getenv("MTR_ENGINE_SLOTS")) n_slots = std::max(1, atoi(e));
    Config cfg = default_config(n_slots, n, pos_off[first + n], max_len, ctx->n_sm);
    n_slots = cfg.n_reads;
    if (const char *e = getenv("MTR_ENGINE_WALK_STREAMS")) cfg.walk_streams = std::max(1, std::min(4, atoi(e)));
    if (const char *e = getenv("MTR_ENGINE_WALK_CTAS")) cfg.uf_ctas = std::max(1, atoi(e));
    if (const char *e = getenv("MTR_ENGINE_WALK_CTAS_BIG")) cfg.uf_ctas_big = std::max(1, atoi(e));
    if (const char *e = getenv("MTR_ENGINE_POLISH_CTAS")) cfg.polish_ctas = std::max(1, atoi(e));
    // threads of a walk cta: every warp builds the count table, warps 0 and 1 walk (forward / backward); the registers of
    // the idle warps are held as long as the walks last
    int walk_threads = 128, walk_threads_big = 128;
    if (const char *e = getenv("MTR_ENGINE_WALK_THREADS")) walk_threads = atoi(e) <= 32 ? 32 : (atoi(e) <= 64 ? 64 : 128);
    if (const char *e = getenv("MTR_ENGINE_WALK_THREADS_BIG")) walk_threads_big = atoi(e) <= 32 ? 32 : (atoi(e) <= 64 ? 64 : 128);
    if (const char *e = getenv("MTR_ENGINE_DIR_MB")) cfg.dir_cap = std::max(64LL, atoll(e)) << 20;
    if (const char *e = getenv("MTR_ENGINE_LONG_DIR_MB")) cfg.long_dir_cap = std::max(64LL, atoll(e)) << 20;
    if (const char *e = getenv("MTR_ENGINE_LONG_ROWS")) cfg.long_rows = std::max(1, atoi(e));
    if (const char *e = getenv("MTR_ENGINE_COMPACT_CAP")) cfg.compact_cap = (unsigned)std::max(64, std::min((int)kCompactCap, atoi(e)));
    int n_short = kShortInst, n_long = kLongInst;            // queues in use (fewer: less memory, more deferred emissions)
    if (const char *e = getenv("MTR_ENGINE_SHORT_QUEUES")) n_short = std::max(1, std::min(kShortInst, atoi(e)));
    if (const char *e = getenv("MTR_ENGINE_LONG_QUEUES")) n_long = std::max(1, std::min(kLongInst, atoi(e)));
    Layout lay = make_layout(cfg);
    MTR_CUDA(ctx, E.d_main.reserve(lay.total));
    const size_t n_slices = (size_t)(cfg.uf_ctas + cfg.uf_ctas_big) * (size_t)cfg.walk_streams + (size_t)cfg.polish_ctas;
    MTR_CUDA(ctx, E.d_scratch.reserve((size_t)lay.uf_stride * n_slices));
    MTR_CUDA(ctx, E.d_wide.reserve((size_t)lay.table_cap * 8 * n_slices));
    {
        // direction arenas: the caps are what a full group would like; a context takes what the device still has (several
        // contexts share a GPU and allocate one after the other: a smaller arena only means more deferred emissions, the
        // queue's own cap below follows the size it got).  One task of the longest read must fit, and kArenaHeadroom stays
        // free for the buffers that grow later (accepted repeats, alignments).
        static std::mutex arena_mu;
        std::lock_guard<std::mutex> ga(arena_mu);
        constexpr size_t kArenaHeadroom = 4ull << 30;
        const size_t floor_b = std::max<size_t>(16u << 20, (size_t)1024 * (size_t)std::max(max_len, 1));
        for (int i = 0; i < kQueues; i++) {
            const bool used = i < kShortInst ? i < n_short : i - kShortInst < n_long;
            size_t want = (size_t)(i < kShortInst ? cfg.dir_cap : cfg.long_dir_cap);
            E.q_busy[i] = false;
            if (!used || want <= E.d_dirs_q[i].cap) continue;
            size_t free_b = 0, total_b = 0;
            MTR_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
            const size_t room = free_b + E.d_dirs_q[i].cap > kArenaHeadroom ? free_b + E.d_dirs_q[i].cap - kArenaHeadroom : 0;
            if (want > room) want = std::max(floor_b, room & ~(size_t)0xfffff);
            if (want <= E.d_dirs_q[i].cap) continue;
            cudaError_t e = E.d_dirs_q[i].reserve_exact(want);
            while (e == cudaErrorMemoryAllocation && want / 2 >= floor_b) {     // (fragmentation, a racing allocation of another process)
                cudaGetLastError();
                want /= 2;
                e = E.d_dirs_q[i].reserve_exact(want);
            }
            MTR_CUDA(ctx, e);
        }
    }
    MTR_CUDA(ctx, E.h_snap.reserve(sizeof(EngSnapshot)));
    MTR_CUDA(ctx, E.h_ctr.reserve(sizeof(Counters)));
    E.cfg = cfg; E.lay = lay;
    Ptrs P = bind(E.d_main.p, lay, cfg);
    P.packed = (const uint32_t *)ctx->d_packed.p;
    P.end = (int *)E.d_end.p; P.w = (int *)E.d_w.p;
    P.uf_scratch = (unsigned char *)E.d_scratch.p;
    P.uf_wide = (unsigned char *)E.d_wide.p;
    P.min_match_ratio = min_match_ratio;
    P.speculate = E.speculate;
    if (const char *e = getenv("MTR_SPECULATE")) P.speculate = std::max(0, atoi(e));
    static const bool timeline = getenv("MTR_TIMELINE") != nullptr;
    if (timeline) {
        P.stamp_waves = 4096;
        MTR_CUDA(ctx, E.d_stamps.reserve((size_t)P.stamp_waves * 16 * 8));
        MTR_CUDA(ctx, cudaMemsetAsync(E.d_stamps.p, 0, (size_t)P.stamp_waves * 16 * 8, s));
        P.stamps = (unsigned long long *)E.d_stamps.p;
    }
    E.P = P;
    Ptrs Psmall = P;                                         // the small-window walk kernel: 32 KB count tables
    Psmall.compact_cap = std::min(P.compact_cap, kCompactCapSmall);

    // zero: chain stages (ST_FREE), lists, counters, histograms, the scratch epochs; then the slots and the reads
    MTR_CUDA(ctx, cudaMemsetAsync((char *)E.d_main.p + lay.chains, 0, sizeof(Chain) * (size_t)lay.n_chains, s));
    MTR_CUDA(ctx, cudaMemsetAsync((char *)E.d_main.p + lay.zero_begin, 0, lay.total - lay.zero_begin, s));
    MTR_CUDA(ctx, cudaMemsetAsync(E.d_scratch.p, 0, (size_t)lay.uf_stride * n_slices, s));
    MTR_CUDA(ctx, cudaMemsetAsync(P.walk_ring, 0xff, 4 * (size_t)cfg.walk_cap, s));            // walk queues: every entry "not written" (-1)
    MTR_CUDA(ctx, cudaMemsetAsync(P.walk_ring_big, 0xff, 4 * (size_t)cfg.walk_cap, s));
    std::vector<Read> slots;
    std::vector<ReadDesc> descs;
    init_slots(slots, n_slots);
    // the slots take the reads heaviest first (summed widths of the candidate ranges): the reads with the longest chains
    // of dependent DPs start early instead of making the group's tail
    std::vector<int> order((size_t)n);
    for (int r = 0; r < n; r++) order[r] = r;
    if (n > n_slots / 2 && !getenv("MTR_ENGINE_INPUT_ORDER")) {
        DevBuf d_w8, d_pos, d_len;
        std::vector<long long> w8((size_t)n), pos_rel((size_t)n);
        for (int r = 0; r < n; r++) pos_rel[r] = pos_off[first + r];
        MTR_CUDA(ctx, d_w8.reserve(8 * (size_t)n)); MTR_CUDA(ctx, d_pos.reserve(8 * (size_t)n)); MTR_CUDA(ctx, d_len.reserve(4 * (size_t)n));
        MTR_CUDA(ctx, cudaMemcpyAsync(d_pos.p, pos_rel.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, s));
        MTR_CUDA(ctx, cudaMemcpyAsync(d_len.p, ctx->len.data() + first, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
        eng_read_weight<<<(n + 3) / 4, 128, 0, s>>>((const int *)E.d_end.p, (const long long *)d_pos.p, (const int *)d_len.p, n, (long long *)d_w8.p);
        MTR_CUDA(ctx, cudaMemcpyAsync(w8.data(), d_w8.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
        MTR_CUDA(ctx, mtr_sync(ctx));
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w8[a] > w8[b]; });
        d_w8.release(); d_pos.release(); d_len.release();
    }
    init_descs(descs, ctx->word_off.data() + first, ctx->len.data() + first, n, order.data());
    MTR_CUDA(ctx, cudaMemcpyAsync(P.reads, slots.data(), sizeof(Read) * (size_t)n_slots, cudaMemcpyHostToDevice, s));
    MTR_CUDA(ctx, cudaMemcpyAsync((void *)P.descs, descs.data(), sizeof(ReadDesc) * (size_t)n, cudaMemcpyHostToDevice, s));
    {
        Counters c0;
        memset(&c0, 0, sizeof c0);
        c0.unfinished = n;
        MTR_CUDA(ctx, cudaMemcpyAsync(P.ctr, &c0, sizeof c0, cudaMemcpyHostToDevice, s));
    }
    MTR_CUDA(ctx, mtr_sync(ctx));
    volatile EngSnapshot *snap = (volatile EngSnapshot *)E.h_snap.p;
    memset((void *)snap, 0, sizeof(EngSnapshot));
    snap->unfinished = n;

    // launch descriptions of K3 over the queues
    DpQueue Qs[kQueues];
    WdpDevLaunch LL[kQueues];
    int fill_smem = 0;                                       // > 0: bounds the fill blocks per SM (room for the waves' kernels)
    if (const char *e = getenv("MTR_ENGINE_FILL_SMEM_KB")) fill_smem = std::max(0, std::min(48, atoi(e))) << 10;
    int grid_cap = ctx->n_sm * 8;
    if (const char *e = getenv("MTR_ENGINE_GRID_CAP")) grid_cap = std::max(1, atoi(e));
    for (int i = 0; i < kQueues; i++) {
        Qs[i] = bind_queue(E.d_main.p, lay, cfg, i);
        Qs[i].dir_cap = std::min<long long>(Qs[i].dir_cap, (long long)E.d_dirs_q[i].cap);
        WdpDevLaunch &L = LL[i];
        memset(&L, 0, sizeof L);
        const DpQueue &Q = Qs[i];
        L.tasks = Q.tasks; L.class_begin = Q.class_begin; L.seg_task = Q.seg_task; L.seg_slot = Q.seg_slot; L.nseg_family = kSegs / 2;
        L.counters = Q.slot_counter; L.class_share = &P.ctr->class_share[0]; L.packed = P.packed; L.units = P.units; L.dirs = (uint8_t *)E.d_dirs_q[i].p; L.results = P.results; L.aux = Q.aux;
        L.pending0 = (char *)&P.chains[0].pending; L.pending_stride = (int)sizeof(Chain); L.pending_total = &P.ctr->dp_pending;
        L.blocks = ctx->n_sm * 2; L.fused = 1; L.fill_smem = fill_smem; L.prof = prof ? &P.ctr->prof_k3[0] : nullptr;
        L.side[0] = E.q_side[i]; L.fork = E.q_fork[i]; L.join[0] = E.q_join[i]; L.n_side = 1;
    }
    const DpQueue none = no_queue();

    // (the attribute belongs to the function, not to the launch: every context sets the same constant)
    MTR_CUDA(ctx, cudaFuncSetAttribute(eng_unitfinder<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUfDynSmem));
    MTR_CUDA(ctx, cudaFuncSetAttribute(eng_unitfinder<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUfDynSmem));
    cudaEvent_t base = busy_base(ctx->device);
    double dp_ms = 0;
    long long launches = di_launches, k3_uses = 0, idle_ticks = 0, rescues = 0;
    int wave_no = 0, next_short = 0, next_long = 0, quiet = 0;
    unsigned long long last_tasks = ~0ull, last_sig = ~0ull;
    int last_accepted = -1, last_unfinished = -1;
    double t_launch = 0, t_wait = 0, t_k3 = 0;
    // a queue whose K3 kernels have ended is free again: its results are consumed by the advance pass of the next wave,
    // which runs before that wave's emission refills the queue
    auto reap = [&](int i, bool wait) -> cudaError_t {
        if (!E.q_busy[i]) return cudaSuccess;
        cudaError_t e = wait ? cudaEventSynchronize(E.q_done[i]) : cudaEventQuery(E.q_done[i]);
        if (e == cudaErrorNotReady) { cudaGetLastError(); return cudaSuccess; }
        if (e != cudaSuccess) return e;
        E.q_busy[i] = false;
        float t0 = 0, t1 = 0;
        if (base && cudaEventElapsedTime(&t0, base, E.q_ev0[i]) == cudaSuccess && cudaEventElapsedTime(&t1, base, E.q_ev1[i]) == cudaSuccess) {
            std::lock_guard<std::mutex> gl(g_busy.mu);
            g_busy.iv[ctx->device].push_back(std::make_pair((double)t0, (double)t1));
            dp_ms += t1 - t0;
        }
        return cudaSuccess;
    };
    for (;;) {
        const double tl0 = wall_ms();
        int qi[2] = {-1, -1};
        for (int i = 0; i < kQueues; i++) MTR_CUDA(ctx, reap(i, false));
        for (int k = 0; k < n_short && qi[0] < 0; k++) { const int i = (next_short + k) % n_short; if (!E.q_busy[i]) qi[0] = i; }
        for (int k = 0; k < n_long && qi[1] < 0; k++) { const int i = kShortInst + (next_long + k) % n_long; if (!E.q_busy[i]) qi[1] = i; }
        if (qi[0] >= 0) next_short = (qi[0] + 1) % n_short;
        if (qi[1] >= 0) next_long = (qi[1] - kShortInst + 1) % n_long;
        const DpQueue &QS = qi[0] >= 0 ? Qs[qi[0]] : none, &QL = qi[1] >= 0 ? Qs[qi[1]] : none;
        eng_begin<<<1, 32, 0, t>>>(P, QS, QL);
        eng_advance<<<ctx->n_sm * 4, 128, 0, t>>>(P, lay.n_chains);
        eng_unitfinder<1><<<cfg.polish_ctas, 128, kUfDynSmem, t>>>(P, (cfg.uf_ctas + cfg.uf_ctas_big) * cfg.walk_streams, kUfSmemWords, 0);
        eng_sched<<<n_slots, 32, 0, t>>>(P);
        {
            // the walk kernels of this wave (small and big windows): on the next walk stream pair, behind the scheduler pass,
            // beside everything else
            const int ws = wave_no++ % cfg.walk_streams;
            MTR_CUDA(ctx, cudaEventRecord(E.sched_done[ws], t));
            MTR_CUDA(ctx, cudaStreamWaitEvent(E.walk_stream[ws], E.sched_done[ws], 0));
            MTR_CUDA(ctx, cudaStreamWaitEvent(E.walk_stream_big[ws], E.sched_done[ws], 0));
            eng_unitfinder<0><<<cfg.uf_ctas, walk_threads, kUfDynSmemSmall, E.walk_stream[ws]>>>(Psmall, ws * cfg.uf_ctas, kUfSmemWordsSmall, 0);
            eng_unitfinder<0><<<cfg.uf_ctas_big, walk_threads_big, kUfDynSmem, E.walk_stream_big[ws]>>>(P, cfg.uf_ctas * cfg.walk_streams + ws * cfg.uf_ctas_big, kUfSmemWords, 1);
            launches++;
        }
        eng_emit<<<(lay.n_chains + 255) / 256, 256, 0, t>>>(P, QS, QL, lay.n_chains);
        for (int k = 0; k < 2; k++) {
            if (qi[k] < 0) continue;
            const DpQueue &Q = Qs[qi[k]];
            eng_plan<<<1, 32, 0, t>>>(P, Q);
            eng_scatter<<<ctx->n_sm * 2, 256, 0, t>>>(P, Q);
            eng_zero_aux<<<ctx->n_sm * 2, 256, 0, t>>>(P, Q);
            launches += 3;
        }
        eng_publish<<<1, 1, 0, t>>>(P, QS, QL, (EngSnapshot *)E.h_snap.p);
        MTR_CUDA(ctx, cudaGetLastError());
        MTR_CUDA(ctx, cudaEventRecord(E.tick_done, t));
        launches += 7;
        const double tl1 = wall_ms();
        // the wave is a millisecond of GPU time: poll (a sleeping wait costs more than the wave)
        for (;;) {
            const cudaError_t e = cudaEventQuery(E.tick_done);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) MTR_CUDA(ctx, e);
            cudaGetLastError();
            std::this_thread::yield();
        }
        const double tl2 = wall_ms();
        // K3 over the tasks this wave has emitted, on the queues' own streams
        int emitted = 0;
        for (int k = 0; k < 2; k++) {
            if (qi[k] < 0 || snap->q_tasks[k] <= 0) continue;
            const int i = qi[k];
            WdpDevLaunch &L = LL[i];
            L.blocks_i32 = std::min((snap->q_slots[k][0] + 1) / 2, grid_cap);
            L.blocks_p16 = std::min((snap->q_slots[k][1] + 1) / 2, grid_cap);
            MTR_CUDA(ctx, cudaEventRecord(E.q_ev0[i], E.q_stream[i]));
            MTR_CUDA(ctx, wdp_launch_dev(L, E.q_stream[i]));
            MTR_CUDA(ctx, cudaEventRecord(E.q_ev1[i], E.q_stream[i]));
            MTR_CUDA(ctx, cudaEventRecord(E.q_done[i], E.q_stream[i]));
            E.q_busy[i] = true;
            emitted += snap->q_tasks[k];
            launches += (L.blocks_i32 > 0) + (L.blocks_p16 > 0);
            k3_uses++;
        }
        const double tl3 = wall_ms();
        t_launch += tl1 - tl0; t_wait += tl2 - tl1; t_k3 += tl3 - tl2;
        if (prof && (snap->waves <= 64 || snap->waves % 16 == 0))
            fprintf(stderr, "[mtr engine] ctx %p wave %d: unfinished %d accepted %d tasks short %d (slots %d+%d) long %d (slots %d+%d) deferred %d in flight %d\n", (void *)ctx, snap->waves, snap->unfinished, snap->n_accepted,
                    snap->q_tasks[0], snap->q_slots[0][0], snap->q_slots[0][1], snap->q_tasks[1], snap->q_slots[1][0], snap->q_slots[1][1], snap->deferred, snap->in_flight);
        if (snap->error) break;
        if (snap->unfinished <= 0) break;
        const bool moved = !(snap->tasks_total == last_tasks && snap->progress_sig == last_sig && snap->n_accepted == last_accepted && snap->unfinished == last_unfinished);
        last_tasks = snap->tasks_total; last_sig = snap->progress_sig; last_accepted = snap->n_accepted; last_unfinished = snap->unfinished;
        if (moved) { quiet = 0; continue; }
        // nothing moved in this wave: every chain at work waits for a walk or a DP -- wait for the first queue to finish
        // instead of spinning waves (they scan every chain slot); nothing in flight at all for many waves = a bug: fail loudly
        idle_ticks++;
        bool any_busy = false;
        for (int i = 0; i < kQueues; i++) any_busy = any_busy || E.q_busy[i];
        if (snap->in_flight == 0 && !any_busy) {
            // nothing in flight and nothing moves: chains that wait for shared DP results run their own instead (should
            // never be needed; Counters::unshared says how often it was)
            if (quiet == 8 || quiet == 32) { eng_unshare<<<(lay.n_chains + 255) / 256, 256, 0, t>>>(P, lay.n_chains); rescues++; }
            if (++quiet > 64) {
                mtr_set_error(ctx, "engine_run: no progress after wave %d (%d reads unfinished, %d tasks deferred)", snap->waves, snap->unfinished, snap->deferred);
                if (getenv("MTR_ENGINE_DUMP")) { eng_dump_stuck<<<1, 1, 0, t>>>(P, lay.n_chains); cudaStreamSynchronize(t); }
                return MTR_ECUDA;
            }
            continue;
        }
        quiet = 0;
        if (any_busy) {
            // poll the busy queues for up to 200 us
            const double w0 = wall_ms();
            for (bool got = false; !got && wall_ms() - w0 < 0.2;) {
                for (int i = 0; i < kQueues && !got; i++) if (E.q_busy[i] && cudaEventQuery(E.q_done[i]) == cudaSuccess) got = true;
                cudaGetLastError();
                if (!got) std::this_thread::yield();
            }
        } else {
            const double w0 = wall_ms();
            while (wall_ms() - w0 < 0.05) std::this_thread::yield();   // walks only: they publish through the chains' stages
        }
    }
    for (int i = 0; i < cfg.walk_streams; i++) { MTR_CUDA(ctx, cudaStreamSynchronize(E.walk_stream[i])); MTR_CUDA(ctx, cudaStreamSynchronize(E.walk_stream_big[i])); }   // walks nobody waits for any more
    for (int i = 0; i < kQueues; i++) MTR_CUDA(ctx, reap(i, true));                                       // ... and DPs of dropped candidates
    s = t;
    Counters *hc = (Counters *)E.h_ctr.p;
    MTR_CUDA(ctx, cudaMemcpyAsync(hc, P.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    MTR_CUDA(ctx, cudaStreamSynchronize(s));
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
        fprintf(stderr, "[mtr engine] ctx %p: rescue passes %lld, chains unshared %d\n", (void *)ctx, rescues, hc->unshared);
    if (rescues > 0 && !prof) fprintf(stderr, "[mtr engine] note: %lld rescue pass(es), %d chain(s) re-ran shared DPs on their own\n", rescues, hc->unshared);
    if (prof)
        fprintf(stderr, "[mtr engine] ctx %p: host ms: launching waves %.1f, waiting for waves %.1f, launching K3 %.1f (%lld uses), idle waves %lld, whole call so far %.1f (di %.1f)\n", (void *)ctx, t_launch, t_wait, t_k3, k3_uses, idle_ticks, wall_ms() - t_wall0, di_ms);
    if (prof)
        fprintf(stderr, "[mtr engine] ctx %p: %d reads, %d waves | unit finder: %llu tables (direct %llu compact %llu wide %llu), %llu with walks, %llu walks | Mticks: build %.1f list %.1f walk fwd %.1f bwd %.1f, slowest task %.3f, slowest walk %.3f | walk steps %llu (memo hits %llu, steps reaching level 4 %llu), probe rounds %llu, failed walks %llu | dp %.1f ms uf %.1f ms\n", (void *)ctx, n, hc->waves,
                hc->tables, hc->prof_kind[2], hc->prof_kind[1], hc->prof_kind[0], hc->prof_walk_tasks, hc->walks, hc->prof_build / 1e6, hc->prof_list / 1e6, hc->prof_walk[0] / 1e6, hc->prof_walk[1] / 1e6,
                hc->prof_max_task / 1e6, hc->prof_max_walk / 1e6, hc->prof_steps, hc->prof_memo_hits, hc->prof_deep_steps, hc->prof_probe_rounds, hc->prof_fail_walks, dp_ms, t_wait);
    if (prof)
        for (int f = 0; f < 2; f++) {
            const unsigned long long *k = hc->prof_k3 + 5 * f;
            fprintf(stderr, "[mtr engine] ctx %p: K3 %s: %llu slots, fill %.1f Mticks over %llu slot rows (%.0f ticks per row), traceback %.1f Mticks over %llu rows (%.0f ticks per row)\n", (void *)ctx,
                    f ? "int16x2" : "int32", k[3], k[0] / 1e6, k[2], k[2] ? (double)k[0] / k[2] : 0.0, k[1] / 1e6, k[4], k[4] ? (double)k[1] / k[4] : 0.0);
        }
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
        MTR_CUDA(ctx, cudaStreamSynchronize(s));
    }
    export_repeats((const Accepted *)E.h_acc.p, na, E.reps, E.units, first);
    if (repeats) *repeats = E.reps.data();
    if (n_repeats) *n_repeats = na;
    if (units) *units = E.units.data();
    if (stats) {
        export_stats(*hc, stats);
        stats->launches = launches;
        stats->di_ms = di_ms; stats->dp_ms = dp_ms; stats->uf_ms = t_wait;
        stats->h2d_bytes = di_h2d + (int64_t)sizeof(ReadDesc) * n + (int64_t)sizeof(Read) * n_slots + (int64_t)sizeof(Counters);
        stats->d2h_bytes = (int64_t)sizeof(Accepted) * na + (int64_t)sizeof(Counters);
        stats->wall_ms = wall_ms() - t_wall0;
    }
    ctx->stats.launches = (int32_t)std::min<long long>(launches, 0x7fffffff);
    return MTR_OK;
}
