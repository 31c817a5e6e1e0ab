// tests/hostsim/sim_engine.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The CPU twin of mtr_b200/csrc/eng.cu: the SAME engine code (mtr_b200/csrc/eng_core.h, compiled here by g++ with one
// lane per warp, see simt.h) driven wave by wave on host memory.  What is a kernel launch on the GPU is a plain loop
// over the work items here; the wrap-around DP tasks the engine emits are answered by the oracle's DP, the directional
// index by the oracle's (sim_device.cpp).  So the per-read scheduler with its look-ahead and pruning, the maxFreq
// gate, count tables, greedy walks with memo and cycle cut, polish, the revise chain with its vote, the two-penalty
// pick, the DP task emission / bucket sort and the accepted-repeat list are all checked against the reference's
// digests on a machine without a GPU.  Never linked into libmtr_b200.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include "../../mtr_b200/csrc/eng_host.h"
#include "../../oracle/mtr_oracle.h"

using namespace eng;

struct EngState {
    std::vector<unsigned char> main_buf, scratch, wide;
    std::vector<int> end, w;
    std::vector<mtr_repeat> reps;
    std::vector<uint8_t> units;
    int speculate = 8;
};

void eng_state_free(mtr_ctx *ctx) { delete ctx->eng; ctx->eng = nullptr; }

extern "C" int mtr_di_run_range(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off,
                                double *di, int32_t *end, int32_t *w, int first, int count);
// provided by sim_device.cpp
const uint32_t *sim_packed(mtr_ctx *ctx);
mtro_ctx *sim_oracle(mtr_ctx *ctx, int manhattan);

extern "C" int mtr_engine_set_speculate(mtr_ctx *ctx, int depth)
{
    if (!ctx || depth < 0) return MTR_EINVAL;
    if (!ctx->eng) ctx->eng = new EngState();
    ctx->eng->speculate = depth;
    return MTR_OK;
}

extern "C" double mtr_engine_dp_busy_ms(int, int) { return 0.0; }

// walk-queue entries start as -1 and go back to -1 when taken (eng.cu: a taker waits for an entry that is still being written)
static int take_entry(int *ring, unsigned at)
{
    const int v = ring[at];
    if (v < 0) { fprintf(stderr, "hostsim: walk queue entry %u taken before it was written\n", at); abort(); }
    ring[at] = -1;
    return v;
}

static void run_dp_tasks(mtr_ctx *ctx, const Ptrs &P, const DpQueue &Q, mtro_ctx *o)
{
    const int total = Q.class_begin[WDP_NCLASS];
    std::vector<int> x, u;
    for (int i = 0; i < total; i++) {
        const WdpTask &t = Q.tasks[i];
        x.assign((size_t)t.rows + 2, 0);
        for (int r = 1; r <= t.rows; r++) { const long long b = t.base0 + r; x[r] = (int)((P.packed[b >> 4] >> ((b & 15) * 2)) & 3u); }
        u.assign((size_t)t.ulen + 2, 0);
        for (int j = 1; j <= t.ulen; j++) u[j] = P.units[t.unit_off + j - 1];
        for (int p = 0; p < (int)t.n_param; p++) {
            mtro_dp_result r;
            int *cons = nullptr, *miss = nullptr;
            if (t.mode == MTR_TB_CONSENSUS) { cons = Q.aux + t.aux_off; miss = cons + (size_t)(t.ulen + 1) * 5; }
            mtro_wrap_dp(o, x.data(), t.rows, u.data(), t.ulen, t.gain[p], t.mis[p], t.indel[p], t.mode, &r, cons, miss, nullptr, nullptr);
            mtr_wdp_result &d = P.results[t.result_idx + p];
            d.best = r.best; d.max_i = r.max_i; d.max_j = r.max_j; d.end_i = r.end_i; d.end_j = r.end_j;
            d.n_match = r.n_match; d.n_mismatch = r.n_mismatch; d.n_ins = r.n_ins; d.n_del = r.n_del; d.n_scanned = r.n_scanned;
            d.path_len = r.path_len; d.flags = 0;
            P.chains[t.result_idx >> kWdpOwnerShift].pending--;              // what wdp_traceback_dev does on the GPU
            P.ctr->dp_pending--;
        }
    }
    (void)ctx;
}

extern "C" int mtr_engine_run_range(mtr_ctx *ctx, int first, int count, int manhattan, float min_match_ratio, const uint16_t *stale,
                                    const int64_t *stale_off, const mtr_repeat **repeats, int64_t *n_repeats, const uint8_t **units,
                                    mtr_engine_stats *stats);
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
    if (first < 0 || count < 0 || first + count > ctx->n_reads) return MTR_EINVAL;
    const int n = count;
    if (n == 0) return MTR_OK;
    if (!ctx->eng) ctx->eng = new EngState();
    EngState &E = *ctx->eng;
    std::vector<int64_t> pos_off((size_t)first + n + 1, 0);
    int max_len = 0;
    for (int r = 0; r < n; r++) { pos_off[first + r + 1] = pos_off[first + r] + ctx->len[first + r]; max_len = std::max(max_len, (int)ctx->len[first + r]); }
    E.end.assign((size_t)pos_off[first + n] + 1, -1); E.w.assign((size_t)pos_off[first + n] + 1, -1);
    int rc = mtr_di_run_range(ctx, manhattan, stale, stale_off, pos_off.data(), nullptr, E.end.data(), E.w.data(), first, n);
    if (rc) return rc;

    // MTR_ENGINE_SLOTS: reads at work at the same time (a slot takes the next read of the batch when its read has finished)
    int n_slots = n;
    if (const char *e = getenv("MTR_ENGINE_SLOTS")) n_slots = std::max(1, atoi(e));
    Config cfg = default_config(n_slots, n, pos_off[first + n], max_len, 1);
    n_slots = cfg.n_reads;
    cfg.uf_ctas = 1;
    // MTR_SIM_COMPACT_CAP / MTR_SIM_DIRECT_K: shrink the shared-memory table layouts so that small test windows reach
    // the COMPACT and the WIDE (+ probe cache) paths too
    if (const char *e = getenv("MTR_SIM_COMPACT_CAP")) cfg.compact_cap = (unsigned)std::max(64, std::min(4096, atoi(e)));
    if (const char *e = getenv("MTR_SIM_DIRECT_K")) cfg.direct_max_k = std::max(0, std::min(7, atoi(e)));
    // small per-wave budgets on request, to exercise the deferral path on the CPU
    if (const char *e = getenv("MTR_ENGINE_DIR_KB")) cfg.dir_cap = std::max(1LL, atoll(e)) << 10;
    if (const char *e = getenv("MTR_ENGINE_TASK_CAP")) cfg.task_cap = std::max(4, atoi(e));
    // MTR_ENGINE_LONG_ROWS: tasks with at least this many rows take the long queues; MTR_SIM_LONG_EVERY = n: a long queue is
    // free only every n-th wave and its results arrive n - 1 waves late (chains wait in WAIT_*, dropped candidates leave
    // zombie sets behind)
    if (const char *e = getenv("MTR_ENGINE_LONG_ROWS")) cfg.long_rows = std::max(1, atoi(e));
    const int long_every = getenv("MTR_SIM_LONG_EVERY") ? std::max(1, std::min(5, atoi(getenv("MTR_SIM_LONG_EVERY")))) : 1;
    const Layout lay = make_layout(cfg);
    E.main_buf.assign(lay.total, 0);
    E.scratch.assign((size_t)lay.uf_stride, 0);
    E.wide.assign((size_t)lay.table_cap * 8, 0);
    std::vector<unsigned> smem((size_t)kUfSmemWords, 0u);
    int sh[16] = {0}, near[4 * kTiesNear] = {0};
    std::vector<MemoEntry> memos(2 * kMemoSlots);
    memset(memos.data(), 0, sizeof(MemoEntry) * memos.size());
    const Cta cta = cta_of_warp(sh);
    Ptrs P = bind(E.main_buf.data(), lay, cfg);
    P.packed = sim_packed(ctx);
    P.end = E.end.data(); P.w = E.w.data();
    P.uf_scratch = E.scratch.data();
    P.uf_wide = E.wide.data();
    P.min_match_ratio = min_match_ratio;
    P.speculate = E.speculate;
    if (const char *e = getenv("MTR_SPECULATE")) P.speculate = std::max(0, atoi(e));
    for (int i = 0; i < cfg.walk_cap; i++) { P.walk_ring[i] = -1; P.walk_ring_big[i] = -1; }
    Ptrs Psmall = P;
    Psmall.compact_cap = std::min(P.compact_cap, kCompactCapSmall);
    std::vector<Read> slots;
    std::vector<ReadDesc> descs;
    init_slots(slots, n_slots);
    // the order in which the slots take the reads is scheduling only (eng.cu: heaviest first); MTR_SIM_READ_ORDER = 1: last first
    std::vector<int> order((size_t)n);
    for (int r = 0; r < n; r++) order[r] = getenv("MTR_SIM_READ_ORDER") && atoi(getenv("MTR_SIM_READ_ORDER")) == 1 ? n - 1 - r : r;
    init_descs(descs, ctx->word_off.data() + first, ctx->len.data() + first, n, order.data());
    memcpy(P.reads, slots.data(), sizeof(Read) * (size_t)n_slots);
    memcpy((void *)P.descs, descs.data(), sizeof(ReadDesc) * (size_t)n);
    P.ctr->unfinished = n;
    mtro_ctx *o = sim_oracle(ctx, manhattan);
    std::vector<unsigned long long> inline_tab(kInlineSlots);
    const Scratch S = scratch_of(P, 0);
    unsigned long long last_sig = ~0ull;
    int idle = 0;
    std::vector<unsigned> tails, tails_big;
    DpQueue Qs[kQueues];
    for (int i = 0; i < kQueues; i++) Qs[i] = bind_queue(E.main_buf.data(), lay, cfg, i);
    const DpQueue none = no_queue();
    // K3 of a queue "runs" for every[k] waves: its results arrive that many waves after the emission (minus one), and the
    // queue cannot be used again before
    const int every[2] = {getenv("MTR_SIM_SHORT_EVERY") ? std::max(1, std::min(5, atoi(getenv("MTR_SIM_SHORT_EVERY")))) : 1, long_every};
    int due[kQueues];
    for (int i = 0; i < kQueues; i++) due[i] = -1;             // wave at which the queue's results arrive (-1: free)
    int wave = 0;
    while (P.ctr->unfinished > 0 && !P.ctr->error) {
        wave++;
        for (int i = 0; i < kQueues; i++) if (due[i] >= 0 && wave >= due[i]) { run_dp_tasks(ctx, P, Qs[i], o); due[i] = -1; }
        int qi[2] = {-1, -1};
        for (int k = 0; k < kShortInst && qi[0] < 0; k++) { const int i = (wave + k) % kShortInst; if (due[i] < 0) qi[0] = i; }
        for (int k = 0; k < kLongInst && qi[1] < 0; k++) { const int i = kShortInst + (wave + k) % kLongInst; if (due[i] < 0) qi[1] = i; }
        if (every[0] > 1 && wave % 3 == 0) qi[0] = -1;          // now and then no queue is free at all
        if (every[1] > 1 && wave % 4 == 0) qi[1] = -1;
        const DpQueue &QS = qi[0] >= 0 ? Qs[qi[0]] : none, &QL = qi[1] >= 0 ? Qs[qi[1]] : none;
        wave_begin(P, QS, QL);
        for (int c = 0; c < lay.n_chains; c++) { take_shared(P, c); if (chain_ready(P, c)) advance_chain(P, c); }
        for (int i = 0; i < P.ctr->n_polish; i++) polish_chain(P, P.polish_list[i], S, cta, smem.data());
        for (int r = 0; r < n_slots; r++) sched_read(P, r, inline_tab.data(), kInlineSlots, sh);
        // the walk queue: with MTR_SIM_WALK_LAG = n the entries pushed by a scheduler pass are only walked n waves later,
        // like walks that are still running when the next passes start (dropped candidates then meet chains in ST_WALK)
        {
            static const int lag = getenv("MTR_SIM_WALK_LAG") ? std::min(5, atoi(getenv("MTR_SIM_WALK_LAG"))) : 0;
            tails.push_back(P.ctr->walk_tail);
            const unsigned limit = (int)tails.size() > lag ? tails[tails.size() - 1 - (size_t)lag] : 0u;
            while ((int)(limit - P.ctr->walk_head) > 0) walk_chain(Psmall, take_entry(P.walk_ring, P.ctr->walk_head++ & P.walk_ring_mask), S, cta, smem.data(), near, memos.data());
            // (the windows that need the big shared-memory count table have a queue and a kernel of their own)
            tails_big.push_back(P.ctr->walk_tail_big);
            const unsigned limit_big = (int)tails_big.size() > lag ? tails_big[tails_big.size() - 1 - (size_t)lag] : 0u;
            while ((int)(limit_big - P.ctr->walk_head_big) > 0) walk_chain(P, take_entry(P.walk_ring_big, P.ctr->walk_head_big++ & P.walk_ring_mask), S, cta, smem.data(), near, memos.data());
        }
        // (one thread per chain on the GPU, in any order: MTR_SIM_EMIT_ORDER = 1 emits in descending, 2 in a scrambled order)
        {
            static const int order = getenv("MTR_SIM_EMIT_ORDER") ? atoi(getenv("MTR_SIM_EMIT_ORDER")) : 0;
            const int nc = lay.n_chains;
            if (order == 0) for (int c = 0; c < nc; c++) emit_chain(P, QS, QL, c);
            else if (order == 1) for (int c = nc - 1; c >= 0; c--) emit_chain(P, QS, QL, c);
            else {
                long long step = 7919;
                while (std::__gcd((long long)nc, step) != 1) step++;
                for (long long i = 0, c = (wave * 31) % nc; i < nc; i++, c = (c + step) % nc) emit_chain(P, QS, QL, (int)c);
            }
        }
        for (int k = 0; k < 2; k++) {
            if (qi[k] < 0) continue;
            const DpQueue &Q = Qs[qi[k]];
            plan_tasks(P, Q);
            const int nt = std::min(Q.qc->n_tasks, Q.task_cap);
            for (int i = 0; i < nt; i++) scatter_task(Q, i);
            const long long na = std::min<long long>((long long)Q.qc->aux_used, Q.aux_cap);
            for (long long i = 0; i < na; i++) Q.aux[i] = 0;
            if (every[k] <= 1) run_dp_tasks(ctx, P, Q, o);
            else due[qi[k]] = wave + every[k] - 1;
        }
        const unsigned long long sig = (unsigned long long)(unsigned)P.ctr->dp_pending * 7919ull + P.ctr->tasks_total * 1315423911ull + P.ctr->tables * 2654435761ull + (unsigned)P.ctr->progress * 97ull + (unsigned)P.ctr->unfinished;
        if (sig == last_sig) {
            if (++idle > 24) { mtr_set_error(ctx, "hostsim engine_run: no progress (wave %d, %d tasks deferred)", P.ctr->waves, P.ctr->deferred); return MTR_ECUDA; }
        } else idle = 0;
        last_sig = sig;
    }
    const Counters &c = *P.ctr;
    if (c.error) {
        switch (c.error) {
        case ERR_WRAPCAP: mtr_set_error(ctx, "You need to increse the value of WrapDPsize."); return MTR_ERANGE;
        case ERR_EMPTY_UNIT: mtr_set_error(ctx, "the revised repeat unit is empty (read %d)", c.error_read); return MTR_ERANGE;
        case ERR_ACCEPTED_FULL: mtr_set_error(ctx, "engine_run: more than %d accepted repeats in one group", cfg.acc_cap); return MTR_ENOMEM;
        default: mtr_set_error(ctx, "engine_run: device error %d", c.error); return MTR_ECUDA;
        }
    }
    export_repeats(P.acc, c.n_accepted, E.reps, E.units, first);
    if (repeats) *repeats = E.reps.data();
    if (n_repeats) *n_repeats = c.n_accepted;
    if (units) *units = E.units.data();
    export_stats(c, stats);
    if (stats) {                                                // what the real engine would move over PCIe
        stats->h2d_bytes = (int64_t)sizeof(ReadDesc) * n + (int64_t)sizeof(Read) * n_slots + (int64_t)sizeof(Counters);
        stats->d2h_bytes = (int64_t)sizeof(Accepted) * c.n_accepted + (int64_t)sizeof(Counters);
    }
    return MTR_OK;
}
