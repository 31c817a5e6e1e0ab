// eng_host.h -- host-side layout of one engine group (shared by eng.cu and by the CPU twin in tests/hostsim).
// Everything the waves touch is carved out of ONE allocation (plus the direction matrices and the unit-finder scratch,
// which are sized separately), so binding eng::Ptrs is pointer arithmetic on a base address -- device or host.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "eng_core.h"

namespace eng {

struct Config {
    int n_reads = 0;             // read SLOTS: reads of the batch at work at the same time
    int n_total = 0;             // reads of the batch
    long long total_bases = 0;
    int max_len = 0;
    int uf_ctas = 0;             // persistent ctas of the small-window walk kernel (each owns a scratch slice and a WIDE table)
    int uf_ctas_big = 0;         // ... of the big-window walk kernel
    int walk_streams = 4;        // walk kernel instances in flight (one scratch bank of uf_ctas slices each)
    int polish_ctas = 0;         // ... of the polish kernel (scratch slices behind the walk banks: it runs beside the walks)
    unsigned compact_cap = kCompactCap;
    int direct_max_k = 7;
    int task_cap = 0;            // DP tasks per use of a short queue
    int acc_cap = 0;             // accepted repeats of the whole group
    long long aux_cap = 0;       // its consensus pool (int32)
    long long dir_cap = 0;       // its direction arena (bytes)
    int long_rows = 2048;        // DP tasks with at least this many rows go to the long queues
    int long_task_cap = 0;       // tasks per use of a long queue
    long long long_dir_cap = 0, long_aux_cap = 0;   // its direction arena (bytes) and consensus pool (int32)
    int walk_cap = 0;
};

struct Layout {
    size_t reads, descs, chains, units, scores, results, polish_list, walk_ring, walk_ring_big, acc, ctr, zero_begin, total;
    struct Q { size_t tasks_in, tasks, aux, hist, seg_task, seg_slot, bucket_cursor, class_begin, slot_counter, qc; } q[kQueues];
    int n_chains;
    unsigned table_cap;
    long long uf_stride;
};

inline Config default_config(int n_slots, int n_total, long long total_bases, int max_len, int n_sm)
{
    Config c;
    const int n_reads = std::max(1, std::min(n_slots, n_total));
    c.n_reads = n_reads; c.n_total = n_total; c.total_bases = total_bases; c.max_len = max_len;
    // every cta owns a WIDE table sized for the longest read: at most 1 GB of them per group
    unsigned cap = 64;
    while (cap < 2u * (unsigned)(max_len + 8)) cap <<= 1;
    c.uf_ctas = (int)std::max<long long>(4, std::min<long long>(n_sm / 2, (1LL << 28) / ((long long)cap * 8)));
    c.uf_ctas_big = c.uf_ctas;
    c.uf_ctas = std::max(c.uf_ctas, std::min(2 * n_sm, 4 * c.uf_ctas));   // small windows: several ctas per SM
    c.polish_ctas = std::max(c.uf_ctas, std::min(2 * n_sm, 4 * c.uf_ctas));   // polish runs inside the wave: as wide as the GPU
    const long long n_chains = (long long)n_reads * kSets * kMaxK;
    c.task_cap = (int)std::min<long long>(std::max<long long>(4096, n_chains / 2), 1 << 19);
    c.acc_cap = (int)std::min<long long>((long long)n_total * 32 + total_bases / 48 + 64, 1 << 24);
    c.aux_cap = std::max<long long>((long long)c.task_cap / 8 * 512, 2 * 4500);
    // direction arenas: full size for a batch that fills the slots, scaled down for small batches (tests, short files)
    c.dir_cap = std::min<long long>(2LL << 30, std::max<long long>(std::max<long long>(16LL << 20, 1024LL * max_len), total_bases * 64));
    c.long_task_cap = (int)std::min<long long>(std::max<long long>(1024, 8LL * n_reads), 1 << 18);
    c.long_dir_cap = std::min<long long>(3LL << 30, std::max<long long>(std::max<long long>(16LL << 20, 1024LL * max_len), total_bases * 64));
    c.long_aux_cap = std::max<long long>((long long)c.long_task_cap / 8 * 512, 2 * 4500);
    // walk queue: a power of two well above the chains that can be queued at once (entries of cancelled walks linger
    // until a walk kernel pops them)
    c.walk_cap = 4096;
    while (c.walk_cap < 8 * n_chains && c.walk_cap < (1 << 28)) c.walk_cap <<= 1;
    return c;
}

inline Layout make_layout(const Config &c)
{
    Layout l;
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; };
    l.n_chains = c.n_reads * kSets * kMaxK;
    l.reads = take(sizeof(Read) * (size_t)std::max(c.n_reads, 1));
    l.descs = take(sizeof(ReadDesc) * (size_t)std::max(c.n_total, 1));
    l.chains = take(sizeof(Chain) * (size_t)std::max(l.n_chains, 1));
    l.units = take((size_t)std::max(l.n_chains, 1) * 4 * kUnitStride);
    l.scores = take((size_t)std::max(l.n_chains, 1) * 3 * kUnitStride);
    l.results = take(sizeof(mtr_wdp_result) * (size_t)std::max(l.n_chains, 1) * kResSlots);
    l.polish_list = take(4 * (size_t)std::max(l.n_chains, 1));
    l.walk_ring = take(4 * (size_t)c.walk_cap);
    l.walk_ring_big = take(4 * (size_t)c.walk_cap);
    l.acc = take(sizeof(Accepted) * (size_t)c.acc_cap);
    for (int i = 0; i < kQueues; i++) {
        const int cap = i < kShortInst ? c.task_cap : c.long_task_cap;
        l.q[i].tasks_in = take(sizeof(WdpTask) * (size_t)cap);
        l.q[i].tasks = take(sizeof(WdpTask) * (size_t)cap);
        l.q[i].aux = take(4 * (size_t)(i < kShortInst ? c.aux_cap : c.long_aux_cap));
    }
    // everything from here on starts as zero
    l.zero_begin = at;
    l.ctr = take(sizeof(Counters));
    for (int i = 0; i < kQueues; i++) {
        l.q[i].hist = take(4 * (size_t)kSegs);
        l.q[i].seg_task = take(4 * (size_t)(kSegs + 1));
        l.q[i].seg_slot = take(4 * (size_t)(kSegs + 1));
        l.q[i].bucket_cursor = take(4 * (size_t)kSegs);
        l.q[i].class_begin = take(4 * (size_t)(WDP_NCLASS + 1));
        l.q[i].slot_counter = take(4 * (size_t)WDP_NCLASS);
        l.q[i].qc = take(sizeof(QueueCtr));
    }
    l.total = at;
    unsigned cap = 64;
    while (cap < 2u * (unsigned)(c.max_len + 8)) cap <<= 1;
    l.table_cap = cap;
    l.uf_stride = (kScratchFixed + 255) & ~255LL;
    return l;
}

inline DpQueue bind_queue(void *base, const Layout &l, const Config &c, int i)
{
    unsigned char *b = (unsigned char *)base;
    DpQueue q;
    q.tasks_in = (WdpTask *)(b + l.q[i].tasks_in); q.tasks = (WdpTask *)(b + l.q[i].tasks);
    q.task_cap = i < kShortInst ? c.task_cap : c.long_task_cap;
    q.hist = (int *)(b + l.q[i].hist); q.seg_task = (int *)(b + l.q[i].seg_task); q.seg_slot = (int *)(b + l.q[i].seg_slot);
    q.bucket_cursor = (int *)(b + l.q[i].bucket_cursor); q.class_begin = (int *)(b + l.q[i].class_begin);
    q.slot_counter = (int *)(b + l.q[i].slot_counter);
    q.aux = (int *)(b + l.q[i].aux); q.aux_cap = i < kShortInst ? c.aux_cap : c.long_aux_cap;
    q.dir_cap = i < kShortInst ? c.dir_cap : c.long_dir_cap;
    q.qc = (QueueCtr *)(b + l.q[i].qc);
    q.id = i;
    return q;
}
inline DpQueue no_queue() { DpQueue q; memset(&q, 0, sizeof q); return q; }

inline Ptrs bind(void *base, const Layout &l, const Config &c)
{
    unsigned char *b = (unsigned char *)base;
    Ptrs P;
    memset(&P, 0, sizeof P);
    P.reads = (Read *)(b + l.reads); P.n_reads = c.n_reads;
    P.descs = (const ReadDesc *)(b + l.descs); P.n_total = c.n_total;
    P.chains = (Chain *)(b + l.chains);
    P.units = b + l.units; P.scores = b + l.scores;
    P.results = (mtr_wdp_result *)(b + l.results);
    P.polish_list = (int *)(b + l.polish_list);
    P.walk_ring = (int *)(b + l.walk_ring); P.walk_ring_big = (int *)(b + l.walk_ring_big); P.walk_ring_mask = (unsigned)c.walk_cap - 1u;
    for (int i = 0; i < kQueues; i++) P.aux_of[i] = (int *)(b + l.q[i].aux);
    P.long_rows = c.long_rows;
    P.acc = (Accepted *)(b + l.acc); P.acc_cap = c.acc_cap;
    P.ctr = (Counters *)(b + l.ctr);
    P.table_cap = l.table_cap; P.uf_stride = l.uf_stride;
    P.share_search = getenv("MTR_ENGINE_SHARE") ? atoi(getenv("MTR_ENGINE_SHARE")) : 1;
    P.compact_cap = c.compact_cap; P.direct_max_k = c.direct_max_k;
    return P;
}

// the reads of the batch (host side; copied into the group's buffer) and the initial state of the slots: all free
inline void init_descs(std::vector<ReadDesc> &out, const int64_t *word_off, const int32_t *len, int n, const int *order = nullptr)
{
    // order: the sequence in which the slots take the reads (heaviest first keeps the tail of a group short); ids stay
    // the reads' positions in the batch, so nothing downstream sees the order
    std::vector<long long> pos((size_t)n + 1, 0);
    for (int r = 0; r < n; r++) pos[r + 1] = pos[r] + len[r];
    out.assign((size_t)n, ReadDesc());
    for (int i = 0; i < n; i++) {
        const int r = order ? order[i] : i;
        ReadDesc &d = out[i];
        d.word_off = word_off[r]; d.pos_off = pos[r]; d.L = len[r]; d.id = r;
    }
}
inline void init_slots(std::vector<Read> &out, int n_slots)
{
    out.assign((size_t)n_slots, Read());
    for (Read &rs : out) { memset(&rs, 0, sizeof rs); rs.phase = 1; rs.id = -1; }
}

// converts the accepted list into the ABI records, ordered by (read, insertion order)
inline void export_repeats(const Accepted *acc, int n, std::vector<mtr_repeat> &reps, std::vector<uint8_t> &units, int first_read = 0)
{
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (acc[a].read != acc[b].read) return acc[a].read < acc[b].read;
        return acc[a].seq < acc[b].seq;
    });
    reps.resize(n);
    units.clear();
    for (int i = 0; i < n; i++) {
        const Accepted &a = acc[order[i]];
        mtr_repeat &r = reps[i];
        r.read = a.read + first_read; r.seq = a.seq;
        r.rep_start = a.rec.rep_start; r.rep_end = a.rec.rep_end; r.repeat_len = a.rec.repeat_len; r.rep_period = a.rec.period;
        r.num_freq_unit = a.rec.units; r.num_matches = a.rec.nm; r.num_mismatches = a.rec.nx; r.num_insertions = a.rec.ni;
        r.num_deletions = a.rec.nd; r.kmer = a.rec.kmer; r.match_gain = a.rec.gain; r.mismatch_penalty = a.rec.mis;
        r.indel_penalty = a.rec.indel;
        r.unit_off = (int64_t)units.size();
        units.insert(units.end(), a.unit, a.unit + std::min(a.rec.period, kUnitStride));
    }
}

inline void export_stats(const Counters &c, mtr_engine_stats *s)
{
    if (!s) return;
    s->waves = c.waves; s->candidates = (int64_t)c.candidates; s->dp_jobs = (int64_t)c.jobs; s->dp_tasks = (int64_t)c.tasks_total;
    s->dp_cells = (int64_t)c.cells; s->dp_slot_cells = (int64_t)c.slot_cells; s->dp_dir_bytes = (int64_t)c.dir_bytes;
    s->spec_cells = (int64_t)c.spec_cells; s->shared_cells = (int64_t)c.shared_cells; s->dp_cells_p16 = (int64_t)c.cells_p16; s->tables = (int64_t)c.tables; s->table_positions = (int64_t)c.table_positions;
    s->walks = (int64_t)c.walks; s->repeats = c.n_accepted; s->wrapdp_messages = c.msgs;
}

}  // namespace eng
