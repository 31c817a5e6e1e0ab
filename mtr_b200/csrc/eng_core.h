// eng_core.h -- the resident engine: mTR's per-read candidate loop, unit finder and revise chain as device code.
//
// Replaces, for a whole group of reads at once and without the host in the loop,
//   handle_one_TR's candidate loop            /root/reference/handle_one_read.c:227-246  (eng::sched_read)
//   find_tandem_repeat / _sub                 handle_one_read.c:77-154                   (sched_read, advance_chain)
//   search_De_Bruijn_graph                    consensus.c:507-582                         (gate, walk_chain, advance_chain)
//     init_inputString, generate_freqNode_return_list_maxNodes, freq_node   :37-60,132-253   (Table)
//     search_De_Bruijn_graph_forward / _backward                            :269-505         (walk)
//   polish_repeat                             consensus.c:584-704                         (polish_chain)
//   revise_representative_unit (+ the vote of revise_representative_unit_sub)  :964-1013,1048-1087  (advance_chain, vote_unit)
//   wrap_around_DP's pick between the two penalty sets   wrap_around_DP.c:357-429         (advance_chain)
// The wrap-around DPs themselves are the K3 kernels of wdp.cu; this file only writes their task records (emit_chain,
// plan_tasks, scatter_task) and reads their results.
//
// Execution model: a group of reads advances in WAVES.  One wave is the kernel sequence
//     advance -> polish -> sched -> walk -> emit -> plan -> scatter -> K3 fill + traceback
// and every kernel is a loop over independent work items (a chain, a read, a DP task) handled by one warp or one
// thread.  All cross-kernel state (per-read cursor and candidate ring, chain records, work lists, counters) lives in
// device memory, so the host only launches waves until the `unfinished` counter reaches zero.
//
// The code is written against simt.h: warp-uniform control flow, lane-strided data loops.  tests/hostsim compiles this
// very header with one lane per warp and runs the waves on the CPU (DP and directional index answered by the oracle).
#pragma once
#include "simt.h"
#include "mtr_internal.h"

namespace eng {
using namespace simt;

// state shared by the lanes of a warp is written by lane 0 only, fenced on both sides: every lane has finished reading
// the old value before it changes, and sees the new one afterwards
#define ENG_LANE0(...) do { wsync(); if (lane() == 0) { __VA_ARGS__; } wsync(); } while (0)

constexpr int kMaxPeriod = 500;              // MAX_PERIOD, mTR.h:34
constexpr int kMaxTies = 1024;               // MAX_tiebreaks, mTR.h:46
constexpr long long kWrapCap = 200000000LL;  // WrapDPsize, mTR.h:51
constexpr int kUnitStride = 512;             // bytes reserved per stored unit / score string
constexpr int kRing = 64;                    // candidates of one read in flight (most of them die at the maxFreq gate)
constexpr int kSets = 16;                    // ... of which this many may hold chains (one chain per k)
constexpr int kMaxK = 11;                    // k values per candidate: 2..10, 2..12 or 5..15 (handle_one_read.c:105-120)
constexpr int kInlineWindow = 512;           // windows up to this many bases pass the maxFreq gate inside sched_read
constexpr int kInlineSlots = 2048;           // count-table slots of a scheduler warp (>= 2 * (kInlineWindow + 2), power of two)
constexpr int kMemoSlots = 512;              // walk-memo entries per direction (shared memory, 16 bytes each)
constexpr int kSchedBudget = 96;             // candidates one sched_read call may start (bounds the latency of a wave)
constexpr int kResSlots = 1 << kWdpOwnerShift; // DP result slots of a chain: search direction d, penalty set s at 2d + s; revise at kResRevise
constexpr int kResRevise = 4;
constexpr int kDpClasses = 20;               // 10 int32 + 10 paired int16x2 fill classes (wdp.cu)
constexpr int kRowBuckets = kWdpRowBuckets;  // quarter-octave buckets of a task's row count (longest first)
constexpr int kSegs = 2 * kRowBuckets * 10;  // (family, class, rows bucket) segments of the sorted task list
constexpr int kShortInst = 4;                // DP queues for tasks below Ptrs::long_rows rows (each with its own streams, task list and direction arena)
constexpr int kLongInst = 12;               // ... and for the long tasks
constexpr int kQueues = kShortInst + kLongInst;   // queue ids: short 0 .. kShortInst - 1, long kShortInst .. kQueues - 1

enum Stage : int {
    ST_FREE = 0, ST_DONE, ST_WALK, ST_WALKING, ST_ZOMBIE_WALKING, ST_NEED_SEARCH, ST_WAIT_SEARCH, ST_NEED_POLISH, ST_NEED_CONS, ST_WAIT_CONS, ST_NEED_DP, ST_WAIT_DP
};
enum { U_RR = 0, U_TMP = 1, U_DIR0 = 2, U_DIR1 = 3 };       // unit strings of a chain
enum { S_RR = 0, S_DIR0 = 1, S_DIR1 = 2 };                  // score strings (node counts clamped to 255: only ==1 and <2 are ever tested)
enum { ERR_NONE = 0, ERR_WRAPCAP = 1, ERR_EMPTY_UNIT = 2, ERR_ACCEPTED_FULL = 3, ERR_TASKS_FULL = 4, ERR_STUCK = 5 };

// the per-repeat feature record (mTR.h:99-119), without the strings
struct Rec { int rep_start, rep_end, repeat_len, period, units, nm, nx, ni, nd, kmer, gain, mis, indel, pad[3]; };

struct Chain {                     // one k of one candidate: find_tandem_repeat_sub (handle_one_read.c:77-100)
    Rec rr, tmp;
    int stage, k, pass, found_last;
    float ratio0;
    int read, qs, qe;
    int dir_found[2], dir_period[2];
    int fatal;                     // ERR_* to raise when (and only if) the candidate commits
    int msg;                       // "You need to increse the value of WrapDPsize." lines to print when it commits
    int ring;                      // ring position of its candidate (absolute), for the cell accounting
    int aux_off;                   // CONSENSUS block (int32 offset into the aux pool of queue aux_q)
    int pending;                   // DP results this chain is still waiting for (the traceback kernels count it down)
    int aux_q;                     // queue whose aux pool holds it
    // Search DPs shared inside a candidate: different k (and the two walk directions) very often end in the SAME unit
    // string, and the DP of the same window against the same unit is the same DP.  lead[d] >= 0: direction d takes its two
    // results from chain lead[d] / 2, direction lead[d] % 2 of the same candidate instead of running them again.
    int lead[2];
    int wait_lead;                 // directions still waiting for their leader's results
    int search_done;               // the four search results are final (they stay valid: the revise passes use slot kResRevise)
    int no_share;                  // this chain runs its own search DPs (set by unshare_chain)
    int pad_[1];
};

struct Cand { int qs, qe, set, spec, min_k, n_k; long long cells; };   // set < 0: no k passed the maxFreq gate

// A Read is a SLOT: it holds one read of the resident batch at a time and takes the next unstarted one (Counters::next_read)
// as soon as its read has finished, so that every wave works on n_slots reads until the batch runs out.
struct ReadDesc { long long word_off, pos_off; int L, id; };   // one read of the resident batch (packed words, directional-index arrays, index in the batch); slots take them in array order

struct Read {
    long long word_off, pos_off;
    int L, cursor, head, n_ring;   // ring entries head .. head + n_ring - 1 (mod kRing), in candidate order
    int phase, n_accepted, candidates, id;   // phase 0: at work on read `id` of the batch, 1: free
    unsigned set_mask;             // chain sets in use
    unsigned zombie_mask;          // ... by dropped candidates whose walks have not finished yet
    long long cells_wasted;
    Cand ring[kRing];
};

struct Accepted { int read, seq; Rec rec; unsigned char unit[kUnitStride]; };   // insert_an_alignment, handle_one_read.c:156-176

struct MemoEntry { unsigned key, epoch; int next; unsigned short seen; unsigned char self1, next1; };   // scores: min(count, 254) + 1, 0 = not known yet

struct Counters {
    int n_polish, polish_head;
    int dp_pending;                // (task, penalty set) results outstanding in any queue
    int pad0;
    // Walks run beside everything else (a pathological walk takes tens of milliseconds: it must delay its own candidate,
    // not the group): sched_read pushes chains at walk_tail, the walk kernel instance launched after a scheduler pass pops
    // below the tail it saw at its start.  Both counters only grow.
    unsigned walk_tail, walk_head;
    unsigned walk_tail_big, walk_head_big;   // the same for the windows that need the 96 KB count table (their kernel fits one cta per SM)
    int walks_running, walks_done; // chains a walk kernel is working on right now / has finished so far
    int unfinished, error, error_read, n_accepted;
    int deferred, msgs, waves, progress;
    int next_read;                 // reads of the batch handed to slots so far
    int unshared;                  // chains taken out of the sharing by a rescue pass (should stay 0)
    unsigned long long cells_p16;      // part of `cells` run by the paired int16x2 kernels (both penalty sets in one register)
    unsigned long long shared_cells;   // cells of search DPs that were not run because a sibling chain ran the identical DP
    unsigned long long cells, slot_cells, spec_cells, jobs, candidates, tables, walks, table_positions, dir_bytes, tasks_total;
    // profile (clock64 ticks, thread 0 / lane 0 of the walking warps): table build, node list, walks per direction; walk steps; tasks per table layout
    unsigned long long prof_build, prof_list, prof_walk[2], prof_steps, prof_kind[3], prof_walk_tasks, prof_max_task;
    unsigned long long prof_probe_rounds, prof_memo_hits, prof_deep_steps, prof_fail_walks, prof_max_walk;
    // estimated K3 work per (family, fill class), smoothed over the uses of all queues: the fill kernels of every queue map
    // the classes to the SMs by these shares, so that concurrent launches agree on which SM runs which class
    unsigned long long class_share[20];
    unsigned long long prof_k3[10];   // K3 per family: fill ticks, traceback ticks, slot rows, slots, traceback rows (wdp.cu WdpTb)
};

// One queue of wrap-around DP tasks: the list as emitted, the list sorted by segment, the sort's work arrays, and the
// per-use reservations (task slots, direction-matrix bytes, consensus-histogram ints).  No wave waits for a DP: a wave
// (tick) emits the tasks of the chains that are ready into one free short queue (rows < Ptrs::long_rows) and one free
// long queue, the K3 kernels of a queue run on its own streams beside the following waves, and the chain that owns a
// task stays in WAIT_* until its `pending` count -- counted down by the warp that finishes a task's traceback -- reaches
// zero.  A 20 k-row DP takes milliseconds; the thousands of other reads at work must not wait for it.
struct QueueCtr { int n_tasks, deferred; unsigned long long dir_used, aux_used; };
struct DpQueue {
    WdpTask *tasks_in, *tasks;     // as emitted / sorted by (family, rows descending, class)
    int task_cap;
    int *hist, *seg_task, *seg_slot, *bucket_cursor;   // [kSegs (+1)]
    int *class_begin;              // [WDP_NCLASS + 1]: entries 0 .. 19: estimated work of fill class c of family f at 10 f + c (plan_tasks); last entry: total number of tasks
    int *slot_counter;             // [WDP_NCLASS] work-queue heads of the fill / traceback kernels
    int *aux;                      // consensus histograms
    long long aux_cap;             // in int32
    long long dir_cap;             // bytes of direction matrices
    QueueCtr *qc;
    int id;                        // index among the kQueues queues of the group
};

struct Ptrs {
    const uint32_t *packed;
    Read *reads;                   // the slots
    int n_reads;                   // number of slots
    const ReadDesc *descs;         // the reads of the batch
    int n_total;                   // ... and their number
    int *end, *w;                  // directional_index_end / _w of every read (Read::pos_off)
    Chain *chains;                 // [n_reads * kSets * kMaxK]
    unsigned char *units;          // [chain][4][kUnitStride]
    unsigned char *scores;         // [chain][3][kUnitStride]
    mtr_wdp_result *results;       // [chain][kResSlots]: search: direction d, penalty set s at 2d + s; revise: slot kResRevise
    int *polish_list;
    int *walk_ring;                // queue of chains in ST_WALK (Counters::walk_tail / walk_head, never reset)
    int *walk_ring_big;            // ... whose windows need the big shared-memory count table (same size)
    unsigned walk_ring_mask;
    int *aux_of[kQueues];          // aux pools by Chain::aux_q
    int long_rows;                 // tasks with at least this many rows go to a long queue
    Accepted *acc;
    int acc_cap;
    Counters *ctr;
    unsigned char *uf_scratch;     // per unit-finder cta: memos, tie lists, node list, unit / score strings
    long long uf_stride;
    unsigned char *uf_wide;        // WIDE count tables, one per unit-finder cta
    unsigned table_cap;            // slots of one WIDE table (power of two)
    unsigned compact_cap;          // slots of a COMPACT table (kCompactCap; tests lower it to reach the WIDE path)
    int direct_max_k;              // largest k that gets a DIRECT table (7)
    float min_match_ratio;
    int speculate;
    int share_search;              // identical search DPs of one candidate run once (MTR_ENGINE_SHARE=0: off)
    unsigned long long *stamps;    // MTR_TIMELINE: [wave][16] %globaltimer values written by the wave's kernels (nullptr: off)
    int stamp_waves;
};

// ---------------------------------------------------------------- records
MTR_DEV void rec_clear(Rec &r)                                  // clear_rr, fill_directional_index.c:40-60
{
    r.rep_start = r.rep_end = r.repeat_len = r.period = r.units = -1;
    r.nm = r.nx = r.ni = r.nd = -1; r.kmer = r.gain = r.mis = r.indel = -1;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
}
MTR_DEV float rec_ratio(const Rec &r) { return (float)r.nm / (float)(r.nm + r.nx + r.ni + r.nd); }   // e.g. wrap_around_DP.c:398

// wrap_around_DP_sub's record update (wrap_around_DP.c:337-350)
MTR_DEV void apply_dp(Rec &r, int qs, const mtr_wdp_result &d, int g, int m, int in)
{
    r.rep_start = qs + d.end_i + 1;
    r.rep_end = qs + d.max_i;
    r.repeat_len = d.max_i - d.end_i;
    r.units = d.n_scanned / r.period;
    r.nm = d.n_match; r.nx = d.n_mismatch; r.ni = d.n_ins; r.nd = d.n_del;
    r.gain = g; r.mis = m; r.indel = in;
}

MTR_DEV unsigned char *unit_ptr(const Ptrs &P, int chain, int slot) { return P.units + ((size_t)chain * 4 + slot) * kUnitStride; }
MTR_DEV unsigned char *score_ptr(const Ptrs &P, int chain, int slot) { return P.scores + ((size_t)chain * 3 + slot) * kUnitStride; }
MTR_DEV void copy_bytes(unsigned char *dst, const unsigned char *src, int n)      // warp-cooperative
{
    for (int i = lane(); i < n; i += NL) dst[i] = src[i];
    wsync();
}

// ---------------------------------------------------------------- packed reads
MTR_DEV int base_at(const uint32_t *rd, int i) { return (int)((rd[i >> 4] >> ((i & 15) * 2)) & 3u); }

// code of the k bases starting at i, first base most significant
MTR_DEV unsigned kmer_at(const uint32_t *rd, int i, int k)
{
#ifdef __CUDA_ARCH__
    const unsigned long long w = (unsigned long long)rd[i >> 4] | ((unsigned long long)rd[(i >> 4) + 1] << 32);
    unsigned v = (unsigned)(w >> ((i & 15) * 2));
    if (k < 16) v &= (1u << (2 * k)) - 1u;
    v = __brev(v);
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);     // bit-reversed -> 2-bit groups reversed
    return v >> (32 - 2 * k);
#else
    unsigned v = 0;
    for (int t = 0; t < k; t++) v = v * 4u + (unsigned)base_at(rd, i + t);
    return v;
#endif
}

// ---------------------------------------------------------------- cooperating threads
// A "cta" is the set of threads that work on one task together: a whole thread block (walk / polish kernels: four
// warps build the count table, then warp d walks direction d) or a single warp (the scheduler's inline gate, and every
// cta of the CPU twin).  sh: a few ints of memory all its threads see (shared memory on the GPU).
struct Cta { int tid, size, warp, nwarps; int *sh; bool block; };
MTR_DEV void cta_sync(const Cta &c) { if (c.block) block_sync(); else wsync(); }
MTR_DEV Cta cta_of_warp(int *sh) { Cta c; c.tid = lane(); c.size = NL; c.warp = 0; c.nwarps = 1; c.sh = sh; c.block = false; return c; }
MTR_DEV Cta cta_of_block(int *sh) { Cta c; c.tid = block_tid(); c.size = block_size(); c.warp = block_tid() / NL; c.nwarps = block_size() / NL; c.sh = sh; c.block = true; return c; }
MTR_DEV int cta_bcast(const Cta &c, int v, int slot)           // value of thread 0, to everybody
{
    cta_sync(c);
    if (c.tid == 0) c.sh[slot] = v;
    cta_sync(c);
    return c.sh[slot];
}

// ---------------------------------------------------------------- exact k-mer counts of one window
// init_inputString + generate_freqNode_return_list_maxNodes + freq_node (consensus.c:37-60,132-253).  The reference's own
// table layout (direct for k <= 6, hashing modulo a prime above) is unobservable.  Three layouts, chosen per task so that
// a unit-finder cta never needs more than kUfSmemBytes of shared memory (several ctas of several groups share an SM):
//   DIRECT   k <= 7: one 16-bit counter per k-mer code (4^7 x 2 B = 32 KB), in shared memory; no keys, no probing
//   COMPACT  small windows: open addressing in shared memory, keys (key + 1, 32 bit) and 16-bit counters apart, 6 B / slot
//   WIDE     everything else: open addressing, one 64-bit slot per node (key + 1 high, count low) in global memory (or,
//            for the scheduler's inline gate, in a warp's shared memory); the walks, which are chains of dependent
//            probes into the by then frozen table, go through a direct-mapped (node -> count) CACHE in shared memory
//            (an L2 round trip costs ~700 cycles, a shared-memory hit ~30)
enum { TB_WIDE = 0, TB_COMPACT = 1, TB_DIRECT = 2 };
constexpr int kUfSmemWords = 24576;                      // 96 KB of table / cache per unit-finder cta
constexpr int kUfSmemWordsSmall = 8192;                  // ... of the walk kernel for small windows (32 KB)
constexpr unsigned kCompactCapSmall = 4096;              // its COMPACT tables: 24 KB, windows up to 3070 positions
constexpr unsigned kCompactCap = 16384;                  // slots of a COMPACT table: 16384 * 6 B = 96 KB (windows up to 12286 positions)
constexpr unsigned kCacheSlots = 4096;                   // 64-bit entries of the probe cache
struct Table {
    unsigned long long *slots;      // WIDE
    unsigned *keys, *cnt;           // COMPACT (keys + counters) / DIRECT (counters)
    unsigned long long *cache;      // WIDE: probe cache, nullptr while the counts still change
    unsigned mask;
    int shift, kind;
};

#ifdef __CUDACC__
__host__ __device__
#endif
inline int compact_max_n(unsigned cap) { const unsigned m = cap / 4u * 3u; return (int)(m < 65000u ? m : 65000u) - 2; }

MTR_DEV Table table_wide(unsigned long long *mem, unsigned cap_max, int n)
{
    unsigned cap = 64;
    while (cap < 2u * (unsigned)(n + 2) && cap < cap_max) cap <<= 1;
    Table t;
    t.slots = mem; t.keys = nullptr; t.cnt = nullptr; t.cache = nullptr; t.mask = cap - 1u; t.shift = clz(cap) + 1; t.kind = TB_WIDE;
    return t;
}
MTR_DEV Table table_compact(unsigned *mem, unsigned cap_max, int n)
{
    unsigned cap = 64;
    while (cap < 2u * (unsigned)(n + 2) && cap < cap_max) cap <<= 1;
    Table t;
    t.slots = nullptr; t.keys = mem; t.cnt = mem + cap; t.cache = nullptr; t.mask = cap - 1u; t.shift = clz(cap) + 1; t.kind = TB_COMPACT;
    return t;
}
MTR_DEV Table table_direct(unsigned *mem, int k)
{
    Table t;
    t.slots = nullptr; t.keys = nullptr; t.cnt = mem; t.cache = nullptr; t.mask = (1u << (2 * k)) - 1u; t.shift = 0; t.kind = TB_DIRECT;
    return t;
}
MTR_DEV unsigned table_home(const Table &t, unsigned code) { return ((code + 0x9e3779b9u) * 2654435761u) >> t.shift; }
MTR_DEV void table_clear(const Table &t, const Cta &c)
{
    if (t.kind == TB_WIDE) {
        for (unsigned i = (unsigned)c.tid; i <= t.mask; i += (unsigned)c.size) t.slots[i] = 0ull;
    } else {
        if (t.kind == TB_COMPACT) for (unsigned i = (unsigned)c.tid; i <= t.mask; i += (unsigned)c.size) t.keys[i] = 0u;
        for (unsigned i = (unsigned)c.tid; i <= t.mask / 2u; i += (unsigned)c.size) t.cnt[i] = 0u;
    }
    cta_sync(c);
}
MTR_DEV int half_add(unsigned *cnt, unsigned h, unsigned delta)        // 16-bit counter h: returns the old value
{
    const int sh = (int)(h & 1u) * 16;
    return (int)((atomic_add(&cnt[h >> 1], delta << sh) >> sh) & 0xffffu);
}
MTR_DEV int table_insert(const Table &t, unsigned code)         // count after the insert
{
    if (t.kind == TB_DIRECT) return half_add(t.cnt, code, 1u) + 1;
    unsigned h = table_home(t, code);
    if (t.kind == TB_COMPACT) {
        const unsigned key = code + 1u;
        for (;;) {
            const unsigned old = atomic_cas(&t.keys[h], 0u, key);
            if (old == 0u || old == key) return half_add(t.cnt, h, 1u) + 1;
            h = (h + 1u) & t.mask;
        }
    }
    const unsigned long long key = ((unsigned long long)code + 1ull) << 32;
    for (;;) {
        const unsigned long long old = atomic_cas(&t.slots[h], 0ull, key | 1ull);
        if (old == 0ull) return 1;
        if ((old & 0xffffffff00000000ull) == key) return (int)(unsigned)(atomic_add(&t.slots[h], 1ull) + 1ull);
        h = (h + 1u) & t.mask;
    }
}
MTR_DEV int table_find(const Table &t, unsigned code)           // slot or -1
{
    if (t.kind == TB_DIRECT) return code <= t.mask ? (int)code : -1;
    unsigned h = table_home(t, code);
    if (t.kind == TB_COMPACT) {
        const unsigned key = code + 1u;
        for (;;) {
            const unsigned k = ldv(&t.keys[h]);
            if (k == key) return (int)h;
            if (k == 0u) return -1;
            h = (h + 1u) & t.mask;
        }
    }
    const unsigned long long key = ((unsigned long long)code + 1ull) << 32;
    for (;;) {
        const unsigned long long sl = ldv(&t.slots[h]);
        if ((sl & 0xffffffff00000000ull) == key) return (int)h;
        if (sl == 0ull) return -1;
        h = (h + 1u) & t.mask;
    }
}
MTR_DEV int table_count_at(const Table &t, int h)
{
    if (t.kind == TB_WIDE) return (int)(unsigned)ldv(&t.slots[h]);
    return (int)((ldv(&t.cnt[h >> 1]) >> ((h & 1) * 16)) & 0xffffu);
}
MTR_DEV void table_decrement_at(const Table &t, int h)          // the count is >= 6 here: no borrow into the neighbour / the key
{
    if (t.kind == TB_WIDE) atomic_add(&t.slots[h], ~0ull);
    else atomic_add(&t.cnt[h >> 1], 0u - (1u << ((h & 1) * 16)));
}
MTR_DEV int table_count(const Table &t, unsigned code)          // freq_node
{
    if (t.cache) {
        // frozen WIDE table: (node + 1, count) pairs in shared memory, direct-mapped, absent nodes cached as count 0
        const unsigned cs = ((code + 0x9e3779b9u) * 2654435761u) >> 20 & (kCacheSlots - 1u);
        const unsigned long long e = ldv(&t.cache[cs]);
        if ((unsigned)(e >> 32) == code + 1u) return (int)(unsigned)e;
        const int h = table_find(t, code);
        const int n = h < 0 ? 0 : table_count_at(t, h);
        t.cache[cs] = ((unsigned long long)(code + 1u) << 32) | (unsigned)n;
        return n;
    }
    const int h = table_find(t, code);
    return h < 0 ? 0 : table_count_at(t, h);
}
// after the counts are final: route the probes of a WIDE table through the cache in `smem` (kCacheSlots entries)
MTR_DEV void table_freeze(Table &t, unsigned *smem, const Cta &c)
{
    if (t.kind != TB_WIDE || !smem) return;
    unsigned long long *cache = (unsigned long long *)smem;
    for (unsigned i = (unsigned)c.tid; i < kCacheSlots; i += (unsigned)c.size) cache[i] = 0ull;
    cta_sync(c);
    t.cache = cache;
}

// codes of the window [qs, qe]: k-mer codes below min(qe, L-k+1), the raw base above (Q7); index >= L reads as 0
// (the reference's value there depends on its whole call history, H4b)
struct Window { const uint32_t *rd; int L, k, qs, qe, coded_end; };
MTR_DEV Window window_make(const uint32_t *rd, int L, int k, int qs, int qe)
{
    Window w;
    w.rd = rd; w.L = L; w.k = k; w.qs = qs; w.qe = qe;
    w.coded_end = qe < L - k + 1 ? qe : L - k + 1;
    return w;
}
MTR_DEV unsigned window_code(const Window &w, int i)
{
    if (i < w.coded_end) return kmer_at(w.rd, i, w.k);
    return i < w.L ? (unsigned)base_at(w.rd, i) : 0u;
}

// builds the table of a window with every thread of the cta, returns maxFreq to all of them (counts only grow while
// building: the running maximum is the final one)
MTR_DEV int table_build(const Table &t, const Window &w, const Cta &c)
{
    table_clear(t, c);
    if (c.tid == 0) c.sh[0] = -1;
    cta_sync(c);
    int maxf = -1;
    for (int i = w.qs + c.tid; i <= w.qe; i += c.size) {
        const int n = table_insert(t, window_code(w, i));
        maxf = n > maxf ? n : maxf;
    }
    maxf = wmax(maxf);
    if (lane() == 0) atomic_max(&c.sh[0], maxf);
    cta_sync(c);
    maxf = c.sh[0];
    cta_sync(c);
    return maxf;
}

// generate_freqNode_return_list_maxNodes (:132-229): nodes whose CURRENT count equals maxFreq, in order of first
// occurrence, at most `cap`; a listed node loses one count (Q8).  NL positions per step; duplicates inside a step are
// resolved with match_any so that the order is exactly the sequential one.
MTR_DEV int table_list_max(const Table &t, const Window &w, int maxf, int *nodes, int cap, int n_max)
{
    // n_max: number of distinct nodes whose count is maxFreq (table_count_max): once all of them are listed the rest of
    // the window cannot add anything, and the nodes of a repeat show up within its first copies
    if (n_max < cap) cap = n_max;
    int nn = 0;
    for (int p0 = w.qs; p0 <= w.qe && nn < cap; p0 += NL) {
        const int p = p0 + lane();
        unsigned code = 0x80000000u | (unsigned)lane();         // unique dummy for idle lanes
        int h = -1;
        bool ismax = false;
        if (p <= w.qe) {
            const unsigned c = window_code(w, p);
            h = table_find(t, c);
            ismax = h >= 0 && table_count_at(t, h) == maxf;
            if (ismax) code = c;
        }
        const unsigned grp = match_any(code);
        const bool lead = ismax && (ffs(grp) - 1) == lane();
        const unsigned lst = wballot(lead);
        const int at = nn + popc(lst & lanemask_lt());
        if (lead && at < cap) { nodes[at] = (int)code; table_decrement_at(t, h); }
        nn += popc(lst);
        if (nn > cap) nn = cap;
        wsync();
    }
    return nn;
}

// number of table entries whose count equals maxf (every thread of the cta gets the result); the table is final
MTR_DEV int table_count_max(const Table &t, int maxf, const Cta &c)
{
    if (c.tid == 0) c.sh[0] = 0;
    cta_sync(c);
    int mine = 0;
    if (t.kind == TB_WIDE) { for (unsigned i = (unsigned)c.tid; i <= t.mask; i += (unsigned)c.size) mine += (int)(unsigned)t.slots[i] == maxf && t.slots[i] != 0ull; }
    else for (unsigned i = (unsigned)c.tid; i <= t.mask / 2u; i += (unsigned)c.size) { const unsigned v = t.cnt[i]; mine += (int)(v & 0xffffu) == maxf; mine += (int)(v >> 16) == maxf; }
    mine = wsum(mine);
    if (lane() == 0 && mine) atomic_add(&c.sh[0], mine);
    cta_sync(c);
    const int r = c.sh[0];
    cta_sync(c);
    return r;
}

// ---------------------------------------------------------------- greedy de Bruijn walk (consensus.c:269-505)
// From step 10 on the look-ahead depth is constant (k), so the next node is a pure function of the current node and of
// the (now frozen) count table.  The memo caches that function across the <= 100 start nodes of one (window, k,
// direction) and marks the nodes a walk has visited: coming back to a visited node means the walk is caught in a cycle
// that does not contain its start node, i.e. it can only run out its step limit -- "no loop" is returned at once.
// Both shortcuts leave every observable result unchanged; a full memo simply stops learning.
struct Memo { MemoEntry *tab; unsigned epoch; int used, serial; };

MTR_DEV void memo_reset(Memo &m)
{
    m.epoch++;
    if (m.epoch == 0u) {                                       // (cannot happen within one task; kept for safety)
        for (int i = lane(); i < kMemoSlots; i += NL) m.tab[i].epoch = 0u;
        wsync();
        m.epoch = 1u;
    }
    m.used = 0; m.serial = 0;
}
// slot of `node`: a window of four consecutive slots starting at its home; *e receives the entry and *fresh = false if
// the node is there, else *fresh = true and the slot returned is the one to (over)write: the first free one of the
// window, or -- all four taken by other nodes -- the home slot itself.  Forgetting an entry only costs time: its
// look-ahead is redone, and a walk caught in a cycle goes round until it meets a node it still remembers.  No lane
// writes here: every lane sees the same table.
MTR_DEV int memo_find(const Memo &m, unsigned node, MemoEntry *e, bool *fresh)
{
    const unsigned home = (((node + 0x9e3779b9u) * 2654435761u) >> 16) & (unsigned)(kMemoSlots - 4);   // windows do not wrap
    int free_slot = -1;
    for (unsigned w = 0; w < 4u; w++) {
        const unsigned h = home + w;
        *e = m.tab[h];
        if (e->epoch != m.epoch) { if (free_slot < 0) free_slot = (int)h; continue; }
        if (e->key == node) { *fresh = false; return (int)h; }
    }
    *fresh = true;
    return free_slot >= 0 ? free_slot : (int)home;
}
MTR_DEV int score_byte(int count) { return count > 254 ? 254 : count; }

// A tie list: the first kTiesNear entries live in shared memory (`near`), the rest in global memory (`far`).  A list is
// written by one look-ahead level and read by the next; almost all lists are a handful of entries long, and a store to
// global memory followed by a load of the same address costs an L2 round trip.
constexpr int kTiesNear = 64;
struct TieList { int *near, *far; };
constexpr int kUfDynSmem = kUfSmemWords * 4 + 4 * kTiesNear * 4 + 2 * kMemoSlots * 16;   // dynamic shared memory of a unit-finder cta
constexpr int kUfDynSmemSmall = kUfSmemWordsSmall * 4 + 4 * kTiesNear * 4 + 2 * kMemoSlots * 16;
MTR_DEV int tie_get(const TieList &t, int i) { return i < kTiesNear ? t.near[i] : t.far[i]; }
MTR_DEV void tie_put(const TieList &t, int i, int v) { if (i < kTiesNear) t.near[i] = v; else t.far[i] = v; }

// One walk.  Returns the period (0 = no loop).  ustr / uscore: the unit and its node counts (clamped to 254) in walk
// order (the backward walk's strings are reversed by the caller).  Per step: every lane reads the memo entry of the
// node, the warp computes what the entry does not know yet (look-ahead, counts), lane 0 writes entry and strings, one
// fence on each side of the write.  near-lists: 4 * kTiesNear ints of shared memory per cta (two lists per direction).
MTR_DEV int walk(const Table &tb, Memo &memo, int qs, int qe, unsigned start, int k, bool backward,
                 unsigned char *ustr, unsigned char *uscore, TieList ties, TieList fresh, Counters *prof)
{
    unsigned node = start;
    long long p_steps = 0, p_rounds = 0, p_hits = 0, p_deep = 0;
    const long long p_t0 = clock_now();
    int limit = (qe - qs) / 5;                                  // MIN_NUM_FREQ_UNIT
    if (limit > kMaxPeriod) limit = kMaxPeriod;
    const int serial = ++memo.serial;                           // 1, 2, ... within one (window, k, direction)
    int period = 0;
    for (int l = 0; l < limit; l++) {
        MemoEntry e;
        e.key = node; e.epoch = memo.epoch; e.next = -1; e.seen = 0; e.self1 = 0; e.next1 = 0;
        int me = -1;
        bool is_fresh = true;
        if (l >= 10) {
            MemoEntry got;
            me = memo_find(memo, node, &got, &is_fresh);
            if (me >= 0 && !is_fresh) {
                e = got;
                if (e.seen == serial) { period = 0; break; }   // cycle without the start node
            }
        }
        int self_sc = e.self1 ? e.self1 - 1 : -1, next_sc = e.next1 ? e.next1 - 1 : -1;
        if (!backward && self_sc < 0) self_sc = score_byte(table_count(tb, node));
        int next = e.next;
        p_steps++;
        if (next >= 0) p_hits++;
        if (next < 0) {
            const int depth = l < 10 ? 1 : k;
            int nties = 1, pick = 0, m;
            wsync();
            if (lane() == 0) tie_put(ties, 0, 0);
            wsync();
            TieList cur = ties, nxt = fresh;
            for (m = 1; m <= depth; m++) {
                const int ncand = 4 * nties;
                p_rounds += (ncand + NL - 1) / NL;
                if (m == 4) p_deep++;
                const unsigned keep = (1u << (2 * (k - m))) - 1u;
                // one pass in candidate order: the running maximum, and the list of the extensions that reach it (a new
                // maximum restarts the list; first wins, at most 1024) -- the reference's own loop, 32 candidates a time
                int best = -1, nf = 0;
                for (int c0 = 0; c0 < ncand; c0 += NL) {
                    const int ci = c0 + lane();
                    int c = -2, digits = 0;
                    if (ci < ncand) {
                        const int t = tie_get(cur, ci >> 2), b = ci & 3;
                        digits = backward ? (b << (2 * (m - 1))) + t : 4 * t + b;
                        const unsigned cand = backward ? ((unsigned)digits << (2 * (k - m))) + (node >> (2 * m))
                                                       : ((node & keep) << (2 * m)) + (unsigned)digits;
                        c = table_count(tb, cand);
                    }
                    const int cmax = wmax(c);
                    if (cmax > best) { best = cmax; nf = 0; }
                    const unsigned eq = wballot(c == best);
                    if (eq) {
                        if (nf == 0) pick = bcast(digits, ffs(eq) - 1);
                        const int at = nf + popc(eq & lanemask_lt());
                        if (c == best && at < kMaxTies) tie_put(nxt, at, digits);
                        nf += popc(eq);
                        if (nf > kMaxTies) nf = kMaxTies;
                    }
                }
                wsync();
                if (backward ? nf <= 1 : nf == 1) break;
                const TieList sw = cur; cur = nxt; nxt = sw;
                nties = nf;
            }
            // m == depth + 1 when the ties were never resolved: the appended base is then pick / 4^depth == 0 ('A', :336)
            next = backward ? (int)(((unsigned)(pick & 3) << (2 * (k - 1))) + (node >> 2))
                            : (int)(((node & ((1u << (2 * (k - 1))) - 1u)) << 2) + ((unsigned)pick >> (2 * (m - 1))));
        }
        if (backward && next_sc < 0) next_sc = score_byte(table_count(tb, (unsigned)next));
        const unsigned strnode = backward ? (unsigned)next : node;
        wsync();                                                // every lane has read the entry
        if (lane() == 0) {
            ustr[l] = (unsigned char)(strnode >> (2 * (k - 1)));
            uscore[l] = (unsigned char)(backward ? next_sc : self_sc);
            if (me >= 0) {
                e.seen = (unsigned short)serial;
                e.next = next;                                  // (only entries of steps >= 10 exist: full look-ahead depth)
                if (!backward) e.self1 = (unsigned char)(self_sc + 1); else e.next1 = (unsigned char)(next_sc + 1);
                memo.tab[me] = e;
            }
        }
        wsync();
        node = (unsigned)next;
        if (node == start) { period = l + 1; if (kMaxPeriod <= period) period = 0; break; }
    }
    wsync();
    if (prof && lane() == 0) {
        atomic_add(&prof->prof_steps, (unsigned long long)p_steps); atomic_add(&prof->prof_probe_rounds, (unsigned long long)p_rounds);
        atomic_add(&prof->prof_memo_hits, (unsigned long long)p_hits); atomic_add(&prof->prof_deep_steps, (unsigned long long)p_deep);
        if (period == 0) atomic_add(&prof->prof_fail_walks, 1ull);
        const unsigned long long dt = (unsigned long long)(clock_now() - p_t0);
        unsigned long long old = prof->prof_max_walk;
        while (dt > old) { const unsigned long long seen = atomic_cas(&prof->prof_max_walk, old, dt); if (seen == old) break; old = seen; }
    }
    return period;
}

// ---------------------------------------------------------------- unit-finder scratch of one cta (global memory)
// Per direction (= walking warp): the far parts of the two tie lists, unit / score strings.  Shared: the list of start nodes, the
// polish output and the WIDE count table.
struct Scratch {
    int *ties[2], *fresh[2];       // far parts of the tie lists
    unsigned char *ustr[2], *uscore[2];
    unsigned *epoch[2];            // running memo epochs (survive from task to task and from wave to wave)
    int *nodes;
    unsigned char *revised;
    unsigned long long *wide;      // this cta's WIDE table
};
constexpr long long kScratchFixed = 4LL * kMaxTies * 4 + 4LL * kUnitStride + 128 * 4 + kUnitStride + 256;
MTR_DEV Scratch scratch_of(const Ptrs &P, int cta)
{
    unsigned char *b = P.uf_scratch + (size_t)cta * (size_t)P.uf_stride;
    Scratch s;
    for (int d = 0; d < 2; d++) { s.ties[d] = (int *)b; b += (size_t)kMaxTies * 4; s.fresh[d] = (int *)b; b += (size_t)kMaxTies * 4; }
    for (int d = 0; d < 2; d++) { s.ustr[d] = b; b += kUnitStride; s.uscore[d] = b; b += kUnitStride; }
    s.nodes = (int *)b; s.epoch[0] = (unsigned *)b + 120; s.epoch[1] = (unsigned *)b + 121; b += 128 * 4;
    s.revised = b; b += kUnitStride + 256;
    s.wide = (unsigned long long *)(P.uf_wide + (size_t)cta * (size_t)P.table_cap * 8);
    return s;
}

// the count table of a window of n positions and k-mer length k for this cta (smem: kUfSmemWords words)
MTR_DEV Table table_for(const Scratch &S, const Ptrs &P, unsigned *smem, int n, int k)
{
    if (k <= P.direct_max_k && n <= 65000) return table_direct(smem, k);
    if (n <= compact_max_n(P.compact_cap)) return table_compact(smem, P.compact_cap, n);
    return table_wide(S.wide, P.table_cap, n);
}

// ---------------------------------------------------------------- search_De_Bruijn_graph up to wrap_around_DP (consensus.c:507-549)
// One chain in stage ST_WALK, one cta: counts (all threads), maxFreq gate, list of maximum-frequency nodes (warp 0), then
// the forward walks over that list until the first loop on warp 0 and the backward walks likewise on warp 1 (both on
// the same warp, one after the other, when the cta has only one).
// counts, gate, node list and walks of one window on one cta.  Results: c.sh[8] = maxFreq, c.sh[2 + d] = direction d found
// a loop, c.sh[4 + d] = its period, unit / score strings in S.ustr[d] / S.uscore[d] in WALK order (the caller reverses
// the backward one, :458-470).  walks_ctr: optional counter of walks run.
MTR_DEV void unit_walks(Table tb, const Window &win, const Scratch &S, const Cta &c, unsigned *smem, int *near, MemoEntry *memos, Counters *ctr)
{
    const long long t0 = clock_now();
    const int maxf = table_build(tb, win, c);
    const long long t1 = clock_now();
    if (c.tid == 0) { c.sh[2] = 0; c.sh[3] = 0; c.sh[4] = 0; c.sh[5] = 0; c.sh[6] = 0; c.sh[8] = maxf; }
    cta_sync(c);
    long long t2 = t1;
    if (5 < maxf) {                                             // MIN_NUM_FREQ_UNIT < maxFreq, :532
        const int n_max = table_count_max(tb, maxf, c);
        if (c.warp == 0) {
            const int nn = table_list_max(tb, win, maxf, S.nodes, 100, n_max);
            if (lane() == 0) c.sh[6] = nn;
        }
        cta_sync(c);
        const int nn = c.sh[6];
        table_freeze(tb, smem, c);
        t2 = clock_now();
        for (int d = 0; d < 2; d++) {
            if (c.warp != d % c.nwarps) continue;
            const long long tw0 = clock_now();
            Memo memo;
            memo.tab = memos + (size_t)d * kMemoSlots; memo.epoch = *S.epoch[d]; memo.used = 0; memo.serial = 0;
            memo_reset(memo);
            int nwalks = 0;
            for (int i = 0; i < nn; i++) {
                nwalks++;
                TieList ta, tf;
                ta.near = near + (2 * d) * kTiesNear; ta.far = S.ties[d]; tf.near = near + (2 * d + 1) * kTiesNear; tf.far = S.fresh[d];
                const int p = walk(tb, memo, win.qs, win.qe, (unsigned)S.nodes[i], win.k, d == 1, S.ustr[d], S.uscore[d], ta, tf, ctr);
                if (p == 0) continue;
                if (lane() == 0) { c.sh[2 + d] = 1; c.sh[4 + d] = p; }
                break;
            }
            ENG_LANE0(*S.epoch[d] = memo.epoch; if (ctr) { atomic_add(&ctr->walks, (unsigned long long)nwalks); atomic_add(&ctr->prof_walk[d], (unsigned long long)(clock_now() - tw0)); });
        }
    }
    cta_sync(c);
    if (ctr && c.tid == 0) {
        const long long t3 = clock_now();
        atomic_add(&ctr->prof_build, (unsigned long long)(t1 - t0)); atomic_add(&ctr->prof_list, (unsigned long long)(t2 - t1));
        atomic_add(&ctr->prof_kind[tb.kind], 1ull); atomic_add(&ctr->prof_walk_tasks, 5 < maxf ? 1ull : 0ull);
        unsigned long long old = ctr->prof_max_task;
        const unsigned long long mine = (unsigned long long)(t3 - t0);
        while (mine > old) { const unsigned long long seen = atomic_cas(&ctr->prof_max_task, old, mine); if (seen == old) break; old = seen; }
    }
}

MTR_DEV void walk_chain(const Ptrs &P, int chain, const Scratch &S, const Cta &c, unsigned *smem, int *near, MemoEntry *memos)
{
    Chain &ch = P.chains[chain];
    // A set freed and taken again inside one sched_read call leaves two entries for the same chain in the walk list
    // (the first one from the dropped candidate): whoever claims the chain first does the work, the other one leaves.
    int mine = 0;
    if (c.tid == 0) { mine = atomic_cas(&ch.stage, (int)ST_WALK, (int)ST_WALKING) == (int)ST_WALK; if (mine) atomic_add(&P.ctr->walks_running, 1); }
    if (!cta_bcast(c, mine, 1)) return;
    fence();                                                   // the scheduler wrote the chain's fields before it set ST_WALK
    const Read &rs = P.reads[ch.read];
    const int k = ch.k, qs = ch.qs, qe = ch.qe;
    const Window win = window_make(P.packed + rs.word_off, rs.L, k, qs, qe);
    const Table tb = table_for(S, P, smem, qe - qs + 1, k);
    if (c.tid == 0) { atomic_add(&P.ctr->tables, 1ull); atomic_add(&P.ctr->table_positions, (unsigned long long)(qe - qs + 1)); }
    unit_walks(tb, win, S, c, smem, near, memos, P.ctr);
    for (int d = 0; d < 2; d++) {
        if (c.warp != d % c.nwarps || !c.sh[2 + d]) continue;
        const int p = c.sh[4 + d];
        unsigned char *du = unit_ptr(P, chain, U_DIR0 + d), *ds = score_ptr(P, chain, S_DIR0 + d);
        for (int x = lane(); x < p; x += NL) {
            const int sx = d == 1 ? p - 1 - x : x;             // the backward walk is reversed (:458-470)
            du[x] = S.ustr[d][sx]; ds[x] = S.uscore[d][sx];
        }
    }
    cta_sync(c);
    if (c.tid == 0) {
        // foundLoop of the LAST walk attempted (Q4): the backward pass tries the start nodes until its first loop, so its
        // last attempt found one iff the pass did
        ch.found_last = c.sh[3];
        ch.dir_found[0] = c.sh[2]; ch.dir_found[1] = c.sh[3];
        ch.dir_period[0] = c.sh[4]; ch.dir_period[1] = c.sh[5];
        if (!(c.sh[2] || c.sh[3])) rec_clear(ch.rr);           // nothing found: find_tandem_repeat_sub clears (:86-88)
        // the walk kernel of a wave runs beside the wave's emission / DP kernels: publish the stage last, behind a fence
        fence();
        const int done = (c.sh[2] || c.sh[3]) ? (int)ST_NEED_SEARCH : (int)ST_DONE;
        if (atomic_cas(&ch.stage, (int)ST_WALKING, done) != (int)ST_WALKING) atomic_exch(&ch.stage, (int)ST_DONE);   // dropped meanwhile
        atomic_add(&P.ctr->walks_done, 1);
        atomic_add(&P.ctr->walks_running, -1);
    }
    cta_sync(c);
}

// ---------------------------------------------------------------- polish_repeat (consensus.c:584-704)
MTR_DEV int align_score(const Table &tb, int start, int k, int node, int period, const unsigned char *unit)
{
    int sum = 0;
    for (int j = start; 0 <= j && start - k < j; j--) {
        node = unit[j % period] * (1 << (2 * (k - 1))) + node / 4;
        sum += table_count(tb, (unsigned)node);
    }
    return sum;
}
MTR_DEV bool suspicious(const unsigned char *score, int nscore, int kmer, int j)
{
    int c = 0;
    for (int i = 0; i < kmer - 1 && 0 <= j - i; i++) {
        const int sc = (j - i) < nscore ? score[j - i] : -1;
        if (sc < 2) c++;
    }
    return (kmer - 1) * 0.8 < (double)c;
}

// polish rr of a chain in place: the table build uses the whole cta, the polish itself is warp-uniform scalar code on
// warp 0
MTR_DEV void polish_rr(const Ptrs &P, int chain, const Scratch &S, const Cta &c, unsigned *smem)
{
    Chain &ch = P.chains[chain];
    const Read &rs = P.reads[ch.read];
    const int k = ch.rr.kmer, period = ch.rr.period;
    if (period <= k) return;
    const Window win = window_make(P.packed + rs.word_off, rs.L, k, ch.rr.rep_start, ch.rr.rep_end);
    Table tb = table_for(S, P, smem, ch.rr.rep_end - ch.rr.rep_start + 1, k);
    table_build(tb, win, c);
    table_freeze(tb, smem, c);
    if (c.tid == 0) { atomic_add(&P.ctr->tables, 1ull); atomic_add(&P.ctr->table_positions, (unsigned long long)(ch.rr.rep_end - ch.rr.rep_start + 1)); }
    if (c.warp != 0) return;
    unsigned char *unit = unit_ptr(P, chain, U_RR);
    const unsigned char *score = score_ptr(P, chain, S_RR);
    const int nscore = period;                                  // the score string of the walk that produced the unit
    const int p4 = 1 << (2 * (k - 1));
    unsigned char *revised = S.revised;                         // filled from the back, kMaxPeriod entries
    int jr = kMaxPeriod - 1;
    int best = 0;
    for (int i = 0; i < k; i++) best = unit[i] * (1 << (2 * (k - 1 - i))) + best;
    for (int j = period - 1; 0 <= j;) {
        const int ref = unit[j] * p4 + best / 4;
        int best_freq = table_count(tb, (unsigned)ref);
        best = ref;
        const int sc = j < nscore ? score[j] : -1;
        unsigned char out;
        if (sc == 1 && suspicious(score, nscore, k, j)) {
            for (int l = 0; l < 4; l++) {
                const int alt = (ref + (l - unit[j]) * p4) % (4 * p4);
                const int f = table_count(tb, (unsigned)alt);
                if (best_freq < f) { best_freq = f; best = alt; }
            }
            if (best == ref) {
                out = unit[j]; j--;
            } else {
                const int s_del = align_score(tb, j, k, best, period, unit);
                const int s_sub = align_score(tb, j - 1, k, best, period, unit);
                int s_ins = -1;
                if (j >= 1 && best / p4 == unit[(j - 1) % period]) s_ins = align_score(tb, j - 2, k, best, period, unit);
                out = (unsigned char)(best / p4);
                int mx = s_del > s_sub ? s_del : s_sub;
                if (s_ins > mx) mx = s_ins;
                if (mx == s_del) {} else if (mx == s_sub) j -= 1; else j -= 2;
            }
        } else {
            out = unit[j]; j--;
        }
        if (lane() == 0) revised[jr] = out;
        jr--;
        if (jr < 0) return;                                     // "fails to revise": record unchanged
    }
    wsync();
    const int np = (kMaxPeriod - 1) - jr;
    copy_bytes(unit, revised + jr + 1, np);
    ENG_LANE0(ch.rr.period = np);
}

// ---------------------------------------------------------------- min_missing (consensus.c:714-820)
// each row of min_missing_bases[10][10][20] starts at 1 and steps by 0 or 1: stored as 19 step bits
MTR_CONST unsigned kMissingSteps[10][10] = {
    {0x21127,0x42227,0x08447,0x1084b,0x0108b,0x08113,0x40423,0x04043,0x00205,0x00041},
    {0x21127,0x04227,0x0844b,0x2088b,0x0210b,0x08213,0x00823,0x04085,0x00409,0x00081},
    {0x42227,0x0444b,0x1084b,0x4108b,0x04113,0x20413,0x01023,0x10085,0x00809,0x00101},
    {0x0422b,0x0844b,0x2088b,0x0208b,0x08213,0x20423,0x02045,0x20105,0x01009,0x00201},
    {0x0444b,0x1084b,0x4108b,0x04113,0x10213,0x00823,0x04045,0x40205,0x02011,0x00402},
    {0x1084b,0x2108b,0x02113,0x08213,0x20423,0x01045,0x08085,0x00409,0x08021,0x01002},
    {0x1088b,0x41113,0x04113,0x10423,0x40825,0x02085,0x10109,0x00809,0x10021,0x04002},
    {0x42113,0x04213,0x10423,0x40845,0x02085,0x08109,0x00409,0x02011,0x00081,0x40004},
    {0x04225,0x10425,0x40845,0x02085,0x08109,0x40209,0x01011,0x10041,0x00202,0x00010},
    {0x20845,0x01085,0x04089,0x08209,0x40411,0x01021,0x08041,0x00102,0x01004,0x00040},
};
MTR_DEV int min_missing(int period, double error, int coverage)
{
    const int plim[9] = {200, 150, 100, 75, 50, 30, 20, 10, 5};
    const double elim[9] = {0.25, 0.225, 0.2, 0.175, 0.15, 0.125, 0.1, 0.075, 0.05};
    int i = 9, j = 9;
    for (int t = 0; t < 9; t++) if (period > plim[t]) { i = t; break; }
    for (int t = 0; t < 9; t++) if (error > elim[t]) { j = t; break; }
    const int k = coverage <= 1 ? 0 : (coverage >= 20 ? 19 : coverage - 1);
    return 1 + popc(kMissingSteps[i][j] & ((1u << k) - 1u));
}

// majority vote of revise_representative_unit_sub (consensus.c:964-1013) from the two histograms of the CONSENSUS
// traceback: per unit column at most one consensus base (none if the gap wins) then at most one inserted base.
// Columns are voted NL at a time; the output order is the column order (ballot compaction).  Writes at most
// kUnitStride bases; the returned period may exceed that (the caller discards periods >= MAX_PERIOD).
MTR_DEV int vote_unit(const Rec &r, const int *cons, const int *miss, unsigned char *out)
{
    const int ulen = r.period;
    const int coverage = r.repeat_len / r.period;
    const bool ins_ok = 5 <= coverage && coverage <= 20;
    int need = 0;
    if (ins_ok) {
        const double mismatch_ratio = (double)(r.nx + r.ni + r.nd) / r.repeat_len;
        need = min_missing(r.period, mismatch_ratio, coverage);
    }
    int n = 0;
    for (int j0 = 1; j0 <= ulen; j0 += NL) {
        const int j = j0 + lane();
        int cb = -1, ib = -1;                                   // consensus base (0..3) / inserted base, -1 = none
        if (j <= ulen) {
            int mv = -1, mb = -1;
            for (int q = 0; q < 5; q++) if (mv < cons[j * 5 + q]) { mv = cons[j * 5 + q]; mb = q; }
            if (mb < 4) cb = mb;
            mv = -1; int mm = -1;
            for (int q = 0; q < 4; q++) if (mv < miss[j * 4 + q]) { mv = miss[j * 4 + q]; mm = q; }
            if (ins_ok && need <= mv && 0 <= mm && mm <= 3) ib = mm;
        }
        const unsigned bc = wballot(cb >= 0), bi = wballot(ib >= 0);
        const unsigned lt = lanemask_lt();
        int at = n + popc(bc & lt) + popc(bi & lt);
        if (cb >= 0) { if (at < kUnitStride) out[at] = (unsigned char)cb; at++; }
        if (ib >= 0) { if (at < kUnitStride) out[at] = (unsigned char)ib; }
        n += popc(bc) + popc(bi);
    }
    wsync();
    return n;
}

// ---------------------------------------------------------------- chain state machine: consume DP results
MTR_CONST int kSearchParams[2][3] = {{1, 1, 3}, {1, 3, 1}};    // wrap_around_DP.c:395,405
MTR_CONST int kReviseParams[2][3] = {{5, 1, 1}, {1, 1, 3}};    // consensus.c:1062,1076

MTR_DEV void start_revise_pass(const Ptrs &P, int chain)      // revise_representative_unit_sub's input: a copy of rr
{
    Chain &ch = P.chains[chain];
    copy_bytes(unit_ptr(P, chain, U_TMP), unit_ptr(P, chain, U_RR), ch.rr.period < kUnitStride ? ch.rr.period : kUnitStride);
    wsync();
    if (lane() == 0) {
        ch.tmp = ch.rr;
        ch.tmp.gain = kReviseParams[ch.pass][0]; ch.tmp.mis = kReviseParams[ch.pass][1]; ch.tmp.indel = kReviseParams[ch.pass][2];
        ch.stage = ST_NEED_CONS;
    }
    wsync();
}

MTR_DEV void next_pass_or_done(const Ptrs &P, int chain)
{
    Chain &ch = P.chains[chain];
    if (ch.pass == 0) {
        ENG_LANE0(ch.pass = 1);
        wsync();
        start_revise_pass(P, chain);
    } else {
        ENG_LANE0(ch.stage = ST_DONE);
        wsync();
    }
}

// a chain in WAIT_* whose DP results have all arrived (pending == 0) and are in P.results
MTR_DEV bool chain_ready(const Ptrs &P, int chain)
{
    const Chain &ch = P.chains[chain];
    const int st = ch.stage;
    return (st == ST_WAIT_SEARCH || st == ST_WAIT_CONS || st == ST_WAIT_DP) && ldv(&ch.pending) == 0 && ldv(&ch.wait_lead) == 0;
}
// one thread: a chain that waits for a leader whose results are final takes them itself (the leader hands them out when
// it advances; a follower that registers around that moment must not depend on having been seen)
MTR_DEV void take_shared(const Ptrs &P, int chain)
{
    Chain &ch = P.chains[chain];
    if (ldv(&ch.stage) != (int)ST_WAIT_SEARCH || ldv(&ch.wait_lead) <= 0) return;
    for (int d = 0; d < 2; d++) {
        const int ld = ldv(&ch.lead[d]);
        if (ld < 0) continue;
        const Chain &o = P.chains[ld / 2];
        if (!ldv(&o.search_done)) continue;
        fence();
        if (atomic_cas(&ch.lead[d], ld, -1) != ld) continue;
        const mtr_wdp_result *src = P.results + (size_t)(ld / 2) * kResSlots + 2 * (ld % 2);
        mtr_wdp_result *dst = P.results + (size_t)chain * kResSlots + 2 * d;
        dst[0] = src[0]; dst[1] = src[1];
        fence();
        atomic_add(&ch.wait_lead, -1);
    }
}
// one thread, only while nothing at all is in flight (the host's rescue pass): a chain still waiting for shared results
// goes back to NEED_SEARCH and will run every one of its search DPs itself -- same results, the sharing is an optimisation
MTR_DEV void unshare_chain(const Ptrs &P, int chain)
{
    Chain &ch = P.chains[chain];
    if (ch.stage != ST_WAIT_SEARCH || ch.wait_lead <= 0 || ch.pending != 0) return;
    ch.lead[0] = ch.lead[1] = -1; ch.wait_lead = 0; ch.no_share = 1;
    ch.stage = ST_NEED_SEARCH;
    atomic_add(&P.ctr->unshared, 1);
}

MTR_DEV void advance_chain(const Ptrs &P, int chain)
{
    Chain &ch = P.chains[chain];
    fence();                                                   // (pending reached zero: the results are behind the fence)
    const mtr_wdp_result *res = P.results + (size_t)chain * kResSlots;
    const int stage = ch.stage;
    if (stage == ST_WAIT_SEARCH) {
        // a direction that shares the DPs of my own other direction
        {
            mtr_wdp_result *mine = P.results + (size_t)chain * kResSlots;
            for (int d = 0; d < 2; d++)
                if (ch.lead[d] <= -2) {
                    const int d2 = -2 - ch.lead[d];
                    ENG_LANE0(mine[2 * d] = mine[2 * d2]; mine[2 * d + 1] = mine[2 * d2 + 1]; ch.lead[d] = -1);
                }
            wsync();
        }
        // my four results are final: hand them to the sibling chains that wait for them (lane = sibling, direction)
        {
            const int set0 = chain - chain % kMaxK;
            for (int q = lane(); q < 2 * kMaxK; q += NL) {
                const int c2 = set0 + q / 2, d2 = q % 2;
                Chain &o = P.chains[c2];
                const int ld = c2 != chain && ldv(&o.stage) == (int)ST_WAIT_SEARCH ? ldv(&o.lead[d2]) : -1;
                // (the follower may be taking the results itself at this moment, see take_shared: whoever clears `lead` delivers)
                if (ld >= 0 && ld / 2 == chain && atomic_cas(&o.lead[d2], ld, -1) == ld) {
                    mtr_wdp_result *dst = P.results + (size_t)c2 * kResSlots + 2 * d2;
                    dst[0] = res[2 * (ld % 2)]; dst[1] = res[2 * (ld % 2) + 1];
                    fence();
                    atomic_add(&o.wait_lead, -1);
                }
            }
            ENG_LANE0(ch.search_done = 1);
        }
        // max_rr of search_De_Bruijn_graph starts cleared; wrap_around_DP (wrap_around_DP.c:357-429) keeps the strictly
        // better of the two penalty sets of a direction (NaN never wins); a cleared record never qualifies
        int best_d = -1, best_s = -1;
        float best_ratio = -1;
        for (int d = 0; d < 2; d++) {
            if (!ch.dir_found[d]) continue;
            int pick_s = -1;
            float pick_ratio = -1;
            for (int s = 0; s < 2; s++) {
                const mtr_wdp_result &r = res[2 * d + s];
                const float ratio = (float)r.n_match / (float)(r.n_match + r.n_mismatch + r.n_ins + r.n_del);
                if (pick_ratio < ratio) { pick_s = s; pick_ratio = ratio; }
            }
            if (pick_s < 0) continue;
            const int period = ch.dir_period[d];
            const int units = res[2 * d + pick_s].n_scanned / period;
            if (best_ratio < pick_ratio && P.min_match_ratio <= pick_ratio && 5 < units && 2 <= period && period < kMaxPeriod) {
                best_ratio = pick_ratio; best_d = d; best_s = pick_s;
            }
        }
        bool cleared = best_d < 0 || !ch.found_last;            // Q4: the LAST walk attempted must have found a loop
        if (!cleared) {
            copy_bytes(unit_ptr(P, chain, U_RR), unit_ptr(P, chain, U_DIR0 + best_d), ch.dir_period[best_d]);
            copy_bytes(score_ptr(P, chain, S_RR), score_ptr(P, chain, S_DIR0 + best_d), ch.dir_period[best_d]);
            Rec r = ch.rr;
            r.period = ch.dir_period[best_d];
            apply_dp(r, ch.qs, res[2 * best_d + best_s], kSearchParams[best_s][0], kSearchParams[best_s][1], kSearchParams[best_s][2]);
            int msg = 0;
            if ((long long)r.period * (ch.qe - ch.qs + 1) > kWrapCap) { msg = 1; cleared = true; }   // handle_one_read.c:89-91
            ENG_LANE0(ch.rr = r; ch.msg += msg);
            wsync();
        }
        if (cleared) {
            ENG_LANE0(rec_clear(ch.rr); ch.stage = ST_DONE);
            wsync();
            return;
        }
        const int coverage = ch.rr.repeat_len / ch.rr.period;
        if (!(5 <= coverage && coverage <= 20 && 5 < ch.rr.period)) {
            ENG_LANE0(ch.stage = ST_DONE);
            wsync();
            return;
        }
        // revise_representative_unit (consensus.c:1048-1087), after polish_repeat
        ENG_LANE0(ch.stage = ST_NEED_POLISH; P.polish_list[atomic_add(&P.ctr->n_polish, 1)] = chain);
        return;
    }
    if (stage == ST_WAIT_CONS) {
        const int *cons = P.aux_of[ch.aux_q] + ch.aux_off;
        const int np = vote_unit(ch.tmp, cons, cons + (size_t)(ch.tmp.period + 1) * 5, unit_ptr(P, chain, U_TMP));
        ENG_LANE0(ch.tmp.period = np);
        wsync();
        if (np < kMaxPeriod) {
            if (np <= 0) {                                      // the reference divides by zero here (H9)
                ENG_LANE0(ch.fatal = ERR_EMPTY_UNIT; rec_clear(ch.rr); ch.stage = ST_DONE);
                wsync();
                return;
            }
            ENG_LANE0(ch.stage = ST_NEED_DP);
            wsync();
            return;
        }
        next_pass_or_done(P, chain);
        return;
    }
    if (stage == ST_WAIT_DP) {
        Rec t = ch.tmp;
        apply_dp(t, ch.tmp.rep_start, res[kResRevise], kReviseParams[ch.pass][0], kReviseParams[ch.pass][1], kReviseParams[ch.pass][2]);
        if (ch.ratio0 < rec_ratio(t)) {                         // ratio0 is never refreshed (Q10)
            copy_bytes(unit_ptr(P, chain, U_RR), unit_ptr(P, chain, U_TMP), t.period);
            ENG_LANE0(ch.rr = t);
        }
        ENG_LANE0(ch.tmp = t);
        wsync();
        next_pass_or_done(P, chain);
        return;
    }
}

// one chain of the polish list: polish_repeat, then the first revise pass
MTR_DEV void polish_chain(const Ptrs &P, int chain, const Scratch &S, const Cta &c, unsigned *smem)
{
    Chain &ch = P.chains[chain];
    polish_rr(P, chain, S, c, smem);
    if (c.warp == 0) {
        ENG_LANE0(ch.ratio0 = rec_ratio(ch.rr); ch.pass = 0);
        start_revise_pass(P, chain);
    }
    cta_sync(c);
}

// ---------------------------------------------------------------- DP task emission (one thread per chain slot)
MTR_CONST int kClassCap[10] = {16, 32, 48, 64, 96, 128, 192, 256, 384, 512};   // G * C of the throughput classes of wdp.cu
MTR_DEV int dp_class_of(int ulen)
{
    for (int c = 0; c < 10; c++) if (ulen <= kClassCap[c]) return c;
    return 9;
}
MTR_DEV int seg_of(int cls20, int rows);
MTR_DEV int row_bucket(int rows)
{
    if (rows < 4) return rows < 0 ? 0 : rows;
    const int lz = 31 - clz((unsigned)rows);
    const int b = lz * 4 + ((rows >> (lz - 2)) & 3);
    return b < kRowBuckets - 1 ? b : kRowBuckets - 1;
}

// segment of a task: family, then fill class, then longest rows first.  All tasks of a class are contiguous: the fill
// kernels keep the warps of an SM on ONE class as long as it has work (the ten instantiations of a family are 200-270 KB
// of code; an SM that runs several of them at once lives on instruction-cache misses)
MTR_DEV int seg_of(int cls20, int rows)
{
    return cls20 * kRowBuckets + (kRowBuckets - 1 - row_bucket(rows));
}
MTR_DEV int bucket_rows(int b)                                 // a row count inside row bucket b
{
    if (b < 4) return b;
    return (4 + (b & 3)) << ((b >> 2) - 2);
}

struct TaskSpec { int first, rows, ulen, uslot, n_param, mode, res_slot; const int *params; };

// reserves direction space (+ consensus block) and task slots, then appends the tasks; false = no room in this wave
// (the chain stays NEED_* and is emitted again by the next wave)
MTR_DEV bool emit_tasks(const Ptrs &P, const DpQueue &Q, int chain, const TaskSpec *sp, int n)
{
    Chain &ch = P.chains[chain];
    const Read &rs = P.reads[ch.read];
    long long dir_need = 0, aux_need = 0;
    int slots = 0;
    bool paired[2] = {false, false};
    for (int i = 0; i < n; i++) {
        const int cls = dp_class_of(sp[i].ulen);
        dir_need += (((long long)sp[i].rows * (kClassCap[cls] / 4) + 15) & ~15LL) * sp[i].n_param;
        if (sp[i].mode == MTR_TB_CONSENSUS) aux_need += ((long long)(sp[i].ulen + 1) * 9 + 3) & ~3LL;
        int gmax = 0;
        for (int p = 0; p < sp[i].n_param; p++) gmax = sp[i].params[3 * p] > gmax ? sp[i].params[3 * p] : gmax;
        // both penalty sets in one int16x2 task while every score (x4, plus tag) fits 15 bits
        paired[i] = sp[i].n_param == 2 && sp[i].mode == MTR_TB_COUNTS && 4LL * gmax * sp[i].rows <= 32760;
        slots += (sp[i].n_param == 2 && !paired[i]) ? 2 : 1;
    }
    const long long d0 = (long long)atomic_add(&Q.qc->dir_used, (unsigned long long)dir_need);
    const long long a0 = (long long)atomic_add(&Q.qc->aux_used, (unsigned long long)aux_need);
    const int t0 = atomic_add(&Q.qc->n_tasks, slots);
    if (d0 + dir_need > Q.dir_cap || a0 + aux_need > Q.aux_cap || t0 + slots > Q.task_cap) {
        // over budget: the slots it took in the task list are marked empty (the counters restart with the queue's next use)
        for (int i = 0; i < slots && t0 + i < Q.task_cap; i++) Q.tasks_in[t0 + i].rows = -1;
        atomic_add(&Q.qc->deferred, 1);
        atomic_add(&P.ctr->deferred, 1);
        return false;
    }
    long long doff = d0, aoff = a0, cells = 0;
    int at = t0, results = 0;
    for (int i = 0; i < n; i++) {
        WdpTask t;
        t.base0 = rs.word_off * 16 + sp[i].first;
        t.rows = sp[i].rows; t.ulen = sp[i].ulen;
        t.unit_off = (long long)((size_t)chain * 4 + sp[i].uslot) * kUnitStride;
        const int cls = dp_class_of(sp[i].ulen);
        t.dir_stride = kClassCap[cls] / 4;
        t.dir_bytes = ((long long)t.rows * t.dir_stride + 15) & ~15LL;
        t.dir_off = doff; doff += t.dir_bytes * sp[i].n_param;
        t.aux_off = 0; t.aux_cap = 0;
        if (sp[i].mode == MTR_TB_CONSENSUS) {
            t.aux_off = aoff; aoff += ((long long)(sp[i].ulen + 1) * 9 + 3) & ~3LL;
            ch.aux_off = (int)t.aux_off; ch.aux_q = Q.id;
        }
        t.result_idx = chain * kResSlots + sp[i].res_slot;
        t.n_param = (unsigned char)sp[i].n_param; t.mode = (unsigned char)sp[i].mode; t.pad_ = 0;
        for (int p = 0; p < 2; p++) {
            const int *q = sp[i].params + 3 * (p < sp[i].n_param ? p : 0);
            t.gain[p] = (short)q[0]; t.mis[p] = (short)q[1]; t.indel[p] = (short)q[2];
        }
        t.cls = (unsigned char)(paired[i] ? 10 + cls : cls);
        if (sp[i].n_param == 2 && !paired[i]) {                 // two int32 tasks
            WdpTask u = t;
            t.n_param = 1; u.n_param = 1;
            u.gain[0] = t.gain[1]; u.mis[0] = t.mis[1]; u.indel[0] = t.indel[1];
            u.dir_off = t.dir_off + t.dir_bytes; u.result_idx = t.result_idx + 1;
            Q.tasks_in[at++] = u;
            atomic_add(&Q.hist[seg_of(u.cls, u.rows)], 1);
        }
        Q.tasks_in[at++] = t;
        atomic_add(&Q.hist[seg_of(t.cls, t.rows)], 1);
        results += sp[i].n_param;
        cells += (long long)t.rows * t.ulen * sp[i].n_param;
        if (paired[i]) atomic_add(&P.ctr->cells_p16, (unsigned long long)((long long)t.rows * t.ulen * 2));
        atomic_add(&P.ctr->slot_cells, (unsigned long long)((long long)t.rows * kClassCap[cls] * sp[i].n_param));
    }
    atomic_add(&P.ctr->cells, (unsigned long long)cells);
    atomic_add(&P.ctr->jobs, (unsigned long long)n);
    atomic_add(&P.ctr->dir_bytes, (unsigned long long)dir_need);
    atomic_add((unsigned long long *)&P.reads[ch.read].ring[ch.ring % kRing].cells, (unsigned long long)cells);
    ch.pending = results;                                      // (the traceback kernels of the queue count it down)
    atomic_add(&P.ctr->dp_pending, results);
    return true;
}

// QS / QL: the short and the long queue that take this wave's tasks (tasks_in == nullptr: none is free, the chain is
// emitted again by a later wave)
MTR_DEV void emit_chain(const Ptrs &P, const DpQueue &QS, const DpQueue &QL, int chain)
{
    Chain &ch = P.chains[chain];
    const int stage = ldv(&ch.stage);
    if (stage != ST_NEED_SEARCH && stage != ST_NEED_CONS && stage != ST_NEED_DP) return;
    fence();                                                   // (a walk may have published NEED_SEARCH a moment ago: read its data after the stage)
    TaskSpec sp[2];
    int n = 0;
    int follow[2] = {-1, -1};                                  // direction d shares the results of chain follow[d] / 2, direction follow[d] % 2
    if (stage == ST_NEED_SEARCH) {
        const int set0 = chain - chain % kMaxK;
        for (int d = 0; d < 2; d++) {
            if (!ch.dir_found[d]) continue;
            // wrap_around_DP.c:260-263 (checked for every direction, shared or not: a leader that aborts hands nothing out)
            if ((long long)(ch.dir_period[d] + 1) * (ch.qe - ch.qs + 1) + ch.dir_period[d] >= kWrapCap) {
                ch.fatal = ERR_WRAPCAP; rec_clear(ch.rr); ch.stage = ST_DONE;
                return;
            }
        }
        for (int d = 0; d < 2; d++) {
            if (!ch.dir_found[d]) continue;
            // A sibling chain of the same candidate (or my own other direction) with the very same unit string runs, or
            // has run, the very same two DPs.  Leaders are chains that own their results (lead < 0): siblings whose
            // search tasks are out already (WAIT_SEARCH or later), or -- among those still to be emitted -- lower ones.
            const int period = ch.dir_period[d];
            const unsigned char *mine = unit_ptr(P, chain, U_DIR0 + d);
            for (int q = 0; q < 2 * kMaxK && follow[d] < 0 && P.share_search && !ch.no_share; q++) {
                const int c2 = set0 + q / 2, d2 = q % 2;
                if (c2 == chain && d2 >= d) continue;
                const Chain &o = P.chains[c2];
                const int st2 = ldv(&o.stage);
                const bool out = st2 == ST_WAIT_SEARCH || st2 == ST_NEED_POLISH || st2 == ST_NEED_CONS || st2 == ST_WAIT_CONS || st2 == ST_NEED_DP || st2 == ST_WAIT_DP ||
                                 (st2 == ST_DONE && o.search_done);
                if (!(out || (st2 == ST_NEED_SEARCH && c2 < chain) || c2 == chain)) continue;
                fence();
                if (o.read != ch.read || o.ring != ch.ring || o.qs != ch.qs || o.qe != ch.qe) continue;      // (a slot of an older candidate)
                if (!o.dir_found[d2] || o.dir_period[d2] != period) continue;
                // a sibling that is out must own these results (lead == -1) and must not wait for anybody itself: results
                // are handed out when a chain advances, so following a chain that waits (perhaps, through others, for me)
                // could close a circle; siblings still to be emitted are followed downwards only
                if (c2 != chain && out && (ldv(&o.lead[d2]) != -1 || ldv(&o.wait_lead) != 0)) continue;
                const unsigned char *other = unit_ptr(P, c2, U_DIR0 + d2);
                bool same = true;
                for (int x = 0; x < period && same; x++) same = other[x] == mine[x];
                if (same) follow[d] = 2 * c2 + d2;
            }
            if (follow[d] >= 0) continue;
            sp[n].first = ch.qs; sp[n].rows = ch.qe - ch.qs + 1; sp[n].ulen = ch.dir_period[d]; sp[n].uslot = U_DIR0 + d;
            sp[n].n_param = 2; sp[n].mode = MTR_TB_COUNTS; sp[n].res_slot = 2 * d; sp[n].params = &kSearchParams[0][0];
            n++;
        }
    } else {
        sp[0].first = ch.tmp.rep_start; sp[0].rows = ch.tmp.rep_end - ch.tmp.rep_start + 1; sp[0].ulen = ch.tmp.period; sp[0].uslot = U_TMP;
        sp[0].n_param = 1; sp[0].mode = stage == ST_NEED_CONS ? MTR_TB_CONSENSUS : MTR_TB_COUNTS; sp[0].res_slot = kResRevise;
        sp[0].params = &kReviseParams[ch.pass][0];
        n = 1;
    }
    for (int i = 0; i < n; i++) {
        // wrap_around_DP.c:260-263: the reference aborts the whole run when (ulen + 1) * i + j reaches WrapDPsize
        if ((long long)(sp[i].ulen + 1) * sp[i].rows + sp[i].ulen >= kWrapCap) {
            ch.fatal = ERR_WRAPCAP; rec_clear(ch.rr); ch.stage = ST_DONE;
            return;
        }
    }
    if (n > 0) {
        const bool is_long = sp[0].rows >= P.long_rows;
        const DpQueue &Q = is_long ? QL : QS;
        if (!Q.tasks_in) { atomic_add(&P.ctr->deferred, 1); return; }
        if (!emit_tasks(P, Q, chain, sp, n)) return;
    }
    if (stage == ST_NEED_SEARCH) {
        // shared directions: results that exist already are copied now, the others arrive when the leader's do
        int waits = 0;
        for (int d = 0; d < 2; d++) {
            if (follow[d] < 0) continue;
            const int c2 = follow[d] / 2, d2 = follow[d] % 2;
            const long long cells = 2LL * (ch.qe - ch.qs + 1) * ch.dir_period[d];
            atomic_add(&P.ctr->shared_cells, (unsigned long long)cells);
            if (c2 != chain && P.chains[c2].search_done) {
                const mtr_wdp_result *src = P.results + (size_t)c2 * kResSlots + 2 * d2;
                mtr_wdp_result *dst = P.results + (size_t)chain * kResSlots + 2 * d;
                dst[0] = src[0]; dst[1] = src[1];
            } else if (c2 == chain) {
                ch.lead[d] = -2 - d2;                           // my own other direction: copied when my results are in
            } else {
                ch.lead[d] = follow[d];
                waits++;
            }
        }
        ch.wait_lead = waits;
        fence();
    }
    ch.stage = stage == ST_NEED_SEARCH ? ST_WAIT_SEARCH : (stage == ST_NEED_CONS ? ST_WAIT_CONS : ST_WAIT_DP);
}

// single thread, first thing in a wave: the reservations of the queues this wave fills start from zero
MTR_DEV void queue_reset(const DpQueue &Q)
{
    Q.qc->n_tasks = 0; Q.qc->deferred = 0; Q.qc->dir_used = 0; Q.qc->aux_used = 0;
}
MTR_DEV void wave_begin(const Ptrs &P, const DpQueue &QS, const DpQueue &QL)
{
    Counters &c = *P.ctr;
    c.n_polish = 0; c.polish_head = 0; c.deferred = 0;
    if (QS.tasks_in) queue_reset(QS);
    if (QL.tasks_in) queue_reset(QL);
    c.waves++;
}

// one warp: prefix sums over the segments (tasks and warp slots), and the reset of the histogram for the next wave
MTR_CONST int kClassJpw[10] = {8, 8, 8, 4, 4, 4, 2, 2, 1, 1};  // tasks per warp slot = 32 / G of the fill classes of wdp.cu
MTR_CONST int kClassCols[10] = {4, 8, 12, 8, 12, 16, 12, 16, 12, 16};   // columns per lane
MTR_DEV void plan_tasks(const Ptrs &P, const DpQueue &Q)
{
    int at = 0, slots = 0;
    for (int s0 = 0; s0 < kSegs; s0 += NL) {                    // (kRowBuckets is a multiple of NL or NL == 1: a step never straddles two classes)
        const int sg = s0 + lane();
        const int h = sg < kSegs ? Q.hist[sg] : 0;
        const int cls = (sg < kSegs ? sg : 0) / kRowBuckets;    // 0 .. 19
        const int jpw = kClassJpw[cls % 10];
        const int sl = (h + jpw - 1) / jpw;
        const int off = wscan_excl(h), soff = wscan_excl(sl);
        if (sg < kSegs) { Q.seg_task[sg] = at + off; Q.seg_slot[sg] = slots + soff; Q.hist[sg] = 0; Q.bucket_cursor[sg] = 0; }
        // estimated work of the class: slot rows x columns per lane, in units of 1024 (saturating)
        const int rows = bucket_rows(kRowBuckets - 1 - (sg < kSegs ? sg : 0) % kRowBuckets);
        long long work = (long long)sl * rows * kClassCols[cls % 10];
        work = wsum(work);
        if (lane() == 0) {
            const long long prev = (s0 % kRowBuckets) == 0 ? 0 : (long long)Q.class_begin[cls] * 1024;
            const long long tot = prev + work;
            Q.class_begin[cls] = (int)((tot + 1023) / 1024 < 0x7fffffffLL ? (tot + 1023) / 1024 : 0x7fffffffLL);
        }
        at += wsum(h); slots += wsum(sl);
    }
    ENG_LANE0(Q.seg_task[kSegs] = at; Q.seg_slot[kSegs] = slots; Q.class_begin[WDP_NCLASS] = at; atomic_add(&P.ctr->tasks_total, (unsigned long long)at));
    for (int q = lane(); q < 20; q += NL) {
        const unsigned long long old = P.ctr->class_share[q];
        P.ctr->class_share[q] = old - old / 16 + (unsigned long long)(unsigned)Q.class_begin[q];
    }
    for (int q = lane(); q < WDP_NCLASS; q += NL) Q.slot_counter[q] = 0;
    wsync();
}

MTR_DEV void scatter_task(const DpQueue &Q, int i)
{
    const WdpTask t = Q.tasks_in[i];
    if (t.rows < 0) return;                                    // slot of a deferred emission
    const int sg = seg_of(t.cls, t.rows);
    Q.tasks[Q.seg_task[sg] + atomic_add(&Q.bucket_cursor[sg], 1)] = t;
}

// ---------------------------------------------------------------- per-read scheduler: handle_one_TR's candidate loop
// (handle_one_read.c:227-246) with candidate look-ahead.  An accepted repeat found from candidate (qs', qe') prunes
// only ranges that start inside it and end before its rep_end (:178-188): a later candidate whose range ends beyond
// every in-flight qe' can (almost) never be pruned by them and is evaluated concurrently; up to `speculate` further
// candidates are evaluated ahead of ones that could still prune them.  Results are committed in candidate order; when
// a commit accepts a repeat, every in-flight candidate whose start it has just pruned is dropped -- the reference
// would never have visited it.  Same bytes for any depth.
MTR_DEV void free_set(const Ptrs &P, Read &rs, int read, int set)
{
    if (set < 0) return;
    for (int c = lane(); c < kMaxK; c += NL) P.chains[((size_t)read * kSets + set) * kMaxK + c].stage = ST_FREE;
    wsync();
    ENG_LANE0(rs.set_mask &= ~(1u << set));
    wsync();
}

// chain set of a candidate that is dropped while its chains may still be with the walk kernel or a long DP queue: queued
// walks are cancelled, running ones are told that nobody waits for them, DP results in flight are left to arrive; the set is
// freed by a later scheduler pass once all of that has ended
MTR_DEV void drop_set(const Ptrs &P, Read &rs, int read, int set)
{
    if (set < 0) return;
    int busy = 0;
    for (int c = lane(); c < kMaxK; c += NL) {
        Chain &ch = P.chains[((size_t)read * kSets + set) * kMaxK + c];
        int st = ldv(&ch.stage);
        if (st == ST_WALK) st = atomic_cas(&ch.stage, (int)ST_WALK, (int)ST_DONE) == (int)ST_WALK ? (int)ST_DONE : ldv(&ch.stage);
        if (st == ST_WALKING) st = atomic_cas(&ch.stage, (int)ST_WALKING, (int)ST_ZOMBIE_WALKING) == (int)ST_WALKING ? (int)ST_ZOMBIE_WALKING : ldv(&ch.stage);
        if (st == ST_ZOMBIE_WALKING) busy = 1;
        else if (st != ST_FREE) ch.stage = ST_DONE;            // nothing of this chain is emitted or advanced any more ...
        if (ldv(&ch.pending) > 0) busy = 1;                    // ... but DP tasks of a long queue may still write into it
    }
    busy = wmax(busy);
    wsync();
    if (busy) ENG_LANE0(rs.zombie_mask |= 1u << set);
    else free_set(P, rs, read, set);
}

MTR_DEV void sched_read(const Ptrs &P, int read, unsigned long long *table_mem, unsigned table_cap, int *sh)
{
    Read &rs = P.reads[read];
    int started = 0;
    // chain sets of dropped candidates whose walks were still running: free them once the walks have let go
    for (unsigned zm = rs.zombie_mask; zm; zm &= zm - 1u) {
        const int set = ffs(zm) - 1;
        bool busy = false;
        for (int c = 0; c < kMaxK; c++) {
            const int st = ldv(&P.chains[((size_t)read * kSets + set) * kMaxK + c].stage);
            if (st == ST_WALK || st == ST_WALKING || st == ST_ZOMBIE_WALKING) busy = true;
            if (ldv(&P.chains[((size_t)read * kSets + set) * kMaxK + c].pending) > 0) busy = true;
        }
        if (!busy) { ENG_LANE0(rs.zombie_mask &= ~(1u << set)); free_set(P, rs, read, set); }
    }
  next_read:
    if (rs.phase != 0) {
        // a free slot takes the next read of the batch -- once nothing of its previous read is in flight any more
        if (rs.set_mask != 0u || started >= kSchedBudget) return;
        int idx = P.n_total;
        if (lane() == 0 && ldv(&P.ctr->next_read) < P.n_total) idx = atomic_add(&P.ctr->next_read, 1);
        idx = bcast(idx, 0);
        if (idx >= P.n_total) return;
        const ReadDesc d = P.descs[idx];
        ENG_LANE0(rs.word_off = d.word_off; rs.pos_off = d.pos_off; rs.L = d.L; rs.cursor = 0; rs.head = 0; rs.n_ring = 0;
                  rs.n_accepted = 0; rs.candidates = 0; rs.id = d.id; rs.zombie_mask = 0u; rs.cells_wasted = 0; rs.phase = 0);
        wsync();
    }
    const uint32_t *rd = P.packed + rs.word_off;
    int *END = P.end + rs.pos_off, *WW = P.w + rs.pos_off;
    const int L = rs.L;
    for (;;) {
        // ---- commit finished candidates in candidate order
        while (rs.n_ring > 0) {
            Cand &cd = rs.ring[rs.head % kRing];
            int best_c = -1;
            bool done = true;
            int fatal = 0, msgs = 0;
            if (cd.set >= 0) {
                const size_t c0 = ((size_t)read * kSets + cd.set) * kMaxK;
                for (int c = 0; c < cd.n_k; c++) if (P.chains[c0 + c].stage != ST_DONE) { done = false; break; }
                if (!done) break;
                // find_tandem_repeat's pick over k (handle_one_read.c:135-146)
                float best_ratio = -1;
                for (int c = 0; c < cd.n_k; c++) {
                    const Chain &ch = P.chains[c0 + c];
                    if (ch.fatal && !fatal) fatal = ch.fatal;
                    msgs += ch.msg;
                    const float ratio = rec_ratio(ch.rr);
                    if (best_ratio < ratio && P.min_match_ratio <= ratio && 5 < ch.rr.units && 2 <= ch.rr.period) { best_ratio = ratio; best_c = c; }
                }
            }
            if (fatal) {
                ENG_LANE0(if (atomic_max(&P.ctr->error, fatal) < fatal) P.ctr->error_read = rs.id; rs.phase = 1; atomic_add(&P.ctr->unfinished, -1));
                wsync();
                return;
            }
            ENG_LANE0(rs.candidates++; atomic_add(&P.ctr->progress, 1); if (msgs) atomic_add(&P.ctr->msgs, msgs));
            bool accepted = false;
            Rec pick;
            rec_clear(pick);
            if (best_c >= 0) {
                const int chain = (int)(((size_t)read * kSets + cd.set) * kMaxK + best_c);
                pick = P.chains[chain].rr;
                if (pick.repeat_len > 0 && pick.rep_start + 10 < pick.rep_end) {        // handle_one_TR's accept (:236-243)
                    accepted = true;
                    int slot = 0;
                    ENG_LANE0(slot = atomic_add(&P.ctr->n_accepted, 1));
                    slot = bcast(slot, 0);
                    if (slot >= P.acc_cap) {
                        ENG_LANE0(atomic_max(&P.ctr->error, ERR_ACCEPTED_FULL); rs.phase = 1; atomic_add(&P.ctr->unfinished, -1));
                        wsync();
                        return;
                    }
                    Accepted &a = P.acc[slot];
                    copy_bytes(a.unit, unit_ptr(P, chain, U_RR), pick.period < kUnitStride ? pick.period : kUnitStride);
                    ENG_LANE0(a.read = rs.id; a.seq = rs.n_accepted; a.rec = pick; rs.n_accepted++);
                    // remove_redundant_ranges_from_directional_index (:178-188)
                    const int hi = pick.rep_end < L ? pick.rep_end : L;
                    for (int i = pick.rep_start + lane(); i < hi; i += NL)
                        if (END[i] >= 0 && END[i] < pick.rep_end) { END[i] = -1; WW[i] = -1; }
                    wsync();
                }
            }
            const int front_set = cd.set;
            ENG_LANE0(rs.head = (rs.head + 1) % kRing; rs.n_ring--);
            wsync();
            free_set(P, rs, read, front_set);
            if (accepted) {
                // in-flight candidates whose range this repeat has just pruned would never have been visited: drop them
                int keep = 0;
                const int n = rs.n_ring;
                for (int c = 0; c < n; c++) {
                    const Cand cc = rs.ring[(rs.head + c) % kRing];
                    if (END[cc.qs] < 0) {
                        ENG_LANE0(rs.cells_wasted += cc.cells; atomic_add(&P.ctr->spec_cells, (unsigned long long)cc.cells));
                        drop_set(P, rs, read, cc.set);
                    } else {
                        if (keep != c) {
                            ENG_LANE0(rs.ring[(rs.head + keep) % kRing] = cc);
                            if (cc.set >= 0)
                                for (int q = lane(); q < cc.n_k; q += NL) P.chains[((size_t)read * kSets + cc.set) * kMaxK + q].ring = rs.head + keep;
                            wsync();
                        }
                        keep++;
                    }
                }
                ENG_LANE0(rs.n_ring = keep);
                wsync();
            }
        }
        // ---- start further candidates
        int cur = rs.cursor;
        for (;;) {                                              // next live range: -1 < END[qs] < L (:228-229)
            const int i = cur + lane();
            const bool live = i < L && END[i] > -1 && END[i] < L;
            const unsigned m = wballot(live || i >= L);
            if (m) { cur += ffs(m) - 1; break; }
            cur += NL;
        }
        if (cur > L) cur = L;
        ENG_LANE0(rs.cursor = cur);
        wsync();
        if (cur >= L) {
            if (rs.n_ring == 0) {
                ENG_LANE0(rs.phase = 1; atomic_add(&P.ctr->unfinished, -1); atomic_add(&P.ctr->candidates, (unsigned long long)rs.candidates));
                wsync();
                goto next_read;
            }
            return;
        }
        const int qs = cur, qe = END[cur];
        if (rs.n_ring >= kRing || started >= kSchedBudget) return;
        bool safe = true;
        int n_spec = 0;
        for (int c = 0; c < rs.n_ring; c++) {
            const Cand &cc = rs.ring[(rs.head + c) % kRing];
            if (cc.set < 0) continue;                           // died at the gate: can never accept, never prunes
            if (cc.qe >= qe) safe = false;
            n_spec += cc.spec;
        }
        if (!safe && n_spec >= P.speculate) return;
        const int cw = WW[cur];
        int min_k, max_k;                                       // handle_one_read.c:105-120
        if (cw < 100) { min_k = 2; max_k = 10; } else if (cw < 1000) { min_k = 2; max_k = 12; } else { min_k = 5; max_k = 15; }
        const int n_k = max_k - min_k + 1;
        // The maxFreq gate (consensus.c:532).  A k'-mer that occurs c times has a k-prefix (k < k') that occurs at
        // least c times at the same coded positions, so maxFreq(k') <= maxFreq(k) + (raw-base entries of the k'
        // window, Q7): once that bound is <= 5 no larger k can pass.  Small windows are counted right here; for large
        // ones every k goes to the walk kernel, which applies the gate itself.
        const int width = qe - qs + 1;
        unsigned pass_mask = 0;
        if (width <= kInlineWindow) {
            int low_maxf = 1 << 30;
            for (int k = min_k; k <= max_k; k++) {
                const int coded_end = qe < L - k + 1 ? qe : L - k + 1;
                const int raw = qe - coded_end + 1;
                if (low_maxf + raw <= 5) continue;
                const Window win = window_make(rd, L, k, qs, qe);
                const Table tb = table_wide(table_mem, table_cap, width);
                const int maxf = table_build(tb, win, cta_of_warp(sh));
                if (lane() == 0) { atomic_add(&P.ctr->tables, 1ull); atomic_add(&P.ctr->table_positions, (unsigned long long)width); }
                if (maxf < low_maxf) low_maxf = maxf;
                if (5 < maxf) pass_mask |= 1u << (k - min_k);
            }
        } else {
            pass_mask = (1u << n_k) - 1u;
        }
        int set = -1;
        if (pass_mask) {
            const unsigned free_mask = ~rs.set_mask & ((1u << kSets) - 1u);
            if (!free_mask) return;                             // every chain set is busy: wait for a commit
            set = ffs(free_mask) - 1;
        }
        const int pos = rs.head + rs.n_ring;
        wsync();
        if (lane() == 0) {
            Cand &cd = rs.ring[pos % kRing];
            cd.qs = qs; cd.qe = qe; cd.set = set; cd.spec = safe ? 0 : 1; cd.min_k = min_k; cd.n_k = n_k; cd.cells = 0;
            rs.n_ring++;
            rs.cursor = cur + 1;
            if (set >= 0) rs.set_mask |= 1u << set;
        }
        wsync();
        started++;
        if (set >= 0) {
            for (int c = lane(); c < n_k; c += NL) {
                const int chain = (int)(((size_t)read * kSets + set) * kMaxK + c);
                Chain &ch = P.chains[chain];
                rec_clear(ch.rr); rec_clear(ch.tmp);
                ch.rr.kmer = min_k + c;
                ch.k = min_k + c; ch.pass = 0; ch.found_last = 0; ch.ratio0 = 0;
                ch.read = read; ch.qs = qs; ch.qe = qe;
                ch.dir_found[0] = ch.dir_found[1] = 0; ch.dir_period[0] = ch.dir_period[1] = 0;
                ch.fatal = 0; ch.msg = 0; ch.ring = pos; ch.aux_off = 0; ch.pending = 0; ch.aux_q = 0;
                ch.lead[0] = ch.lead[1] = -1; ch.wait_lead = 0; ch.search_done = 0; ch.no_share = 0;
                if (pass_mask & (1u << c)) {
                    // a walk kernel may be running right now and may hold a stale queue entry for this very chain (a
                    // cancelled walk of the set's previous owner): the fields first, then the stage
                    fence();
                    atomic_exch(&ch.stage, (int)ST_WALK);
                    // two walk kernels: 32 KB count tables (DIRECT k <= 7, COMPACT up to 3070 positions; several ctas per SM)
                    // and 96 KB ones (COMPACT up to 12286 positions; one cta per SM)
                    if (min_k + c > P.direct_max_k && width > compact_max_n(kCompactCapSmall))
                        P.walk_ring_big[atomic_add(&P.ctr->walk_tail_big, 1u) & P.walk_ring_mask] = chain;
                    else
                        P.walk_ring[atomic_add(&P.ctr->walk_tail, 1u) & P.walk_ring_mask] = chain;
                } else {
                    rec_clear(ch.rr);
                    ch.stage = ST_DONE;
                }
            }
            wsync();
        }
    }
}

}  // namespace eng
