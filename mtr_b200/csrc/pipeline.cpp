// pipeline.cpp -- host side of the drop-in: mTR's entry points (handle_one_file / handle_one_read) on top of
// the CUDA kernels.  Replaces /root/reference/handle_one_file.c, handle_one_read.c, the control logic of
// consensus.c (unit finder, polish, revise) and chaining.cpp; the arithmetic-heavy stages run on the GPU:
//   fill_directional_index_with_end  -> mtr_di_run   (K1/K2, di.cu)
//   wrap_around_DP / _sub, the DP of revise_representative_unit_sub, pretty_print_alignment -> mtr_wdp_run (K3, wdp.cu)
//
// The reference evaluates one candidate range of one read at a time and calls the DP synchronously.  Here every
// read of a batch is a small state machine; one "round" advances all reads in parallel on the host cores until
// each of them needs DP results, the DP jobs of all reads go to the GPU as one batch, and the next round consumes
// them.  Inside one candidate the 9-11 k values, both walk directions and both penalty sets are independent, so
// they share a round; the revise chain (consensus DP -> DP, twice) and the candidate loop itself (an accepted
// repeat prunes later candidates, handle_one_read.c:178-188,243) stay sequential per read, exactly as in the
// reference.  There is no CPU implementation of the DP or of the directional index in this file: without a
// usable GPU handle_one_file fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <time.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <vector>
#include "mtr_internal.h"

// ================================================================ the reference's globals (mTR.h:61-96,142-143)
extern "C" {
int   Manhattan_Distance = 1;
float min_match_ratio = 0.6f;
int  *orgInputString = nullptr;
float time_all, time_memory, time_range, time_period, time_initialize_input_string, time_wrap_around_DP,
      time_count_table, time_chaining;
int   query_counter;
}

namespace {

constexpr int kMaxLen = 1000000;      // MAX_INPUT_LENGTH, mTR.h:31
constexpr int kMaxPeriod = 500;       // MAX_PERIOD, mTR.h:34
constexpr int kMaxTies = 1024;        // MAX_tiebreaks, mTR.h:46
constexpr long long kWrapCap = 200000000LL;   // WrapDPsize, mTR.h:51
constexpr int kMaxInflightCands = 24;         // per read; bounds the jobs a read can queue in one round

struct Pow4 { int v[16]; Pow4() { v[0] = 1; for (int i = 1; i < 16; i++) v[i] = v[i - 1] * 4; } };
const Pow4 P4;

double thread_cpu_s()
{
    timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- the per-repeat feature record (mTR.h:99-119)
struct Rec {
    int inputLen = -1, rep_start = -1, rep_end = -1, repeat_len = -1, period = -1, units = -1;
    int nm = -1, nx = -1, ni = -1, nd = -1, kmer = -1, gain = -1, mis = -1, indel = -1;
    std::vector<uint8_t> unit;        // bases 0..3
    std::vector<int> score;           // string_score of the walk that produced the unit
    void clear() { *this = Rec(); }   // clear_rr, fill_directional_index.c:40-60
    float ratio() const { return (float)nm / (nm + nx + ni + nd); }       // e.g. wrap_around_DP.c:398
};

// apply one DP result the way wrap_around_DP_sub fills its record (wrap_around_DP.c:337-350)
void apply_dp(Rec &r, int qs, const mtr_wdp_result &d, int g, int m, int in)
{
    r.rep_start = qs + d.end_i + 1;
    r.rep_end = qs + d.max_i;
    r.repeat_len = d.max_i - d.end_i;
    r.units = d.n_scanned / r.period;
    r.nm = d.n_match; r.nx = d.n_mismatch; r.ni = d.n_ins; r.nd = d.n_del;
    r.gain = g; r.mis = m; r.indel = in;
}

// ---------------------------------------------------------------- exact k-mer counts of one window
// (init_inputString + generate_freqNode_*, consensus.c:37-253; the hash layout is unobservable)
constexpr int kDirectK = 8;           // k <= 8: direct table of 4^k counters (256 KB at k = 8; 3 ns per position against 6-10 ns for the
                                      // hash table, and cheaper look-ups in the walks); k = 9, 10 direct (1-4 MB) gained nothing in the
                                      // pipeline; larger k: open addressing
struct Counter {
    std::vector<int> codes;           // codes[i - qs] for i in [qs, qe]
    std::vector<int> direct;          // k <= kDirectK: 4^k counters
    // k > kDirectK: open addressing, one 8-byte slot per node (key + 1 in the high word so that 0 means empty); the slots
    // a window touched are cleared again from `codes`, so a build costs O(window), not O(table)
    std::vector<uint64_t> slots;
    std::vector<uint32_t> touched;
    std::vector<uint32_t> pos_slot;   // slot of every window position (k > kDirectK): the listing pass needs no second probe
    uint32_t hmask = 0;
    int k = 0, maxf = -1;

    static inline uint32_t hash(uint32_t node) { return node * 2654435761u; }
    void build(const uint8_t *org, int L, int kk, int qs, int qe)
    {
        if (k <= kDirectK) { for (int c : codes) direct[c] = 0; }      // undo the previous window (O(window), not O(4^k))
        else { for (uint32_t h : touched) slots[h] = 0; touched.clear(); }
        k = kk;
        const int n = qe - qs + 1;
        codes.resize(n);
        const int coded_end = std::min(qe, L - k + 1);        // codes for i < coded_end (:48-51)
        int carry = 0;
        for (int i = qs; i < qs + k - 1; i++) carry = 4 * carry + (i < L ? org[i] : 0);
        const int mask = P4.v[k - 1];
        for (int i = qs; i <= qe; i++) {
            if (i < coded_end) {
                const int c = 4 * carry + org[i + k - 1];
                codes[i - qs] = c;
                carry = c & (mask - 1);                       // mask = 4^(k-1)
            } else {
                // raw base left by the copy loop (:42-44); index L itself is stale in the reference (H4b): 0 here
                codes[i - qs] = i < L ? org[i] : 0;
            }
        }
        maxf = -1;
        if (k <= kDirectK) {
            if (direct.size() < (size_t)P4.v[kDirectK]) direct.assign((size_t)P4.v[kDirectK], 0);
            for (int c : codes) maxf = std::max(maxf, ++direct[c]);
        } else {
            uint32_t cap = 1024;
            while (cap < 2u * (uint32_t)(n + 1)) cap <<= 1;
            if (slots.size() < cap) slots.assign(cap, 0);
            hmask = cap - 1;
            touched.reserve(n);
            pos_slot.resize(n);
            constexpr int kAhead = 12;                          // the table of a long window does not fit L1/L2: prefetch
            for (int i = 0; i < n; i++) {
                if (i + kAhead < n) __builtin_prefetch(&slots[hash((uint32_t)codes[i + kAhead]) & hmask], 1, 1);
                const int c = codes[i];
                const uint64_t key = ((uint64_t)(uint32_t)c + 1) << 32;
                uint32_t h = hash((uint32_t)c) & hmask;
                for (;;) {
                    const uint64_t sl = slots[h];
                    if (sl == 0) { slots[h] = key | 1u; touched.push_back(h); maxf = std::max(maxf, 1); break; }
                    if ((sl & 0xffffffff00000000ull) == key) { slots[h] = sl + 1; maxf = std::max(maxf, (int)(uint32_t)(sl + 1)); break; }
                    h = (h + 1) & hmask;
                }
                pos_slot[i] = h;
            }
        }
    }
    int get(int node)                                          // freq_node, consensus.c:231-253
    {
        if (k <= kDirectK) return (node >= 0 && node < P4.v[k]) ? direct[node] : 0;
        const uint64_t key = ((uint64_t)(uint32_t)node + 1) << 32;
        uint32_t h = hash((uint32_t)node) & hmask;
        for (;;) {
            const uint64_t sl = slots[h];
            if (sl == 0) return 0;
            if ((sl & 0xffffffff00000000ull) == key) return (int)(uint32_t)sl;
            h = (h + 1) & hmask;
        }
    }
    void decrement(int node)
    {
        if (k <= kDirectK) { direct[node]--; return; }
        const uint64_t key = ((uint64_t)(uint32_t)node + 1) << 32;
        uint32_t h = hash((uint32_t)node) & hmask;
        while ((slots[h] & 0xffffffff00000000ull) != key) h = (h + 1) & hmask;
        slots[h]--;
    }
    int max_freq() const { return maxf; }                      // counts only grow while building: running max == final max
    // generate_freqNode_return_list_maxNodes (:132-229): listing a node decrements its count
    int list_max_nodes(int *list, int cap, int maxfreq)
    {
        int n = 0;
        if (k <= kDirectK) {
            for (int c : codes)
                if (direct[c] == maxfreq) { list[n++] = c; direct[c]--; if (cap <= n) break; }
        } else {
            const int len = (int)codes.size();
            for (int i = 0; i < len; i++) {
                uint64_t &sl = slots[pos_slot[i]];
                if ((int)(uint32_t)sl == maxfreq) { list[n++] = codes[i]; sl--; if (cap <= n) break; }
            }
        }
        return n;
    }
};

// ---------------------------------------------------------------- greedy de Bruijn walk (consensus.c:269-505)
// From step 10 on the look-ahead depth is constant (k), so the next node is a pure function of the current node
// and of the (now frozen) count table.  WalkMemo caches that function across the <= 100 start nodes of one
// (window, k, direction) and marks the nodes a walk has visited: coming back to a visited node means the walk is
// caught in a cycle that does not contain its start node, i.e. it can only run out its step limit -- the result
// ("no loop") is returned at once.  Both shortcuts leave every observable result unchanged.
struct WalkMemo {
    struct Entry { uint32_t key, epoch; int next, seen; };
    // open addressing; the table starts small enough to stay in L1/L2 (most windows touch a few hundred nodes) and
    // doubles when a (window, k, direction) fills it beyond a half
    std::vector<Entry> tab;
    uint32_t epoch = 0, mask = 0, used = 0;
    int serial = 0;
    void reset()
    {
        if (tab.empty()) { tab.assign(1u << 10, Entry{0, 0, 0, 0}); mask = (1u << 10) - 1; }
        if (++epoch == 0) { for (Entry &e : tab) e.epoch = 0; epoch = 1; }
        serial = 0;
        used = 0;
    }
    void grow()
    {
        std::vector<Entry> old;
        old.swap(tab);
        tab.assign(old.size() * 2, Entry{0, 0, 0, 0});
        mask = (uint32_t)tab.size() - 1;
        for (const Entry &e : old) {
            if (e.epoch != epoch) continue;
            uint32_t h = (e.key * 2654435761u) & mask;
            while (tab[h].epoch == epoch) h = (h + 1) & mask;
            tab[h] = e;
        }
    }
    Entry *slot(uint32_t node)
    {
        if (2 * used > mask) grow();
        uint32_t h = (node * 2654435761u) & mask;
        for (;;) {
            Entry &e = tab[h];
            if (e.epoch != epoch) { e.epoch = epoch; e.key = node; e.seen = -1; e.next = -1; used++; return &e; }
            if (e.key == node) return &e;
            h = (h + 1) & mask;
        }
    }
};

bool walk(Counter &cnt, WalkMemo &memo, int qs, int qe, int start, int k, bool backward, Rec &out)
{
    int ustr[kMaxPeriod], uscore[kMaxPeriod];
    int ties[kMaxTies], fresh[kMaxTies];
    int node = start, period = 0;
    const int limit = (qe - qs) / 5;                       // MIN_NUM_FREQ_UNIT
    const int serial = memo.serial++;
    for (int l = 0; l < kMaxPeriod && l < limit; l++) {
        if (!backward) { ustr[l] = node >> (2 * (k - 1)); uscore[l] = cnt.get(node); }
        WalkMemo::Entry *me = nullptr;
        int next = -1;
        if (l >= 10) {
            me = memo.slot((uint32_t)node);
            if (me->seen == serial) return false;           // cycle without the start node
            me->seen = serial;
            next = me->next;
        }
        if (next < 0) {
            int m, pick = 0, nties = 1;
            ties[0] = 0;
            const int depth = l < 10 ? 1 : k;
            for (m = 1; m <= depth; m++) {
                int best = -1, nf = 0;
                pick = 0;
                for (int t = 0; t < nties; t++)
                    for (int b = 0; b < 4; b++) {
                        // 4^x divisions and remainders of the reference as shifts and masks (all values >= 0)
                        const int digits = backward ? (b << (2 * (m - 1))) + ties[t] : 4 * ties[t] + b;
                        const int cand = backward ? (digits << (2 * (k - m))) + (node >> (2 * m))
                                                  : ((node & ((1 << (2 * (k - m))) - 1)) << (2 * m)) + digits;
                        const int c = cnt.get(cand);
                        if (best < c) { best = c; pick = digits; nf = 0; fresh[nf++] = digits; }
                        else if (best == c && nf < kMaxTies) fresh[nf++] = digits;
                    }
                if (backward ? nf <= 1 : nf == 1) break;
                std::copy(fresh, fresh + nf, ties);
                nties = nf;
            }
            next = backward ? ((pick & 3) << (2 * (k - 1))) + (node >> 2)
                            : 4 * (node & ((1 << (2 * (k - 1))) - 1)) + (pick >> (2 * (m - 1)));   // unresolved ties append 'A' (:336)
            if (me) me->next = next;
        }
        node = next;
        if (backward) { ustr[l] = node >> (2 * (k - 1)); uscore[l] = cnt.get(node); }
        if (node == start) { period = l + 1; if (kMaxPeriod <= period) period = 0; break; }
    }
    if (period == 0) return false;
    out.period = period;
    out.unit.resize(period); out.score.resize(period);
    for (int i = 0; i < period; i++) {
        const int s = backward ? period - 1 - i : i;
        out.unit[i] = (uint8_t)ustr[s]; out.score[i] = uscore[s];
    }
    return true;
}

// ---------------------------------------------------------------- polish_repeat (consensus.c:584-704)
int align_score(Counter &cnt, int start, int k, int node, int period, const uint8_t *unit)
{
    int sum = 0;
    for (int j = start; 0 <= j && start - k < j; j--) {
        node = unit[j % period] * P4.v[k - 1] + node / 4;
        sum += cnt.get(node);
    }
    return sum;
}

bool suspicious(const Rec &r, int j)
{
    int c = 0;
    for (int i = 0; i < r.kmer - 1 && 0 <= j - i; i++) {
        const int sc = (j - i) < (int)r.score.size() ? r.score[j - i] : -1;
        if (sc < 2) c++;
    }
    return (r.kmer - 1) * 0.8 < (double)c;
}

void polish(Counter &cnt, const uint8_t *org, int L, Rec &r)
{
    const int k = r.kmer, period = r.period;
    if (period <= k) return;
    cnt.build(org, L, k, r.rep_start, r.rep_end);
    const std::vector<uint8_t> unit = r.unit;
    uint8_t revised[kMaxPeriod];
    int jr = kMaxPeriod - 1;
    int best = 0;
    for (int i = 0; i < k; i++) best = unit[i] * P4.v[k - 1 - i] + best;
    for (int j = period - 1; 0 <= j;) {
        const int ref = unit[j] * P4.v[k - 1] + best / 4;
        int best_freq = cnt.get(ref);
        best = ref;
        const int sc = j < (int)r.score.size() ? r.score[j] : -1;
        if (sc == 1 && suspicious(r, j)) {
            for (int l = 0; l < 4; l++) {
                const int alt = (ref + (l - unit[j]) * P4.v[k - 1]) % P4.v[k];
                if (best_freq < cnt.get(alt)) { best_freq = cnt.get(alt); best = alt; }
            }
            if (best == ref) {
                revised[jr--] = unit[j--];
            } else {
                const int s_del = align_score(cnt, j, k, best, period, unit.data());
                const int s_sub = align_score(cnt, j - 1, k, best, period, unit.data());
                int s_ins = -1;
                if (j >= 1 && best / P4.v[k - 1] == unit[(j - 1) % period])
                    s_ins = align_score(cnt, j - 2, k, best, period, unit.data());
                revised[jr--] = (uint8_t)(best / P4.v[k - 1]);
                const int mx = std::max(std::max(s_del, s_sub), s_ins);
                if (mx == s_del) {} else if (mx == s_sub) j -= 1; else j -= 2;
            }
        } else {
            revised[jr--] = unit[j--];
        }
        if (jr < 0) return;                                 // "fails to revise": record unchanged
    }
    r.period = (kMaxPeriod - 1) - jr;
    r.unit.assign(revised + jr + 1, revised + kMaxPeriod);
}

// ---------------------------------------------------------------- min_missing (consensus.c:714-820)
// each row of min_missing_bases[10][10][20] starts at 1 and steps by 0 or 1: stored as 19 step bits
const unsigned kMissingSteps[10][10] = {
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

int min_missing(int period, double error, int coverage)
{
    static const int plim[9] = {200, 150, 100, 75, 50, 30, 20, 10, 5};
    static const double elim[9] = {0.25, 0.225, 0.2, 0.175, 0.15, 0.125, 0.1, 0.075, 0.05};
    int i = 9, j = 9;
    for (int t = 0; t < 9; t++) if (period > plim[t]) { i = t; break; }
    for (int t = 0; t < 9; t++) if (error > elim[t]) { j = t; break; }
    const int k = coverage <= 1 ? 0 : (coverage >= 20 ? 19 : coverage - 1);
    return 1 + __builtin_popcount(kMissingSteps[i][j] & ((1u << k) - 1u));
}

// majority vote of revise_representative_unit_sub (consensus.c:964-1013) from the two histograms
void vote_unit(Rec &r, const int *cons, const int *miss)
{
    const int ulen = r.period;
    std::vector<uint8_t> revised;
    revised.reserve(2 * ulen);
    const int coverage = r.repeat_len / r.period;
    for (int j = 1; j <= ulen; j++) {
        int mv = -1, mb = -1;
        for (int q = 0; q < 5; q++) if (mv < cons[j * 5 + q]) { mv = cons[j * 5 + q]; mb = q; }
        if (mb < 4) revised.push_back((uint8_t)mb);
        mv = -1; int mm = -1;
        for (int q = 0; q < 4; q++) if (mv < miss[j * 4 + q]) { mv = miss[j * 4 + q]; mm = q; }
        if (5 <= coverage && coverage <= 20) {
            const double mismatch_ratio = (double)(r.nx + r.ni + r.nd) / r.repeat_len;
            if (min_missing(r.period, mismatch_ratio, coverage) <= mv && 0 <= mm && mm <= 3) revised.push_back((uint8_t)mm);
        }
    }
    r.period = (int)revised.size();
    r.unit.swap(revised);
}

// ---------------------------------------------------------------- chaining + printing (chaining.cpp:43-363)
struct ChainItem { Rec rec; int start, end, score; ChainItem *pred; };

// Returns the records of the best chain, oldest first.  The set of alignments is taken in insertion order
// (canonical tie-break for the reference's pointer-ordered std::set); the sweep issues the same
// std::multimap operations in the same order as chaining.cpp:262-335, including its erase-then-increment loop.
std::vector<const Rec *> best_chain(std::vector<ChainItem> &items)
{
    std::vector<const Rec *> out;
    if (items.empty()) return out;
    typedef std::multimap<int, ChainItem *> MM;
    MM by_x, by_y;
    for (ChainItem &a : items)
        if (a.start + 10 <= a.end) {
            by_x.insert(std::make_pair(a.start, &a));
            by_x.insert(std::make_pair(a.end - 10, &a));
        }
    for (MM::iterator ev = by_x.begin(); ev != by_x.end(); ev++) {
        ChainItem *cur = ev->second;
        if (cur->start == ev->first) {
            if (by_y.empty()) continue;
            const int lim = cur->start + 10;
            MM::iterator y = by_y.begin(), prev = y;
            for (; y != by_y.end(); prev = y, y++)
                if (prev->second->end <= lim && y->second->end > lim) {
                    cur->pred = prev->second; cur->score += prev->second->score;
                    break;
                }
            if (prev->second->end <= lim && y == by_y.end()) { cur->pred = prev->second; cur->score += prev->second->score; }
        } else if (by_y.empty()) {
            by_y.insert(std::make_pair(cur->end, cur));
        } else {
            bool keep = true;
            for (MM::iterator y = by_y.begin(); y != by_y.end(); y++) {
                if (y->second->end <= cur->end && y->second->score > cur->score) keep = false;
                if (y->second->end > cur->end) break;
            }
            if (!keep) continue;
            by_y.insert(std::make_pair(cur->end, cur));
            for (MM::iterator y = by_y.begin(); y != by_y.end(); y++)
                if (y->second->end >= cur->end && y->second->score < cur->score) y = by_y.erase(y);
        }
    }
    for (const ChainItem *a = by_y.rbegin()->second; a; a = a->pred) out.push_back(&a->rec);
    std::reverse(out.begin(), out.end());
    return out;
}

void append_record(std::string &out, const std::string &id, const Rec &r)      // print_one_TR, chaining.cpp:127-143
{
    char head[256];
    snprintf(head, sizeof head, "\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%f\t%d\t%d\t%d\t", r.inputLen, r.rep_start + 1, r.rep_end + 1,
             r.repeat_len, r.period, r.units, r.nm, (float)r.nm / r.repeat_len, r.nx, r.ni, r.nd);
    out += id; out += head;
    for (uint8_t b : r.unit) out += "ACGT"[b];
    out += '\n';
}

// pretty_print_alignment's output (wrap_around_DP.c:188-212) from the traceback ops of a PATH job
void append_alignment(std::string &out, const Rec &r, const uint8_t *org, const mtr_wdp_result &d, const uint8_t *path)
{
    const int n = d.path_len;
    std::string a(n, ' '), m(n, ' '), b(n, ' ');
    int i = d.max_i, j = d.max_j;
    if (j == 0) j = r.period;
    const uint8_t *x = org + r.rep_start - 1;              // rows are org[rep_start .. rep_end]
    for (int t = 0; t < n; t++) {
        switch (path[t]) {
        case 0: a[t] = "ACGT"[x[i]]; m[t] = '|'; b[t] = "ACGT"[r.unit[j - 1]]; i--; j--; break;
        case 1: a[t] = "ACGT"[x[i]]; b[t] = "ACGT"[r.unit[j - 1]]; i--; j--; break;
        case 2: a[t] = '-'; b[t] = "ACGT"[r.unit[j - 1]]; j--; break;
        default: a[t] = "ACGT"[x[i]]; b[t] = '-'; i--; break;
        }
        if (j == 0) j = r.period;
    }
    char head[128];
    snprintf(head, sizeof head, "\nmatch gain = %i, mismatch penalty = %i, indel penalty = %i\n\n", r.gain, r.mis, r.indel);
    out += head;
    for (int s = n - 1; s >= 0; s -= 50) {
        const int e = s - 50 >= -1 ? s - 50 : -1;
        for (int t = s; t > e; t--) out += a[t];
        out += '\n';
        for (int t = s; t > e; t--) out += m[t];
        out += '\n';
        for (int t = s; t > e; t--) out += b[t];
        out += "\n\n";
    }
}

// ---------------------------------------------------------------- per-read state machine
struct JobReq {
    int first, rows, unit_off, ulen;
    int8_t g[2], m[2], in[2];
    uint8_t n_param, mode;
    long long aux_need;         // CONSENSUS: int32 count, PATH: bytes
};

struct Chain {                  // one k of one candidate: find_tandem_repeat_sub (handle_one_read.c:77-100)
    enum Stage { UF_WAIT, SEARCH_WAIT, CONS_WAIT, DP_WAIT, DONE } stage = DONE;
    int uf_task = -1;
    int k = 0, pass = 0;
    Rec rr, tmp, dir[2];
    bool dir_found[2] = {false, false};
    int dir_job[2] = {-1, -1};
    bool found_last = false;
    float ratio0 = 0;
    int job = -1;
};

struct ReadState {
    std::string id;
    const uint8_t *org = nullptr;      // len + 2 bases (two stale tail bases, H4a)
    int L = 0, index = 0;
    int32_t *end = nullptr, *w = nullptr; // directional_index_end / _w of this read (views into the batch arrays); dead entries have end < 0
    int cursor = 0;
    struct Cand { int qs = 0, qe = 0; bool spec = false; long long cells = 0; std::vector<Chain> chains; };
    std::vector<Cand> cands;           // candidates in flight, in candidate order (front commits first)
    std::vector<ChainItem> accepted;
    std::vector<Rec> printing;         // -a: the chain waiting for its PATH jobs
    enum Phase { RUN, PRINT_WAIT, FINISHED } phase = RUN;
    // round I/O
    std::vector<JobReq> jobs;
    std::vector<uint8_t> units;
    std::vector<mtr_uf_task> uf_tasks; // unit-finder tasks of this round (K4)
    long long uf_base = 0;             // global index of this read's first unit-finder task of the previous round
    long long job_base = 0;            // global index of this read's first job of the previous round
    std::vector<long long> job_aux;    // byte offset in the round's aux buffer of each job of the previous round
    std::string out;
    long long candidates = 0;
    int print_job0 = 0;                // index of the first PATH job of `printing` among the round's jobs
    long long cells_total = 0, cells_wasted = 0;   // DP cells queued so far / by speculative candidates that were pruned after all
};

struct Worker { Counter cnt; WalkMemo memo; double t_build = 0, t_list = 0, t_walk = 0, t_polish = 0, t_step = 0; long long n_chain = 0, n_walk = 0;
                double hb_t[8] = {0}, hw_t[8] = {0}, hw_max[8] = {0}; long long hb_n[8] = {0}, hw_n[8] = {0}; };

int size_bucket(int n) { int b = 0; while (b < 7 && n > (64 << b)) b++; return b; }   // <=64, 128, ..., >4096

struct RoundResults {
    const mtr_wdp_result *res = nullptr;   // [2 * job]
    const uint8_t *aux = nullptr;
    const mtr_uf_result *uf = nullptr;     // [unit-finder task]
    const uint8_t *uf_units = nullptr;
    const int32_t *uf_scores = nullptr;
};

int g_speculate = 8;                       // MTR_SPECULATE: candidates a read may evaluate ahead of ones that could still prune them
                                           // (0: none, exactly the reference's order of evaluation; output identical either way)
bool g_uf_on_gpu = false;                  // MTR_UNITFINDER=gpu runs the unit finder as K4 on the GPU
int g_uf_gpu_min_window = 0;               // ... for candidate windows of at least this many bases (MTR_UF_GPU_MIN_WINDOW)

int add_job(ReadState &rs, int first, int rows, const std::vector<uint8_t> &unit, int n_param, const int (*p)[3], int mode)
{
    JobReq j;
    memset(&j, 0, sizeof j);
    j.first = first; j.rows = rows; j.unit_off = (int)rs.units.size(); j.ulen = (int)unit.size();
    for (int s = 0; s < n_param; s++) { j.g[s] = (int8_t)p[s][0]; j.m[s] = (int8_t)p[s][1]; j.in[s] = (int8_t)p[s][2]; }
    j.n_param = (uint8_t)n_param; j.mode = (uint8_t)mode;
    if ((long long)(j.ulen + 1) * (rows + 1) >= kWrapCap) {
        // wrap_around_DP.c:260-263: the reference aborts the whole run here
        fprintf(stderr, "You need to increse the value of WrapDPsize.\n");
        exit(EXIT_FAILURE);
    }
    j.aux_need = mode == MTR_TB_CONSENSUS ? (long long)(j.ulen + 1) * 9 : (mode == MTR_TB_PATH ? (long long)6 * rows + 64 : 0);
    rs.units.insert(rs.units.end(), unit.begin(), unit.end());
    rs.jobs.push_back(j);
    rs.cells_total += (long long)rows * j.ulen * n_param;
    return (int)rs.jobs.size() - 1;
}

const int kSearchParams[2][3] = {{1, 1, 3}, {1, 3, 1}};       // wrap_around_DP.c:395,405
const int kReviseParams[2][3] = {{5, 1, 1}, {1, 1, 3}};       // consensus.c:1062,1076

void emit_revise_cons(ReadState &rs, Chain &ch)
{
    ch.tmp = ch.rr;
    const int *p = kReviseParams[ch.pass];
    ch.tmp.gain = p[0]; ch.tmp.mis = p[1]; ch.tmp.indel = p[2];
    ch.job = add_job(rs, ch.tmp.rep_start, ch.tmp.rep_end - ch.tmp.rep_start + 1, ch.tmp.unit, 1, &kReviseParams[ch.pass], MTR_TB_CONSENSUS);
    ch.stage = Chain::CONS_WAIT;
}

// search_De_Bruijn_graph up to the point where it needs wrap_around_DP (consensus.c:507-549)
void start_chain(ReadState &rs, int qs, int qe, Chain &ch, Worker &wk)
{
    ch.rr.clear();
    ch.rr.inputLen = rs.L; ch.rr.kmer = ch.k;
    ch.dir_found[0] = ch.dir_found[1] = false;
    ch.found_last = false;
    if (g_uf_on_gpu && qe - qs + 1 >= g_uf_gpu_min_window) {
        mtr_uf_task t;
        t.read = rs.index; t.qs = qs; t.qe = qe; t.k = ch.k;
        ch.uf_task = (int)rs.uf_tasks.size();
        rs.uf_tasks.push_back(t);
        ch.stage = Chain::UF_WAIT;
        wk.n_chain++;
        return;
    }
    // the stage timers cost four clock reads per chain (8 M chains per 8192-read batch): only with MTR_PROFILE
    static const bool prof = getenv("MTR_PROFILE") != nullptr;
    double tp0 = prof ? now_s() : 0.0;
    wk.cnt.build(rs.org, rs.L, ch.k, qs, qe);
    double tp1 = prof ? now_s() : 0.0;
    wk.t_build += tp1 - tp0; wk.n_chain++;
    const double t_build_this = tp1 - tp0;
    const int maxf = wk.cnt.max_freq();
    int nodes[100];
    // the listing decrements counts (Q8), which only the walks can observe: skip it when they do not run (:532)
    const int nn = 5 < maxf ? wk.cnt.list_max_nodes(nodes, 100, maxf) : 0;
    tp0 = prof ? now_s() : 0.0;
    wk.t_list += tp0 - tp1;
    const int sb = size_bucket(qe - qs + 1);
    wk.hb_n[sb]++; wk.hb_t[sb] += t_build_this;
    struct WalkTimer { Worker &w; double t0; int sb; bool walked, on; ~WalkTimer() { if (!on) return; const double d = now_s() - t0; w.t_walk += d; if (walked) { w.hw_n[sb]++; w.hw_t[sb] += d; if (d > w.hw_max[sb]) w.hw_max[sb] = d; } } } walk_timer{wk, tp0, sb, 5 < maxf, prof};
    bool any = false;
    if (5 < maxf) {
        for (int d = 0; d < 2; d++) {
            wk.memo.reset();
            for (int i = 0; i < nn; i++) {
                Rec r = ch.rr;
                wk.n_walk++;
                const bool found = walk(wk.cnt, wk.memo, qs, qe, nodes[i], ch.k, d == 1, r);
                ch.found_last = found;
                if (!found) continue;
                ch.dir_job[d] = add_job(rs, qs, qe - qs + 1, r.unit, 2, kSearchParams, MTR_TB_COUNTS);
                ch.dir[d] = std::move(r); ch.dir_found[d] = true;
                any = true;
                break;
            }
        }
    }
    if (any) { ch.stage = Chain::SEARCH_WAIT; return; }
    ch.rr.clear();                                          // nothing found: find_tandem_repeat_sub clears (:86-88)
    ch.stage = Chain::DONE;
}

void advance_chain(ReadState &rs, int qs, int qe, Chain &ch, Worker &wk, const RoundResults &rr)
{
    const mtr_wdp_result *res = rr.res + 2 * rs.job_base;
    if (ch.stage == Chain::UF_WAIT) {
        const mtr_uf_result &u = rr.uf[rs.uf_base + ch.uf_task];
        ch.found_last = u.found_last != 0;
        bool any = false;
        for (int d = 0; d < 2; d++) {
            if (!u.found[d]) continue;
            Rec r = ch.rr;
            r.period = u.period[d];
            r.unit.assign(rr.uf_units + u.unit_off[d], rr.uf_units + u.unit_off[d] + u.period[d]);
            r.score.assign(rr.uf_scores + u.unit_off[d], rr.uf_scores + u.unit_off[d] + u.period[d]);
            ch.dir[d] = r; ch.dir_found[d] = true;
            ch.dir_job[d] = add_job(rs, qs, qe - qs + 1, r.unit, 2, kSearchParams, MTR_TB_COUNTS);
            any = true;
        }
        if (any) { ch.stage = Chain::SEARCH_WAIT; return; }
        ch.rr.clear();
        ch.stage = Chain::DONE;
        return;
    }
    if (ch.stage == Chain::SEARCH_WAIT) {
        // max_rr of search_De_Bruijn_graph starts cleared; wrap_around_DP (wrap_around_DP.c:357-429) keeps the strictly
        // better of the two penalty sets of a direction.  Everything the comparisons need is in the DP results, so the
        // record is materialised once, for the winner (a cleared record never qualifies: its Num_freq_unit is -1).
        int best_d = -1, best_s = -1;
        float best_ratio = -1;
        for (int d = 0; d < 2; d++) {
            if (!ch.dir_found[d]) continue;
            int pick_s = -1;
            float pick_ratio = -1;
            for (int s = 0; s < 2; s++) {
                const mtr_wdp_result &r = res[2 * ch.dir_job[d] + s];
                const float ratio = (float)r.n_match / (r.n_match + r.n_mismatch + r.n_ins + r.n_del);
                if (pick_ratio < ratio) { pick_s = s; pick_ratio = ratio; }
            }
            if (pick_s < 0) continue;
            const int period = ch.dir[d].period;
            const int units = res[2 * ch.dir_job[d] + pick_s].n_scanned / period;
            if (best_ratio < pick_ratio && min_match_ratio <= pick_ratio && 5 < units && 2 <= period && period < kMaxPeriod) {
                best_ratio = pick_ratio; best_d = d; best_s = pick_s;
            }
        }
        if (best_d >= 0) {
            ch.rr = std::move(ch.dir[best_d]);
            apply_dp(ch.rr, qs, res[2 * ch.dir_job[best_d] + best_s], kSearchParams[best_s][0], kSearchParams[best_s][1], kSearchParams[best_s][2]);
        } else {
            ch.rr.clear();
        }
        if (!ch.found_last) { ch.rr.clear(); ch.stage = Chain::DONE; return; }            // Q4
        if ((long long)ch.rr.period * (qe - qs + 1) > kWrapCap) {
            fprintf(stderr, "You need to increse the value of WrapDPsize.\n");
            ch.rr.clear(); ch.stage = Chain::DONE; return;
        }
        const int coverage = ch.rr.repeat_len / ch.rr.period;
        if (!(5 <= coverage && coverage <= 20 && 5 < ch.rr.period)) { ch.stage = Chain::DONE; return; }
        // revise_representative_unit (consensus.c:1048-1087)
        { static const bool prof = getenv("MTR_PROFILE") != nullptr;
          const double tq = prof ? now_s() : 0.0; polish(wk.cnt, rs.org, rs.L, ch.rr); if (prof) wk.t_polish += now_s() - tq; }
        ch.ratio0 = ch.rr.ratio();
        ch.pass = 0;
        emit_revise_cons(rs, ch);
        return;
    }
    if (ch.stage == Chain::CONS_WAIT) {
        const int *cons = (const int *)(rr.aux + rs.job_aux[ch.job]);
        vote_unit(ch.tmp, cons, cons + (size_t)(ch.tmp.period + 1) * 5);
        if (ch.tmp.period < kMaxPeriod) {
            if (ch.tmp.period <= 0) {                       // the reference divides by zero here (H9)
                fprintf(stderr, "mTR: the revised repeat unit is empty (read %s)\n", rs.id.c_str());
                exit(EXIT_FAILURE);
            }
            ch.job = add_job(rs, ch.tmp.rep_start, ch.tmp.rep_end - ch.tmp.rep_start + 1, ch.tmp.unit, 1, &kReviseParams[ch.pass], MTR_TB_COUNTS);
            ch.stage = Chain::DP_WAIT;
            return;
        }
    } else if (ch.stage == Chain::DP_WAIT) {
        const int *p = kReviseParams[ch.pass];
        apply_dp(ch.tmp, ch.tmp.rep_start, res[2 * ch.job], p[0], p[1], p[2]);
        if (ch.ratio0 < ch.tmp.ratio()) ch.rr = ch.tmp;     // ratio0 is never refreshed (Q10)
    }
    if (ch.pass == 0) { ch.pass = 1; emit_revise_cons(rs, ch); return; }
    ch.stage = Chain::DONE;
}

void finish_read(ReadState &rs, int print_alignment)
{
    std::vector<const Rec *> chain = best_chain(rs.accepted);
    if (!print_alignment) {
        for (const Rec *r : chain) append_record(rs.out, rs.id, *r);
        rs.phase = ReadState::FINISHED;
        return;
    }
    rs.printing.clear();
    for (const Rec *r : chain) rs.printing.push_back(*r);
    if (rs.printing.empty()) { rs.phase = ReadState::FINISHED; return; }
    rs.print_job0 = (int)rs.jobs.size();                    // jobs of a speculative candidate dropped in this very round may precede
    for (const Rec &r : rs.printing) {
        const int p[1][3] = {{r.gain, r.mis, r.indel}};
        add_job(rs, r.rep_start - 1, r.rep_end - r.rep_start + 1, r.unit, 1, p, MTR_TB_PATH);
    }
    rs.phase = ReadState::PRINT_WAIT;
}

// One round of one read: consume the results of the jobs it emitted last round, then run until it needs the
// GPU again (or is finished).  handle_one_TR's candidate loop, handle_one_read.c:227-246.
void step_read(ReadState &rs, Worker &wk, const RoundResults &rr, int print_alignment)
{
    rs.jobs.clear();
    rs.units.clear();
    rs.uf_tasks.clear();
    if (rs.phase == ReadState::PRINT_WAIT) {
        const mtr_wdp_result *res = rr.res + 2 * rs.job_base;
        for (size_t i = 0; i < rs.printing.size(); i++) {
            append_record(rs.out, rs.id, rs.printing[i]);
            append_alignment(rs.out, rs.printing[i], rs.org, res[2 * (rs.print_job0 + i)], rr.aux + rs.job_aux[rs.print_job0 + i]);
        }
        rs.phase = ReadState::FINISHED;
        return;
    }
    for (ReadState::Cand &cd : rs.cands) {
        const long long before = rs.cells_total;
        for (Chain &ch : cd.chains)
            if (ch.stage != Chain::DONE) advance_chain(rs, cd.qs, cd.qe, ch, wk, rr);
        cd.cells += rs.cells_total - before;
    }
    for (;;) {
        // commit finished candidates in candidate order
        while (!rs.cands.empty()) {
            ReadState::Cand &cd = rs.cands.front();
            bool done = true;
            for (const Chain &ch : cd.chains) if (ch.stage != Chain::DONE) { done = false; break; }
            if (!done) break;
            // find_tandem_repeat's pick over k (handle_one_read.c:135-146), then handle_one_TR's accept (:236-243)
            Rec pick;
            float best_ratio = -1;
            for (const Chain &ch : cd.chains) {
                const float ratio = ch.rr.ratio();
                if (best_ratio < ratio && min_match_ratio <= ratio && 5 < ch.rr.units && 2 <= ch.rr.period) {
                    best_ratio = ratio; pick = ch.rr;
                }
            }
            rs.candidates++;
            if (pick.repeat_len > 0 && pick.rep_start + 10 < pick.rep_end) {
                for (int i = pick.rep_start; i < pick.rep_end && i < rs.L; i++)       // :178-188
                    if (rs.end[i] >= 0 && rs.end[i] < pick.rep_end) { rs.end[i] = -1; rs.w[i] = -1; }
                ChainItem it;
                it.rec = pick; it.start = pick.rep_start; it.end = pick.rep_end; it.score = pick.nm; it.pred = nullptr;
                rs.accepted.push_back(it);
                // speculative candidates whose range this repeat has just pruned would never have been visited by the
                // reference: drop them (results of their jobs still in flight are simply never looked at)
                for (size_t c = 1; c < rs.cands.size();) {
                    if (rs.end[rs.cands[c].qs] < 0) { rs.cells_wasted += rs.cands[c].cells; rs.cands.erase(rs.cands.begin() + c); }
                    else c++;
                }
            }
            rs.cands.erase(rs.cands.begin());
        }
        // Start further candidates.  An accepted repeat of an in-flight candidate (qs', qe') ends at rep_end <= qe'+1
        // and prunes only ranges that end before rep_end (:181-182); a later candidate whose range ends beyond
        // every in-flight qe' can therefore never be pruned by them and is evaluated concurrently -- same results,
        // fewer sequential rounds.  Candidates that could still be pruned wait, exactly as in the reference.
        while (rs.cursor < rs.L && !(rs.end[rs.cursor] > -1 && rs.end[rs.cursor] < rs.L)) rs.cursor++;
        if (rs.cursor >= rs.L) {
            if (rs.cands.empty()) finish_read(rs, print_alignment);
            return;
        }
        const int qs = rs.cursor, qe = rs.end[rs.cursor];
        if ((int)rs.cands.size() >= kMaxInflightCands) return;
        bool safe = true;
        int n_spec = 0;
        for (const ReadState::Cand &cd : rs.cands) { if (cd.qe >= qe) safe = false; n_spec += cd.spec; }
        // A candidate that an in-flight one could still prune waits, exactly as in the reference -- or, with
        // MTR_SPECULATE = S, up to S of them are evaluated ahead: if the earlier candidate does accept a repeat that
        // prunes them they are dropped (wasted DP cells, counted apart), otherwise a dependent round has been saved.
        if (!safe && n_spec >= g_speculate) return;         // wait for the in-flight candidates (they have jobs queued)
        const int cw = rs.w[rs.cursor];
        rs.cursor++;
        int min_k, max_k;                                   // handle_one_read.c:105-120
        if (cw < 100) { min_k = 2; max_k = 10; } else if (cw < 1000) { min_k = 2; max_k = 12; } else { min_k = 5; max_k = 15; }
        rs.cands.emplace_back();
        ReadState::Cand &cd = rs.cands.back();
        cd.qs = qs; cd.qe = qe; cd.spec = !safe;
        cd.chains.resize(max_k - min_k + 1);
        const long long cells_before = rs.cells_total;
        // A k'-mer that occurs c times has a k-prefix (k < k') that occurs at least c times at the same coded
        // positions, so maxFreq(k') <= maxFreq(k) + (number of raw-base entries of the k' window, Q7).  Once that
        // bound is <= 5 the search cannot pass the maxFreq gate (consensus.c:532) for any larger k: those chains
        // end "not found" without building their count tables.
        int low_maxf = 1 << 30;
        for (int k = min_k; k <= max_k; k++) {
            Chain &ch = cd.chains[k - min_k];
            ch.k = k;
            const int raw = qe - std::min(qe, rs.L - k + 1) + 1;
            if (!g_uf_on_gpu && low_maxf + raw <= 5) {
                ch.rr.clear(); ch.found_last = false; ch.dir_found[0] = ch.dir_found[1] = false;
                ch.stage = Chain::DONE;
                continue;
            }
            start_chain(rs, qs, qe, ch, wk);
            if (!g_uf_on_gpu) low_maxf = std::min(low_maxf, wk.cnt.max_freq());
        }
        cd.cells += rs.cells_total - cells_before;
    }
}

// ---------------------------------------------------------------- a small persistent thread pool
class Pool {
public:
    explicit Pool(int n) : n_(std::max(1, n))
    {
        for (int t = 1; t < n_; t++) th_.emplace_back([this, t] { loop(t); });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void run(int count, const std::function<void(int, int)> &fn)
    {
        if (count <= 0) return;
        { std::lock_guard<std::mutex> g(m_); fn_ = &fn; count_ = count; next_ = 0; busy_ = n_ - 1; gen_++; }
        cv_.notify_all();
        work(0);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return busy_ == 0; });
    }
private:
    void work(int tid)
    {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= count_) break;
            (*fn_)(tid, i);
        }
    }
    void loop(int tid)
    {
        unsigned seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work(tid);
            { std::lock_guard<std::mutex> g(m_); busy_--; }
            done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int count_ = 0, busy_ = 0;
    unsigned gen_ = 0;
    bool stop_ = false;
};

// ---------------------------------------------------------------- one GPU + its host workers
struct ReadInput {
    std::string id;
    std::vector<uint8_t> bases;        // len + 2 (the two stale bases of H4a at the end)
    std::vector<uint16_t> stale;       // inputString_w_rand beyond len + 4r (H3)
    int len = 0;
};

[[noreturn]] void die(mtr_ctx *ctx, const char *what, int rc)
{
    fprintf(stderr, "mTR (B200): %s failed (%d): %s\n", what, rc, mtr_last_error(ctx));
    exit(EXIT_FAILURE);
}

struct Engine {
    mtr_ctx *ctx = nullptr;            // owns the resident reads; runs the directional index
    // DP dispatch lanes, grouped in tiers by the longest job (rows) a read queued this round: a batch is as slow as
    // its longest job (rows are sequential in the fill and in the traceback), and 7 of 8 read-rounds carry only jobs
    // of <= 256 rows.  Every lane is its own mtr_ctx (own streams and buffers) sharing the resident reads of ctx, so
    // all lanes run concurrently on the GPU.  The last tier takes everything.
    static constexpr int kMaxTiers = 6;
    int n_tiers = 3;
    int tier_rows[kMaxTiers] = {128, 1536, 1 << 30, 0, 0, 0};
    int tier_spin[kMaxTiers] = {0, 0, 0, 0, 0, 0};     // 1: the tier's dispatcher threads spin on the stream instead of sleeping
    std::vector<mtr_ctx *> tier_lanes[kMaxTiers];
    std::vector<mtr_ctx *> uf_lanes;   // unit finder (K4), opt-in
    Pool *pool = nullptr;
    std::vector<Worker> workers;
    double t_di = 0, t_dp = 0, t_rounds = 0;
    long long candidates = 0, rounds = 0, jobs_total = 0;
    std::atomic<int> unfinished{0}, batch_total{0};     // reads of the current batch still running (handle_one_file's stagger gate)
    std::atomic<long long> bases_left{0}, bases_total{0};   // ... and bases its reads have not scanned yet (reads all finish together
                                                            // under most-work-left-first scheduling; the scan position is the progress)

    static int parse_list(const char *e, int *out, int cap)
    {
        int n = 0;
        while (e && *e && n < cap) {
            out[n++] = atoi(e);
            e = strchr(e, ',');
            if (e) e++;
        }
        return n;
    }

    Engine(int device, int threads)
    {
        int rc = mtr_cuda_init(device, &ctx);
        if (rc) die(nullptr, "mtr_cuda_init", rc);
        if (const char *e = getenv("MTR_UNITFINDER")) g_uf_on_gpu = strcmp(e, "gpu") == 0;
        if (const char *e = getenv("MTR_UF_GPU_MIN_WINDOW")) g_uf_gpu_min_window = atoi(e);
        if (const char *e = getenv("MTR_SPECULATE")) g_speculate = std::max(0, atoi(e));
        int want_uf = g_uf_on_gpu ? 2 : 0;
        if (const char *e = getenv("MTR_UF_LANES")) want_uf = g_uf_on_gpu ? std::max(1, atoi(e)) : 0;
        // MTR_TIER_ROWS="128,1536": upper row bounds of all tiers but the last; MTR_TIER_LANES="2,2,2"; MTR_TIER_SPIN="0,0,0"
        int want[kMaxTiers] = {2, 2, 2, 2, 2, 2};
        if (const char *e = getenv("MTR_TIER_ROWS")) {
            int v[kMaxTiers];
            const int n = parse_list(e, v, kMaxTiers - 1);
            n_tiers = 0;
            for (int i = 0; i < n; i++) if (v[i] > 0 && (n_tiers == 0 || v[i] > tier_rows[n_tiers - 1])) tier_rows[n_tiers++] = v[i];
            tier_rows[n_tiers++] = 1 << 30;
        }
        if (const char *e = getenv("MTR_TIER_LANES")) { int v[kMaxTiers]; const int n = parse_list(e, v, kMaxTiers); for (int i = 0; i < n; i++) want[i] = std::max(1, v[i]); }
        if (const char *e = getenv("MTR_TIER_SPIN")) { int v[kMaxTiers]; const int n = parse_list(e, v, kMaxTiers); for (int i = 0; i < n; i++) tier_spin[i] = v[i] != 0; }
        for (int t = 0; t < n_tiers; t++)
            while ((int)tier_lanes[t].size() < want[t]) {
                mtr_ctx *c = nullptr;
                rc = mtr_cuda_init(device, &c);
                if (rc) die(nullptr, "mtr_cuda_init", rc);
                tier_lanes[t].push_back(c);
            }
        while ((int)uf_lanes.size() < want_uf) {
            mtr_ctx *c = nullptr;
            rc = mtr_cuda_init(device, &c);
            if (rc) die(nullptr, "mtr_cuda_init", rc);
            uf_lanes.push_back(c);
        }
        // a sleeping dispatcher leaves its core to the host workers (tens of microseconds of wake-up latency per call)
        for (int t = 0; t < n_tiers; t++)
            for (mtr_ctx *c : tier_lanes[t]) mtr_set_blocking_sync(c, tier_spin[t] ? 0 : 1);
        for (mtr_ctx *c : uf_lanes) mtr_set_blocking_sync(c, 1);
        mtr_set_blocking_sync(ctx, 1);
        // short-job lanes outrank long-job lanes, which outrank the directional index (MTR_TIER_PRIO=0: all equal)
        if (!getenv("MTR_TIER_PRIO") || atoi(getenv("MTR_TIER_PRIO")) != 0)
            for (int t = 0; t < n_tiers; t++)
                for (mtr_ctx *c : tier_lanes[t]) mtr_set_priority(c, n_tiers - t);
        pool = new Pool(threads);
        workers.resize(pool->size());
    }
    ~Engine()
    {
        delete pool;
        h_end.release(); h_w.release();
        for (int t = 0; t < n_tiers; t++)
            for (mtr_ctx *c : tier_lanes[t]) mtr_cuda_shutdown(c);
        for (mtr_ctx *c : uf_lanes) mtr_cuda_shutdown(c);
        mtr_cuda_shutdown(ctx);
    }

    // optional log of every DP job of the last run (bench: replay them as one batch to time K3 alone)
    bool log_jobs = false;
    std::vector<mtr_wdp_job> job_log;
    std::vector<uint8_t> unit_log;
    // resident batch (prepare) + statistics of the last run
    std::vector<int64_t> b_word_off, b_stale_off, b_pos_off;
    std::vector<uint16_t> b_stale;
    mtr_pipeline_stats ps = {};
    PinBuf h_end, h_w;

    // Processes one batch; returns the text the reference would have printed for these reads, in order.
    std::string process(std::vector<ReadInput> &in, int print_alignment)
    {
        prepare(in);
        return run(in, print_alignment);
    }

    // 2-bit packs the reads and uploads them: after this the batch is resident in HBM.
    void prepare(std::vector<ReadInput> &in)
    {
        const int n = (int)in.size();
        if (n == 0) return;
        b_word_off.assign(n + 1, 0); b_stale_off.assign(n + 1, 0); b_pos_off.assign(n + 1, 0);
        std::vector<int32_t> lens(n);
        for (int r = 0; r < n; r++) {
            lens[r] = in[r].len;
            const int64_t words = (in[r].len + 2 + 15) / 16;
            b_word_off[r + 1] = b_word_off[r] + ((words + 3) / 4) * 4;
            b_stale_off[r + 1] = b_stale_off[r] + (int64_t)in[r].stale.size();
            b_pos_off[r + 1] = b_pos_off[r] + in[r].len;
        }
        std::vector<uint32_t> packed((size_t)b_word_off[n], 0u);
        b_stale.assign((size_t)b_stale_off[n], 0);
        pool->run(n, [&](int, int r) {
            uint32_t *dst = packed.data() + b_word_off[r];
            const uint8_t *b = in[r].bases.data();
            const int nb = in[r].len + 2;
            for (int i = 0; i < nb; i++) dst[i >> 4] |= (uint32_t)b[i] << ((i & 15) * 2);
            if (!in[r].stale.empty()) memcpy(b_stale.data() + b_stale_off[r], in[r].stale.data(), in[r].stale.size() * 2);
        });
        int rc = mtr_reads_upload(ctx, packed.data(), b_word_off.data(), lens.data(), n);
        if (rc) die(ctx, "mtr_reads_upload", rc);
        std::vector<mtr_ctx *> others(uf_lanes);
        for (int t = 0; t < n_tiers; t++) others.insert(others.end(), tier_lanes[t].begin(), tier_lanes[t].end());
        for (mtr_ctx *c : others) {
            if (c == ctx) continue;
            rc = mtr_reads_share(c, ctx);
            if (rc) die(c, "mtr_reads_share", rc);
        }
        ps.h2d_bytes = (int64_t)packed.size() * 4 + (int64_t)(n + 1) * 12;
    }

    // Directional index + candidate rounds + chaining for the resident batch.
    std::string run(std::vector<ReadInput> &in, int print_alignment)
    {
        const int n = (int)in.size();
        std::string out;
        batch_total.store(n); unfinished.store(n);
        { long long tb = 0; for (const ReadInput &r : in) tb += r.len; bases_total.store(tb); bases_left.store(tb); }
        if (n == 0) return out;
        const std::vector<int64_t> &pos_off = b_pos_off;
        const int64_t h2d_prepare = ps.h2d_bytes;
        memset(&ps, 0, sizeof ps);
        ps.h2d_bytes = h2d_prepare;
        ps.reads = n; ps.bases = pos_off[n];
        job_log.clear(); unit_log.clear();
        // pinned, reused across batches: the D2H copy of end / w (8 bytes per base) runs at PCIe speed
        if (h_end.reserve(((size_t)pos_off[n] + 1) * 4) != cudaSuccess || h_w.reserve(((size_t)pos_off[n] + 1) * 4) != cudaSuccess) die(ctx, "cudaMallocHost", MTR_ENOMEM);
        int32_t *end = (int32_t *)h_end.p, *ww = (int32_t *)h_w.p;
        double t0 = now_s();
        // ---- per-read state machines
        std::vector<ReadState> st(n);
        for (int r = 0; r < n; r++) {
            ReadState &rs = st[r];
            rs.id = in[r].id; rs.org = in[r].bases.data(); rs.L = in[r].len; rs.index = r;
            rs.end = end + pos_off[r];
            rs.w = ww + pos_off[r];
        }
        // Asynchronous rounds: host workers advance whichever reads have their DP results, the dispatcher thread
        // sends everything queued so far to the GPU as soon as the previous batch is back.  No barrier: a read with
        // an expensive host step (a de Bruijn walk with many tie-breaks) delays only itself.
        const long long dir_cap = dir_budget();
        t0 = now_s();
        double host_ms = 0, wdp_ms = 0, uf_ms = 0;
        constexpr int UF = kMaxTiers;                             // index of the unit-finder queue / counters
        double lane_ms[kMaxTiers + 1] = {0}; long long lane_batches[kMaxTiers + 1] = {0}, lane_items[kMaxTiers + 1] = {0}, lane_reads[kMaxTiers + 1] = {0};
        struct BatchResult { std::vector<mtr_wdp_result> res; std::vector<uint8_t> aux; };
        struct UfBatchResult { std::vector<mtr_uf_result> res; std::vector<uint8_t> units; std::vector<int32_t> scores; };
        std::vector<std::shared_ptr<BatchResult>> result_of(n);
        std::vector<std::shared_ptr<UfBatchResult>> uf_result_of(n);
        std::vector<int> pending(n, 0);                           // lanes a read is still waiting for
        std::mutex mu;
        std::condition_variable cv_ready, cv_submit;
        std::vector<int> submitted[kMaxTiers + 1];                // per DP tier, [UF]: unit finder
        // Ready reads are served most-bases-left-to-scan first: every read is a chain of dependent rounds, so the batch
        // ends when its slowest chain does; serving the reads with the most work left first keeps them from being
        // starved behind reads that just came back (a LIFO stack left the cores idle at the end of every batch).
        typedef std::pair<int, int> ReadyKey;                     // (bases left, -index)
        std::priority_queue<ReadyKey> ready;
        auto push_ready = [&](int idx) { ready.push(ReadyKey(st[idx].L - st[idx].cursor, -idx)); };
        int remaining = n;
        const bool prof = getenv("MTR_PROFILE") != nullptr;
        const int worker_nice = getenv("MTR_WORKER_NICE") ? atoi(getenv("MTR_WORKER_NICE")) : 10;
        std::vector<double> finish_at(prof ? n : 0, 0.0);
        double idle_s = 0, worker_cpu_s = 0, disp_cpu_s = 0;
        int active_workers = 0;                                   // MTR_PROFILE timeline
        long long rows_hist_n[12] = {0}, rows_hist_cells[12] = {0}, readmax_hist[12] = {0};   // MTR_PROFILE: DP jobs by rows (<=32, 64, ...)
        double lane_fill_ms[kMaxTiers + 1] = {0}, lane_tb_ms[kMaxTiers + 1] = {0};
        auto worker_loop = [&](int tid) {
            // The workers are CPU-bound for the whole batch; the dispatcher threads wake up for microseconds at a time
            // and every microsecond they wait for a core delays a GPU round.  Raising the workers' nice value (needs no
            // privilege) lets the scheduler hand a waking dispatcher a core at once.
            if (worker_nice > 0) setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), worker_nice);
            const double cpu0 = prof ? thread_cpu_s() : 0.0;
            for (;;) {
                int idx;
                {
                    std::unique_lock<std::mutex> g(mu);
                    if (prof && ready.empty() && remaining != 0) {
                        const double ti = now_s();
                        cv_ready.wait(g, [&] { return !ready.empty() || remaining == 0; });
                        idle_s += now_s() - ti;
                    } else
                        cv_ready.wait(g, [&] { return !ready.empty() || remaining == 0; });
                    if (ready.empty()) { if (prof) worker_cpu_s += thread_cpu_s() - cpu0; return; }
                    idx = -ready.top().second;
                    ready.pop();
                }
                if (prof) { std::lock_guard<std::mutex> g(mu); active_workers++; }
                ReadState &rs = st[idx];
                RoundResults cur;
                if (result_of[idx]) { cur.res = result_of[idx]->res.data(); cur.aux = result_of[idx]->aux.data(); }
                if (uf_result_of[idx]) {
                    cur.uf = uf_result_of[idx]->res.data(); cur.uf_units = uf_result_of[idx]->units.data();
                    cur.uf_scores = uf_result_of[idx]->scores.data();
                }
                const double ts = prof ? now_s() : 0.0;
                const int left0 = rs.L - rs.cursor;
                step_read(rs, workers[tid], cur, print_alignment);
                bases_left.fetch_sub(left0 - (rs.phase == ReadState::FINISHED ? 0 : rs.L - rs.cursor), std::memory_order_relaxed);
                if (prof) workers[tid].t_step += now_s() - ts;
                result_of[idx].reset();
                uf_result_of[idx].reset();
                int max_rows = 0, lane = 0;
                for (const JobReq &q : rs.jobs) max_rows = std::max(max_rows, q.rows);
                while (max_rows > tier_rows[lane]) lane++;
                {
                    std::lock_guard<std::mutex> g(mu);
                    if (prof) active_workers--;
                    if (prof && rs.phase == ReadState::FINISHED) finish_at[idx] = now_s() - t0;
                    if (rs.phase == ReadState::FINISHED) unfinished.store(remaining - 1, std::memory_order_relaxed);
                    if (rs.phase == ReadState::FINISHED) { if (--remaining == 0) { cv_ready.notify_all(); cv_submit.notify_all(); } }
                    else {
                        if (!rs.jobs.empty()) { pending[idx]++; submitted[lane].push_back(idx); }
                        if (!rs.uf_tasks.empty()) { pending[idx]++; submitted[UF].push_back(idx); }
                        if (pending[idx] == 0) { fprintf(stderr, "mTR (B200): internal error: read %d stalled\n", idx); exit(EXIT_FAILURE); }
                        cv_submit.notify_all();
                    }
                }
            }
        };
        auto dispatch_loop = [&](int lane, mtr_ctx *lctx) {
            cudaSetDevice(lctx->device);
            std::vector<mtr_wdp_job> jobs;
            std::vector<uint8_t> units;
            std::vector<int> batch;
            const double cpu0 = prof ? thread_cpu_s() : 0.0;
            struct CpuAcc { double &acc; double c0; bool on; std::mutex &m; ~CpuAcc() { if (on) { std::lock_guard<std::mutex> g(m); acc += thread_cpu_s() - c0; } } } cpu_acc{disp_cpu_s, cpu0, prof, mu};
            for (;;) {
                {
                    std::unique_lock<std::mutex> g(mu);
                    cv_submit.wait(g, [&] { return !submitted[lane].empty() || remaining == 0; });
                    if (submitted[lane].empty()) return;
                    // a long queue (the first round of a batch: every read arrives at once) is shared with the other
                    // lanes of the tier instead of going out as one huge launch that everybody waits for
                    std::vector<int> &q = submitted[lane];
                    const size_t lanes_here = tier_lanes[lane].size();
                    const size_t take = q.size() <= 96 ? q.size() : std::max<size_t>(64, (q.size() + lanes_here - 1) / lanes_here);
                    if (take >= q.size()) { batch.swap(q); q.clear(); }
                    else {
                        batch.assign(q.begin(), q.begin() + take);
                        q.erase(q.begin(), q.begin() + take);
                        cv_submit.notify_all();                 // the rest is for a sibling lane
                    }
                }
                jobs.clear(); units.clear();
                long long aux_bytes = 0;
                for (int idx : batch) {
                    ReadState &rs = st[idx];
                    rs.job_base = (long long)jobs.size();
                    rs.job_aux.assign(rs.jobs.size(), 0);
                    for (size_t j = 0; j < rs.jobs.size(); j++) {
                        const JobReq &q = rs.jobs[j];
                        mtr_wdp_job gj;
                        memset(&gj, 0, sizeof gj);
                        gj.read = rs.index; gj.first = q.first; gj.rows = q.rows; gj.ulen = q.ulen;
                        gj.unit_off = (int32_t)(units.size() + (size_t)q.unit_off);
                        for (int s = 0; s < 2; s++) { gj.gain[s] = q.g[s]; gj.mis[s] = q.m[s]; gj.indel[s] = q.in[s]; }
                        gj.n_param = q.n_param; gj.mode = q.mode;
                        if (q.mode != MTR_TB_COUNTS) {
                            aux_bytes = (aux_bytes + 15) & ~15LL;
                            rs.job_aux[j] = aux_bytes;
                            gj.aux_off = q.mode == MTR_TB_CONSENSUS ? aux_bytes / 4 : aux_bytes;
                            gj.aux_cap = q.aux_need;
                            aux_bytes += q.mode == MTR_TB_CONSENSUS ? q.aux_need * 4 : q.aux_need;
                        }
                        jobs.push_back(gj);
                    }
                    units.insert(units.end(), rs.units.begin(), rs.units.end());
                }
                auto br = std::make_shared<BatchResult>();
                br->res.resize(jobs.size() * 2);
                br->aux.assign((size_t)aux_bytes, 0);
                const double tg0 = now_s();
                mtr_pipeline_stats local;
                memset(&local, 0, sizeof local);
                run_jobs(lctx, jobs, units, br->res, br->aux, dir_cap, local);
                {
                    std::lock_guard<std::mutex> g(mu);
                    wdp_ms += (now_s() - tg0) * 1e3;
                    lane_fill_ms[lane] += local.wdp_fill_ms; lane_tb_ms[lane] += local.wdp_tb_ms;
                    if (prof) {
                        for (const mtr_wdp_job &j : jobs) { int b = 0; while (b < 11 && j.rows > (32 << b)) b++; rows_hist_n[b]++; rows_hist_cells[b] += (long long)j.rows * j.ulen * j.n_param; }
                        for (int idx : batch) { int mx = 0; for (const JobReq &q : st[idx].jobs) mx = std::max(mx, q.rows); int b = 0; while (b < 11 && mx > (32 << b)) b++; readmax_hist[b]++; }
                    }
                    lane_ms[lane] += (now_s() - tg0) * 1e3; lane_batches[lane]++; lane_items[lane] += (long long)jobs.size(); lane_reads[lane] += (long long)batch.size();
                    if (log_jobs) {
                        const int32_t shift = (int32_t)unit_log.size();
                        for (mtr_wdp_job j : jobs) { j.unit_off += shift; j.mode = MTR_TB_COUNTS; j.aux_off = 0; j.aux_cap = 0; job_log.push_back(j); }
                        unit_log.insert(unit_log.end(), units.begin(), units.end());
                    }
                    rounds++; jobs_total += (long long)jobs.size();
                    ps.rounds++; ps.jobs += (int64_t)jobs.size();
                    if (lane < n_tiers - 1) ps.rounds_fast++;
                    ps.h2d_bytes += (int64_t)jobs.size() * sizeof(mtr_wdp_job) + (int64_t)units.size();
                    ps.d2h_bytes += (int64_t)jobs.size() * 2 * sizeof(mtr_wdp_result) + aux_bytes;
                    ps.wdp_fill_ms += local.wdp_fill_ms; ps.wdp_tb_ms += local.wdp_tb_ms; ps.wdp_cells += local.wdp_cells;
                    ps.wdp_slot_cells += local.wdp_slot_cells; ps.wdp_dir_bytes += local.wdp_dir_bytes;
                    ps.launches += local.launches; ps.wdp_calls += local.wdp_calls;
                    for (int idx : batch) { result_of[idx] = br; if (--pending[idx] == 0) push_ready(idx); }
                }
                cv_ready.notify_all();
                batch.clear();
            }
        };
        auto uf_dispatch_loop = [&](mtr_ctx *ctx_uf) {
            cudaSetDevice(ctx_uf->device);
            std::vector<mtr_uf_task> tasks;
            std::vector<int> batch;
            for (;;) {
                {
                    std::unique_lock<std::mutex> g(mu);
                    cv_submit.wait(g, [&] { return !submitted[UF].empty() || remaining == 0; });
                    if (submitted[UF].empty()) return;
                    batch.swap(submitted[UF]);
                    submitted[UF].clear();
                }
                tasks.clear();
                long long cap = 16;
                for (int idx : batch) {
                    ReadState &rs = st[idx];
                    rs.uf_base = (long long)tasks.size();
                    for (const mtr_uf_task &t : rs.uf_tasks) { tasks.push_back(t); cap += 2LL * std::min(kMaxPeriod, (t.qe - t.qs) / 5); }
                }
                auto br = std::make_shared<UfBatchResult>();
                br->res.resize(tasks.size());
                br->units.resize((size_t)cap);
                br->scores.resize((size_t)cap);
                int64_t used = 0;
                const double tg0 = now_s();
                const int rc = mtr_uf_run(ctx_uf, tasks.data(), (int)tasks.size(), br->res.data(), br->units.data(), br->scores.data(), cap, &used);
                if (rc) die(ctx_uf, "mtr_uf_run", rc);
                mtr_stats us;
                mtr_get_stats(ctx_uf, &us);
                {
                    std::lock_guard<std::mutex> g(mu);
                    uf_ms += (now_s() - tg0) * 1e3;
                    lane_ms[UF] += (now_s() - tg0) * 1e3; lane_batches[UF]++; lane_items[UF] += (long long)tasks.size(); lane_reads[UF] += (long long)batch.size();
                    ps.rounds_uf++; ps.uf_tasks += (int64_t)tasks.size(); ps.uf_kernel_ms += us.uf_ms; ps.launches += 1;
                    ps.h2d_bytes += (int64_t)tasks.size() * sizeof(mtr_uf_task);
                    ps.d2h_bytes += (int64_t)tasks.size() * sizeof(mtr_uf_result) + used * 5;
                    for (int idx : batch) { uf_result_of[idx] = br; if (--pending[idx] == 0) push_ready(idx); }
                }
                cv_ready.notify_all();
                batch.clear();
            }
        };
        // The directional index sweeps the batch in slices on its own context; the reads of a slice join the ready
        // queue as soon as their candidate ranges are back, so the host workers and the DP lanes start while the
        // later slices are still on the GPU.
        // (default: one slice.  Measured on B200: in 512-read slices next to the busy DP lanes the index kernels take 4x
        // longer in total and the last reads enter 0.45 s late, which costs more than the 0.1 s of overlap gains.)
        const int di_slice = std::max(1, getenv("MTR_DI_SLICE") ? atoi(getenv("MTR_DI_SLICE")) : (1 << 30));
        std::thread di_thread([&] {
            cudaSetDevice(ctx->device);
            const double td0 = now_s();
            for (int first = 0; first < n; first += di_slice) {
                const int count = std::min(di_slice, n - first);
                const int rc = mtr_di_run_range(ctx, Manhattan_Distance, b_stale.data(), b_stale_off.data(), pos_off.data(), nullptr, end, ww, first, count);
                if (rc) die(ctx, "mtr_di_run", rc);
                mtr_stats s;
                mtr_get_stats(ctx, &s);
                {
                    std::lock_guard<std::mutex> g(mu);
                    ps.di_kernel_ms += s.di_ms; ps.di_position_passes += s.di_position_passes; ps.launches += s.launches;
                    ps.di_bytes_in += s.di_bytes_in; ps.di_bytes_out += s.di_bytes_out;
                    for (int r = first; r < first + count; r++) push_ready(r);
                }
                cv_ready.notify_all();
            }
            std::lock_guard<std::mutex> g(mu);
            ps.h2d_bytes += (int64_t)b_stale.size() * 2; ps.d2h_bytes += pos_off[n] * 8;
            ps.di_wall_ms = (now_s() - td0) * 1e3;
            t_di += now_s() - td0;
        });
        std::thread monitor;
        if (prof)
            monitor = std::thread([&] {
                for (;;) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(100));
                    std::lock_guard<std::mutex> g(mu);
                    if (remaining == 0) return;
                    size_t q[kMaxTiers] = {0};
                    for (int t = 0; t < n_tiers; t++) q[t] = submitted[t].size();
                    fprintf(stderr, "[mtr timeline] t %.2f s: unfinished %5d  ready %5zu  workers busy %2d  queued for tiers %zu/%zu/%zu/%zu\n", now_s() - t0, remaining,
                            ready.size(), active_workers, q[0], q[1], q[2], q[3]);
                }
            });
        std::vector<std::thread> dispatchers;
        for (int t = 0; t < n_tiers; t++)
            for (mtr_ctx *c : tier_lanes[t]) dispatchers.emplace_back([&, t, c] { dispatch_loop(t, c); });
        for (mtr_ctx *c : uf_lanes) dispatchers.emplace_back([&, c] { uf_dispatch_loop(c); });
        {
            const double th0 = now_s();
            // fresh threads, not the pool: the nice value of a thread cannot be lowered again without privilege, and
            // the caller's thread must not be touched
            std::vector<std::thread> wt;
            const int nw = pool->size();
            for (int t = 0; t < nw; t++) wt.emplace_back([&, t] { worker_loop(t); });
            for (std::thread &t : wt) t.join();
            host_ms += (now_s() - th0) * 1e3;
        }
        for (std::thread &t : dispatchers) t.join();
        di_thread.join();
        if (monitor.joinable()) monitor.join();
        t_rounds += now_s() - t0;
        t_dp += wdp_ms / 1e3;
        ps.rounds_wall_ms = (now_s() - t0) * 1e3;
        for (int r = 0; r < n; r++) { out += st[r].out; candidates += st[r].candidates; ps.candidates += st[r].candidates; ps.spec_cells += st[r].cells_wasted; }
        ps.host_step_ms = host_ms; ps.wdp_wall_ms = wdp_ms; ps.uf_wall_ms = uf_ms;
        if (getenv("MTR_PROFILE")) {
            double b = 0, l = 0, w = 0, p = 0, t = 0; long long nc = 0, nw = 0;
            for (Worker &k : workers) { b += k.t_build; l += k.t_list; w += k.t_walk; p += k.t_polish; t += k.t_step; nc += k.n_chain; nw += k.n_walk;
                                        k.t_build = k.t_list = k.t_walk = k.t_polish = k.t_step = 0; k.n_chain = k.n_walk = 0; }
            for (int l = 0; l <= UF; l++) {
                if (l >= n_tiers && l != UF) continue;
                char name[64];
                if (l == UF) snprintf(name, sizeof name, "unit finder");
                else if (l == n_tiers - 1) snprintf(name, sizeof name, "DP, longer jobs");
                else snprintf(name, sizeof name, "DP rows <= %d", tier_rows[l]);
                fprintf(stderr, "[mtr profile]   tier %d (%s, %d lanes): batches %6lld  busy %8.1f ms  items/batch %8.1f  reads/batch %6.1f  ms/batch %6.3f  kernel ms: fill %.1f traceback %.1f\n", l, name,
                        l == UF ? (int)uf_lanes.size() : (int)tier_lanes[l].size(), lane_batches[l], lane_ms[l],
                        lane_batches[l] ? (double)lane_items[l] / lane_batches[l] : 0.0, lane_batches[l] ? (double)lane_reads[l] / lane_batches[l] : 0.0,
                        lane_batches[l] ? lane_ms[l] / lane_batches[l] : 0.0, lane_fill_ms[l], lane_tb_ms[l]);
            }
            for (int b = 0; b < 12; b++)
                fprintf(stderr, "[mtr profile]   DP jobs with rows <= %6d: %9lld jobs %8.2f Gcells | read-rounds whose longest job is in this bucket: %8lld\n", 32 << b, rows_hist_n[b], rows_hist_cells[b] / 1e9, readmax_hist[b]);
            for (int b = 0; b < 8; b++) {
                long long bn = 0, wn = 0; double wt = 0, wm = 0, bt = 0;
                for (Worker &k : workers) { bn += k.hb_n[b]; bt += k.hb_t[b]; wn += k.hw_n[b]; wt += k.hw_t[b]; wm = std::max(wm, k.hw_max[b]); k.hb_n[b] = k.hw_n[b] = 0; k.hb_t[b] = k.hw_t[b] = k.hw_max[b] = 0; }
                fprintf(stderr, "[mtr profile]   window <= %5d: chains %9lld  table cpu-s %8.3f  with walks %8lld  walk cpu-s %8.3f  max walk ms %8.2f\n", 64 << b, bn, bt, wn, wt, wm * 1e3);
            }
            {
                std::vector<double> f = finish_at;
                std::sort(f.begin(), f.end());
                auto q = [&](double x) { return f.empty() ? 0.0 : f[std::min(f.size() - 1, (size_t)(x * f.size()))]; };
                fprintf(stderr, "[mtr profile]   read finish times s: p10 %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f | worker idle %.3f s  worker cpu %.3f s  dispatcher cpu %.3f s\n",
                        q(0.10), q(0.50), q(0.90), q(0.99), f.empty() ? 0.0 : f.back(), idle_s, worker_cpu_s, disp_cpu_s);
            }
            fprintf(stderr, "[mtr profile] reads %d rounds %lld | host cpu-s: step %.3f build %.3f maxlist %.3f walk %.3f polish %.3f | chains %lld walks %lld | wall: host %.3f s gpu %.3f s\n",
                    n, (long long)ps.rounds, t, b, l, w, p, nc, nw, host_ms / 1e3, wdp_ms / 1e3);
        }
        return out;
    }

    static long long dir_budget()
    {
        const char *e = getenv("MTR_DIR_BUDGET_MB");
        return (e ? atoll(e) : 8192LL) << 20;
    }

    // Runs the round's jobs in sub-batches whose direction matrices fit the budget.
    void run_jobs(mtr_ctx *lctx, std::vector<mtr_wdp_job> &jobs, std::vector<uint8_t> &units, std::vector<mtr_wdp_result> &res,
                  std::vector<uint8_t> &aux, long long dir_cap, mtr_pipeline_stats &acc)
    {
        size_t a = 0;
        std::vector<mtr_wdp_job> part;
        while (a < jobs.size()) {
            size_t b = a;
            long long dir = 0, aux_lo = -1, aux_hi = 0;
            while (b < jobs.size()) {
                const long long need = wdp_dir_bytes(jobs[b].ulen, jobs[b].rows) * jobs[b].n_param;
                if (b > a && dir + need > dir_cap) break;
                dir += need;
                if (jobs[b].mode != MTR_TB_COUNTS) {
                    const long long lo = jobs[b].mode == MTR_TB_CONSENSUS ? jobs[b].aux_off * 4 : jobs[b].aux_off;
                    const long long sz = jobs[b].mode == MTR_TB_CONSENSUS ? jobs[b].aux_cap * 4 : jobs[b].aux_cap;
                    if (aux_lo < 0) aux_lo = lo;
                    aux_hi = lo + sz;
                }
                b++;
            }
            part.assign(jobs.begin() + a, jobs.begin() + b);
            if (aux_lo > 0)
                for (mtr_wdp_job &j : part)
                    if (j.mode != MTR_TB_COUNTS) j.aux_off -= j.mode == MTR_TB_CONSENSUS ? aux_lo / 4 : aux_lo;
            const double t0 = now_s();
            const int rc = mtr_wdp_run(lctx, part.data(), (int)part.size(), units.data(), (int64_t)units.size(), res.data() + 2 * a,
                                       aux_lo >= 0 ? aux.data() + aux_lo : nullptr, aux_lo >= 0 ? aux_hi - aux_lo : 0);
            if (rc) die(lctx, "mtr_wdp_run", rc);
            (void)t0;
            {
                mtr_stats s;
                mtr_get_stats(lctx, &s);
                acc.wdp_fill_ms += s.wdp_fill_ms; acc.wdp_tb_ms += s.wdp_tb_ms; acc.wdp_cells += s.wdp_cells;
                acc.wdp_slot_cells += s.wdp_slot_cells; acc.wdp_dir_bytes += s.wdp_dir_bytes; acc.launches += s.launches;
                acc.wdp_calls++;
            }
            a = b;
        }
    }
};

// ---------------------------------------------------------------- cross-read stale state
// The reference never clears orgInputString / inputString_w_rand between reads, and reads past the part it
// rewrites.  Both effects depend only on the sequence of reads, so they are reproduced while parsing: every
// read gets the two bases beyond its end and the k = 5 coded values beyond len + 4r that an earlier, longer
// read left behind (zeros in a fresh process).
struct StaleTracker {
    std::vector<uint8_t> org;          // shadow of orgInputString
    std::vector<uint16_t> padded;      // shadow of inputString_w_rand after the k = 5 pass
    std::vector<uint8_t> mt;           // genrand_int32() % 4 after init_genrand(0)
    StaleTracker() : org(kMaxLen + 8, 0), padded(3 * (size_t)kMaxLen, 0)
    {
        mt.resize(1300000);
        uint32_t s[624];
        s[0] = 0;
        for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        int pos = 624;
        for (size_t t = 0; t < mt.size(); t++) {
            if (pos == 624) {
                for (int i = 0; i < 624; i++) {
                    const uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
                    s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
                }
                pos = 0;
            }
            uint32_t y = s[pos++];
            y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
            mt[t] = (uint8_t)(y & 3u);
        }
    }
    struct Geom { int L, r, N, ext, need; };
    static Geom geom(int L)
    {
        Geom g;
        g.L = L; g.r = L < 1000 ? 100 : L / 10; g.N = L + 2 * g.r;
        g.ext = std::min(L + 4 * g.r, kMaxLen);             // area the read re-initialises (:143)
        int wmax = 0;
        for (int w = 5; w <= 10240 && w < L / 2; w *= 2) wmax = w;
        g.need = L + g.r + 2 * wmax + 8;                    // last index the k = 5 passes touch
        return g;
    }
    // inputString_w_rand[x] (x < ext) after init_inputString_surrounded_by_random_seq(k = 5) of this read
    // (fill_directional_index.c:137-169): 5-mer code below N-4, raw padded base above
    int s5_at(const ReadInput &in, const Geom &g, int x) const
    {
        auto base = [&](int i) -> int {
            if (i < g.r) return mt[g.ext + i];
            if (i < g.r + g.L) return in.bases[i - g.r];
            if (i < g.N) return mt[g.ext + g.r + (i - g.r - g.L)];
            return mt[i];
        };
        if (x >= g.N - 4) return base(x);
        return (((base(x) * 4 + base(x + 1)) * 4 + base(x + 2)) * 4 + base(x + 3)) * 4 + base(x + 4);
    }
    // Processes reads v[b..e) as they follow each other in the input: every read gets the two bases past its end
    // and the k = 5 codes beyond its own area that the latest earlier, longer read left behind -- looked up
    // directly in the earlier reads of the batch (in parallel over reads) or in the state carried over from
    // earlier batches; afterwards the carried state is advanced past the batch in O(longest read).
    void visit_batch(std::vector<ReadInput> &v, size_t b, size_t e, int threads, const int *tail_override)
    {
        const int n = (int)(e - b);
        if (n <= 0) return;
        std::vector<Geom> g(n);
        for (int i = 0; i < n; i++) { g[i] = geom(v[b + i].len); v[b + i].bases.resize(v[b + i].len + 2); }
        auto work = [&](int i) {
            ReadInput &in = v[b + i];
            const int L = g[i].L;
            for (int t = 0; t < 2; t++) {
                const int x = L + t;
                int val = org[x];
                for (int m = i - 1; m >= 0; m--) if (g[m].L > x) { val = v[b + m].bases[x]; break; }
                in.bases[x] = tail_override && n == 1 ? (uint8_t)(tail_override[t] & 3) : (uint8_t)val;
            }
            const int lo = g[i].ext, hi = g[i].need;
            in.stale.assign(hi > lo ? hi - lo : 0, 0);
            if (hi <= lo) return;
            int maxext = 0;
            for (int m = i - 1; m >= 0 && maxext < hi; m--) {
                if (g[m].ext <= maxext) continue;
                for (int x = std::max(lo, maxext); x < std::min(g[m].ext, hi); x++) in.stale[x - lo] = (uint16_t)s5_at(v[b + m], g[m], x);
                maxext = g[m].ext;
            }
            for (int x = std::max(lo, maxext); x < hi; x++) in.stale[x - lo] = padded[x];
        };
        const int T = std::max(1, std::min(threads, n / 8));
        if (T <= 1) { for (int i = 0; i < n; i++) work(i); }
        else {
            std::atomic<int> next{0};
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back([&] { for (int i; (i = next.fetch_add(1)) < n;) work(i); });
            for (auto &x : th) x.join();
        }
        int maxext = 0, maxlen = 0;
        for (int i = n - 1; i >= 0; i--) {
            if (g[i].ext > maxext) {
                for (int x = maxext; x < g[i].ext; x++) padded[x] = (uint16_t)s5_at(v[b + i], g[i], x);
                maxext = g[i].ext;
            }
            if (g[i].L > maxlen) { memcpy(org.data() + maxlen, v[b + i].bases.data() + maxlen, (size_t)(g[i].L - maxlen)); maxlen = g[i].L; }
        }
    }
};

// ---------------------------------------------------------------- FASTA reader (handle_one_file.c:169-269)
struct FastaReader {
    FILE *fp = nullptr;
    std::string next_id;
    bool started = false, eof = false;
    std::vector<char> buf;
    explicit FastaReader(const char *path) : buf(1 << 20)
    {
        fp = fopen(path, "r");
        if (!fp) { fprintf(stderr, "fatal error: cannot open %s\n", path); fflush(stderr); exit(EXIT_FAILURE); }
    }
    ~FastaReader() { if (fp) fclose(fp); }
    // Returns false at the end of input or at a zero-length read (which ends the run in the reference, :283).
    bool next(ReadInput &out)
    {
        if (eof) return false;
        out.id.clear(); out.bases.clear(); out.len = 0;
        bool any = false;
        while (fgets(buf.data(), (int)buf.size(), fp)) {
            any = true;
            char *s = buf.data();
            if (s[0] == '>' ) {
                std::string id;
                for (int i = 1; s[i] && s[i] != '\n' && s[i] != '\r'; i++) id += s[i];
                if (!started) { started = true; next_id = id; continue; }
                out.id = next_id; next_id = id;
                out.len = (int)out.bases.size();
                return out.len > 0;
            }
            // one line of bases: table lookup into the tail of out.bases (same checks, in the same order, as the
            // per-character loop of handle_one_file.c:169-188,240-262)
            static const struct Lut { int8_t v[256]; Lut() { memset(v, -1, sizeof v); v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; } } lut;
            const size_t m = strcspn(s, "\r\n");
            const size_t old = out.bases.size();
            const size_t lim = std::min(m, (size_t)kMaxLen - old);
            out.bases.resize(old + lim);
            uint8_t *dst = out.bases.data() + old;
            for (size_t i = 0; i < lim; i++) {
                const int8_t b = lut.v[(unsigned char)s[i]];
                if (b < 0) { fprintf(stderr, "Invalid character: %c \n", s[i]); exit(EXIT_FAILURE); }
                dst[i] = (uint8_t)b;
            }
            if (kMaxLen <= (int)out.bases.size()) {
                fprintf(stderr, "fatal error: The length %d is tentatively at most %i.\nread ID = %s\nSet MAX_INPUT_LENGTH to a larger value",
                        (int)out.bases.size(), kMaxLen, out.id.c_str());
                fprintf(stderr, "cannot allocate space for one of global variables in the heap.\n");
                exit(EXIT_FAILURE);
            }
        }
        eof = true;
        if (!any) return false;
        out.id = next_id;
        out.len = (int)out.bases.size();
        return out.len > 0;
    }
};

// ---------------------------------------------------------------- process-wide state behind the C entry points
struct Runtime {
    std::vector<Engine *> engines;     // per_gpu engines for every GPU: engine e works on GPU e / per_gpu
    int per_gpu = 2;                   // engines per GPU (MTR_INFLIGHT_PER_GPU): the next batch starts on the sibling engine when
                                       // the running one is down to its last reads (stagger_frac), so the ramp-down of one batch
                                       // (a few chains of dependent rounds, cores idle) overlaps the ramp-up of the next.  Two
                                       // batches in full flight would only slow each other's long-job lanes down.
    double stagger_frac = 0.12;        // MTR_STAGGER_FRAC
    bool stagger_by_bases = true;      // the gate looks at the bases left to scan (MTR_STAGGER_BY=reads: at the reads left, which
                                       // under most-work-left-first scheduling all finish in the last 0.1 s of a batch)
    StaleTracker stale;
    std::vector<ReadInput> pending;
    long long pending_bases = 0;
    int batch_reads = 8192;            // bigger batches amortise the ramp-up / ramp-down of a batch
    int prep_threads = 4;
    long long batch_bases = 192LL << 20;
    int print_alignment = 0;
    mtr_pipeline_stats totals = {};    // summed over the batches of handle_one_file / mtr_flush since the last mtr_file_stats

    void account(const mtr_pipeline_stats &p)
    {
        totals.reads += p.reads; totals.bases += p.bases; totals.candidates += p.candidates; totals.rounds += p.rounds;
        totals.rounds_fast += p.rounds_fast; totals.jobs += p.jobs; totals.wdp_calls += p.wdp_calls; totals.wdp_cells += p.wdp_cells;
        totals.wdp_slot_cells += p.wdp_slot_cells; totals.wdp_dir_bytes += p.wdp_dir_bytes; totals.di_position_passes += p.di_position_passes;
        totals.di_bytes_in += p.di_bytes_in; totals.di_bytes_out += p.di_bytes_out; totals.h2d_bytes += p.h2d_bytes; totals.d2h_bytes += p.d2h_bytes;
        totals.launches += p.launches; totals.spec_cells += p.spec_cells; totals.wdp_fill_ms += p.wdp_fill_ms; totals.wdp_tb_ms += p.wdp_tb_ms; totals.di_kernel_ms += p.di_kernel_ms;
        totals.di_wall_ms += p.di_wall_ms; totals.rounds_wall_ms += p.rounds_wall_ms; totals.host_step_ms += p.host_step_ms; totals.wdp_wall_ms += p.wdp_wall_ms;
    }

    Runtime()
    {
        int ngpu = 1;
        if (const char *e = getenv("MTR_GPUS")) ngpu = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_BATCH_READS")) batch_reads = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_BATCH_MBASES")) batch_bases = std::max(1LL, atoll(e)) << 20;
        int base = 0;
        if (const char *e = getenv("MTR_DEVICE")) base = atoi(e);
        int threads = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("MTR_THREADS")) threads = atoi(e);
        if (const char *e = getenv("MTR_INFLIGHT_PER_GPU")) per_gpu = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_STAGGER_FRAC")) stagger_frac = atof(e);
        if (const char *e = getenv("MTR_STAGGER_BY")) stagger_by_bases = strcmp(e, "reads") != 0;
        prep_threads = std::max(1, std::min(8, threads / 2));
        threads = std::max(1, threads / ngpu);
        for (int g = 0; g < ngpu; g++)
            for (int i = 0; i < per_gpu; i++) engines.push_back(new Engine(base + g, threads));
    }
    ~Runtime() { for (Engine *e : engines) delete e; }
};

Runtime *g_rt = nullptr;
Runtime &runtime()
{
    if (!g_rt) g_rt = new Runtime();
    return *g_rt;
}

void publish_timers(Runtime &rt)
{
    for (Engine *e : rt.engines) {
        time_range += (float)e->t_di; time_wrap_around_DP += (float)e->t_dp; time_period += (float)e->t_rounds;
        query_counter += (int)e->candidates;
        e->t_di = e->t_dp = e->t_rounds = 0; e->candidates = 0;
    }
}

struct GlobalsInit {
    GlobalsInit() { orgInputString = (int *)calloc(kMaxLen + 8, sizeof(int)); }
} g_globals_init;

}  // namespace

// ================================================================ the reference's entry points
extern "C" void mtr_flush(void)
{
    Runtime &rt = runtime();
    if (rt.pending.empty()) return;
    const std::string out = rt.engines[0]->process(rt.pending, rt.print_alignment);
    rt.account(rt.engines[0]->ps);
    fwrite(out.data(), 1, out.size(), stdout);
    fflush(stdout);
    rt.pending.clear();
    rt.pending_bases = 0;
    publish_timers(rt);
}

// handle_one_read.c:263: the read is orgInputString[0..inputLen)
extern "C" void handle_one_read(char *readID, int inputLen, int read_cnt, int print_alignment)
{
    (void)read_cnt;
    if (inputLen <= 0) return;
    Runtime &rt = runtime();
    if (!rt.pending.empty() && rt.print_alignment != print_alignment) mtr_flush();
    rt.print_alignment = print_alignment;
    ReadInput in;
    in.id = readID ? readID : "";
    in.len = inputLen;
    in.bases.resize(inputLen + 2);
    for (int i = 0; i < inputLen; i++) in.bases[i] = (uint8_t)(orgInputString[i] & 3);
    const int tail[2] = {orgInputString[inputLen], orgInputString[inputLen + 1]};
    rt.pending_bases += inputLen;
    rt.pending.push_back(std::move(in));
    rt.stale.visit_batch(rt.pending, rt.pending.size() - 1, rt.pending.size(), 1, tail);
    if ((int)rt.pending.size() >= rt.batch_reads || rt.pending_bases >= rt.batch_bases) mtr_flush();
}

// handle_one_file.c:271: batches go to the GPUs round-robin, output is printed in input order
extern "C" int handle_one_file(char *inputFile, int print_alignment)
{
    Runtime &rt = runtime();
    mtr_flush();
    FastaReader reader(inputFile);
    const int n_eng = (int)rt.engines.size();      // engines = batches in flight (per_gpu for every GPU)
    const int n_gpu = n_eng / rt.per_gpu;
    const double t_file0 = now_s();
    struct Slot { std::thread th; std::string out; std::vector<ReadInput> reads; Engine *eng = nullptr; };
    std::vector<Slot *> inflight;
    auto drain_front = [&]() {
        Slot *s = inflight.front();
        s->th.join();
        rt.account(s->eng->ps);
        if (getenv("MTR_PROFILE")) fprintf(stderr, "[mtr profile] batch of %zu reads done at %.3f s (di %.0f ms, rounds %.0f ms)\n", s->reads.size(), now_s() - t_file0, s->eng->ps.di_wall_ms, s->eng->ps.rounds_wall_ms);
        fwrite(s->out.data(), 1, s->out.size(), stdout);
        fflush(stdout);
        delete s;
        inflight.erase(inflight.begin());
    };
    int n_reads = 0;
    long long batch_index = 0;
    bool more = true;
    std::vector<Engine *> last_on_gpu(n_gpu, nullptr);
    while (more) {
        Slot *s = new Slot();
        long long bases = 0;
        while ((int)s->reads.size() < rt.batch_reads && bases < rt.batch_bases) {
            ReadInput in;
            if (!reader.next(in)) { more = false; break; }
            bases += in.len;
            s->reads.push_back(std::move(in));
            n_reads++;
        }
        if (s->reads.empty()) { delete s; break; }
        rt.stale.visit_batch(s->reads, 0, s->reads.size(), rt.prep_threads, nullptr);
        while ((int)inflight.size() >= n_eng) drain_front();
        // consecutive batches alternate between the GPUs first, then between the engines of one GPU
        const int gpu = (int)(batch_index % n_gpu);
        Engine *eng = rt.engines[gpu * rt.per_gpu + (batch_index / n_gpu) % rt.per_gpu];
        batch_index++;
        // stagger: the batch before this one on the same GPU must be down to its last reads
        if (Engine *prev = last_on_gpu[gpu])
            while (prev != eng && prev->unfinished.load() > 0 &&
                   (rt.stagger_by_bases ? prev->bases_left.load() > (long long)(rt.stagger_frac * prev->bases_total.load())
                                        : prev->unfinished.load() > (int)(rt.stagger_frac * prev->batch_total.load())))
                std::this_thread::sleep_for(std::chrono::microseconds(500));
        last_on_gpu[gpu] = eng;
        eng->batch_total.store((int)s->reads.size()); eng->unfinished.store((int)s->reads.size());
        { long long tb = 0; for (const ReadInput &r : s->reads) tb += r.len; eng->bases_total.store(tb); eng->bases_left.store(tb); }
        s->eng = eng;
        if (getenv("MTR_PROFILE")) fprintf(stderr, "[mtr profile] batch %lld (%zu reads) starts on engine %d at %.3f s\n", batch_index - 1, s->reads.size(), (int)(eng == rt.engines[gpu * rt.per_gpu] ? 0 : 1), now_s() - t_file0);
        s->th = std::thread([eng, s, print_alignment] { s->out = eng->process(s->reads, print_alignment); });
        inflight.push_back(s);
    }
    while (!inflight.empty()) drain_front();
    publish_timers(rt);
    return n_reads;
}

// Counters of everything handle_one_file / handle_one_read + mtr_flush processed since the last call (then reset).
extern "C" int mtr_file_stats(mtr_pipeline_stats *out)
{
    if (!out) return MTR_EINVAL;
    Runtime &rt = runtime();
    *out = rt.totals;
    memset(&rt.totals, 0, sizeof rt.totals);
    return MTR_OK;
}

// ================================================================ batch-level pipeline ABI (bench, tests, embedding)
struct mtr_pipeline {
    Engine *eng = nullptr;
    StaleTracker *stale = nullptr;
    std::vector<ReadInput> reads;
    std::string out;
};

extern "C" int mtr_pipeline_open(int device, int threads, mtr_pipeline **out)
{
    if (!out) return MTR_EINVAL;
    *out = nullptr;
    mtr_ctx *probe = nullptr;
    const int rc = mtr_cuda_init(device, &probe);       // fail with a code instead of exiting
    if (rc) return rc;
    mtr_cuda_shutdown(probe);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    mtr_pipeline *p = new mtr_pipeline();
    p->eng = new Engine(device, std::max(1, threads));
    p->stale = new StaleTracker();
    *out = p;
    return MTR_OK;
}

extern "C" void mtr_pipeline_close(mtr_pipeline *p)
{
    if (!p) return;
    delete p->eng; delete p->stale; delete p;
}

// Parses FASTA text held in host memory (same rules as handle_one_file), reproduces the cross-read stale state,
// packs the reads to 2 bit and uploads them.  Returns the number of reads now resident, or a negative code.
extern "C" int mtr_pipeline_load_fasta_shard(mtr_pipeline *p, const char *text, int64_t len, int first, int count);

extern "C" int mtr_pipeline_load_fasta(mtr_pipeline *p, const char *text, int64_t len)
{
    return mtr_pipeline_load_fasta_shard(p, text, len, 0, -1);
}

// Keeps reads [first, first + count) of the text (count < 0: to the end).  The reads before `first` are still
// visited by the stale-state tracker, so a shard behaves exactly as it would inside the whole file (H3/H4a).
extern "C" int mtr_pipeline_load_fasta_shard(mtr_pipeline *p, const char *text, int64_t len, int first, int count)
{
    if (!p || (!text && len > 0) || first < 0) return MTR_EINVAL;
    p->reads.clear();
    // record boundaries: '>' at the start of a line
    std::vector<int64_t> starts;
    for (const char *q = text; q && q < text + len;) {
        q = (const char *)memchr(q, '>', (size_t)(text + len - q));
        if (!q) break;
        if (q == text || q[-1] == '\n') starts.push_back(q - text);
        q++;
    }
    int nrec = (int)starts.size();
    if (count >= 0) nrec = std::min(nrec, first + count);
    starts.push_back(nrec < (int)starts.size() ? starts[nrec] : len);
    p->reads.resize(nrec);
    static const struct Lut { int8_t v[256]; Lut() { memset(v, -1, sizeof v); v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; v['\n'] = v['\r'] = -2; } } lut;
    std::atomic<int> next{0}, bad{0};
    auto parse = [&] {
        for (int i; (i = next.fetch_add(1)) < nrec;) {
            ReadInput &in = p->reads[i];
            const char *q = text + starts[i] + 1, *end = text + starts[i + 1];
            const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
            const char *hend = nl ? nl : end;
            const char *idend = hend;
            for (const char *c = q; c < hend; c++) if (*c == '\r') { idend = c; break; }
            in.id.assign(q, idend);
            in.bases.resize((size_t)(end - hend) + 2);
            uint8_t *dst = in.bases.data();
            for (const char *c = hend; c < end; c++) {
                const int8_t b = lut.v[(unsigned char)*c];
                if (b >= 0) *dst++ = (uint8_t)b;
                else if (b == -1) { bad.store(1); break; }
            }
            in.len = (int)(dst - in.bases.data());
            in.bases.resize((size_t)in.len + 2);
            if (in.len >= kMaxLen) bad.store(2);
        }
    };
    {
        const int T = std::max(1, std::min(p->eng->pool->size(), nrec / 16));
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back(parse);
        parse();
        for (auto &x : th) x.join();
    }
    if (bad.load() == 1) return MTR_EINVAL;                 // the reference aborts: "Invalid character"
    if (bad.load() == 2) return MTR_ERANGE;
    for (int i = 0; i < nrec; i++)
        if (p->reads[i].len == 0) { p->reads.resize(i); break; }   // a zero-length read ends the run (handle_one_file.c:283)
    // the reads before `first` only carry the stale state forward
    p->stale->visit_batch(p->reads, 0, p->reads.size(), p->eng->pool->size(), nullptr);
    if (first > 0) p->reads.erase(p->reads.begin(), p->reads.begin() + std::min<size_t>((size_t)first, p->reads.size()));
    p->eng->prepare(p->reads);
    return (int)p->reads.size();
}

// Runs the pipeline on the resident batch (may be called repeatedly).  *out_text stays valid until the next call.
extern "C" int mtr_pipeline_run(mtr_pipeline *p, int print_alignment, const char **out_text, int64_t *out_len)
{
    if (!p) return MTR_EINVAL;
    p->out = p->eng->run(p->reads, print_alignment);
    if (out_text) *out_text = p->out.data();
    if (out_len) *out_len = (int64_t)p->out.size();
    return MTR_OK;
}

extern "C" int mtr_pipeline_get_stats(const mtr_pipeline *p, mtr_pipeline_stats *out)
{
    if (!p || !out) return MTR_EINVAL;
    *out = p->eng->ps;
    return MTR_OK;
}

// Bench support: log the DP jobs of the next runs / hand the log out (valid until the next run).
extern "C" int mtr_pipeline_log_jobs(mtr_pipeline *p, int on)
{
    if (!p) return MTR_EINVAL;
    p->eng->log_jobs = on != 0;
    return MTR_OK;
}

extern "C" int mtr_pipeline_get_job_log(mtr_pipeline *p, const mtr_wdp_job **jobs, int64_t *n_jobs, const uint8_t **units, int64_t *units_len)
{
    if (!p || !jobs || !n_jobs || !units || !units_len) return MTR_EINVAL;
    *jobs = p->eng->job_log.data(); *n_jobs = (int64_t)p->eng->job_log.size();
    *units = p->eng->unit_log.data(); *units_len = (int64_t)p->eng->unit_log.size();
    return MTR_OK;
}

// The context that owns the resident reads of the pipeline's batch (for replaying logged jobs with mtr_wdp_*).
extern "C" mtr_ctx *mtr_pipeline_ctx(mtr_pipeline *p) { return p ? p->eng->ctx : nullptr; }
