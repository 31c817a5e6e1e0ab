// pipeline.cpp -- host side of the drop-in: mTR's entry points (handle_one_file / handle_one_read) on top of the
// resident engine.  Replaces /root/reference/handle_one_file.c (FASTA reader, the per-read loop) and chaining.cpp
// (sweep-line chaining, the TSV record, the -a alignment text); everything between "a read is in memory" and "the list
// of repeats to chain" -- handle_one_TR, handle_one_read.c:190-261 -- runs on the GPU (mtr_engine_run, eng.cu).
//
//   reader   parses the FASTA in input order, reproduces the reference's cross-read stale state (StaleTracker) and cuts
//            the stream into GROUPS of reads (MTR_GROUP_READS / MTR_GROUP_MBASES)
//   workers  one host thread per engine context (MTR_GROUPS_PER_GPU contexts on each of MTR_GPUS devices) pulls the next
//            group: 2-bit pack, upload, mtr_engine_run (directional index + waves), then chaining and formatting on the
//            host; with -a the printed repeats go back to the GPU once more as PATH jobs of K3 (mtr_wdp_run)
//   writer   prints the groups' text in input order
// Reads are independent units, so the groups are pulled dynamically by whichever context is free, on whichever GPU:
// no collective, no exchange step.  There is no CPU implementation of any device stage in this file: without a usable
// GPU handle_one_file fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <time.h>
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

// host cores this process may count on: the machine's, divided by the ranks a launcher has put on it (torchrun exports
// LOCAL_WORLD_SIZE; one process per GPU share the host)
int host_cores()
{
    int n = (int)std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) n = std::max(1, n / std::max(1, atoi(e)));
    return n;
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
};

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


// ---------------------------------------------------------------- reads as parsed
struct ReadInput {
    std::string id;
    std::vector<uint8_t> bases;        // len + 2 (the two stale bases of H4a at the end)
    std::vector<uint16_t> stale;       // inputString_w_rand beyond len + 4r (H3)
    int len = 0;
    int8_t bad = 0;                    // parse_records: 1 = a character that is not a base (bad_char), 2 = MAX_INPUT_LENGTH reached
    char bad_char = 0;
};

[[noreturn]] void die(mtr_ctx *ctx, const char *what, int rc)
{
    if (rc == MTR_ERANGE) fprintf(stderr, "%s\n", mtr_last_error(ctx));        // the reference's own abort messages
    else fprintf(stderr, "mTR (B200): %s failed (%d): %s\n", what, rc, mtr_last_error(ctx));
    // other workers are still launching and the writer may be printing: leave without running exit handlers and static
    // destructors under their feet (the streams are flushed here)
    fflush(NULL);
    _exit(EXIT_FAILURE);
}

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
    void reset() { std::fill(org.begin(), org.end(), 0); std::fill(padded.begin(), padded.end(), 0); }
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
    std::string fatal;                 // the reference's abort message when the input ends the run (the caller prints it and exits
                                       // AFTER the reads in front of the bad one are done: the reference parses a read only when
                                       // the one before it has been printed, handle_one_file.c:277-289)
    std::vector<char> buf;
    // lines are read in pieces of BLK - 1 = 4095 characters exactly as the reference's fgets(s, BLK, fp) does (handle_one_file.c:208):
    // what follows the first piece of a longer header line is read as bases there, and here
    explicit FastaReader(const char *path) : buf(4096)
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
                if (b < 0) { char m[64]; snprintf(m, sizeof m, "Invalid character: %c \n", s[i]); fatal = m; eof = true; return false; }
                dst[i] = (uint8_t)b;
            }
            if (kMaxLen <= (int)out.bases.size()) {
                char m[2304];
                snprintf(m, sizeof m, "fatal error: The length %d is tentatively at most %i.\nread ID = %.2000s\nSet MAX_INPUT_LENGTH to a larger value"
                         "cannot allocate space for one of global variables in the heap.\n", (int)out.bases.size(), kMaxLen, out.id.c_str());
                fatal = m; eof = true;
                return false;
            }
        }
        eof = true;
        if (!any) return false;
        out.id = next_id;
        out.len = (int)out.bases.size();
        return out.len > 0;
    }
};

// The same parser for FASTA text that is in memory (a block of a file in handle_one_file, the whole text in
// mtr_pipeline_load_fasta*): records are cut at every '>' that starts a line and parsed in parallel, line by line with the
// reference's rules (handle_one_file.c:201-269): a line ends at its first '\r' or '\n' (what follows a '\r' in the same line is
// dropped), a header line is read in pieces of BLK - 1 characters (the rest of a longer one is bases), anything but
// A/C/G/T in either case is fatal, a read of MAX_INPUT_LENGTH bases or more is fatal.  Bases in front of the first header
// join the first read (only where `leading` is set: the start of the input).  A bad record is marked (ReadInput::bad) and
// parsing goes on: the caller decides what happens to the records in front of it.
void parse_records(const char *text, int64_t len, bool leading, int max_records, int threads, std::vector<ReadInput> &reads)
{
    reads.clear();
    std::vector<int64_t> starts;
    for (const char *q = text; q && q < text + len;) {
        q = (const char *)memchr(q, '>', (size_t)(text + len - q));
        if (!q) break;
        if (q == text || q[-1] == '\n') starts.push_back(q - text);
        q++;
    }
    const int64_t prefix_end = starts.empty() ? len : starts[0];      // headerless bases in front
    int nrec = (int)starts.size();
    const bool headerless_only = nrec == 0 && prefix_end > 0 && leading;
    if (headerless_only) { starts.push_back(0); nrec = 1; }
    if (max_records >= 0) nrec = std::min(nrec, max_records);
    starts.push_back(nrec < (int)starts.size() ? starts[nrec] : len);
    reads.resize((size_t)nrec);
    static const struct Lut { int8_t v[256]; Lut() { memset(v, -1, sizeof v); v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; } } lut;
    std::atomic<int> next{0};
    auto parse = [&] {
        for (int i; (i = next.fetch_add(1)) < nrec;) {
            ReadInput &in = reads[(size_t)i];
            const char *q = text + starts[(size_t)i], *end = text + starts[(size_t)i + 1];
            const char *hend = q;
            if (!(headerless_only && i == 0)) {
                q++;                                               // past '>'
                const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
                hend = nl ? nl : end;
                if (hend - (q - 1) > 4095) hend = q - 1 + 4095;     // (the reference reads a header line in pieces of BLK - 1 characters: the rest is bases)
                const char *idend = hend;
                for (const char *c = q; c < hend; c++) if (*c == '\r') { idend = c; break; }
                in.id.assign(q, idend);
            }
            const int64_t pre = (i == 0 && leading && !headerless_only) ? prefix_end : 0;
            in.bases.resize((size_t)std::min<int64_t>((end - hend) + pre, kMaxLen) + 2);
            uint8_t *dst = in.bases.data(), *cap = dst + kMaxLen;
            auto take = [&](const char *a, const char *b) {       // the lines of [a, b)
                while (a < b && !in.bad) {
                    const char *nl = (const char *)memchr(a, '\n', (size_t)(b - a));
                    const char *le = nl ? nl : b;
                    for (const char *c = a; c < le && *c != '\r'; c++) {
                        const int8_t v = lut.v[(unsigned char)*c];
                        if (v < 0) { in.bad = 1; in.bad_char = *c; break; }
                        if (dst < cap) *dst++ = (uint8_t)v;
                        if (dst == cap) { in.bad = 2; break; }     // MAX_INPUT_LENGTH reached (handle_one_file.c:190-197)
                    }
                    a = nl ? nl + 1 : b;
                }
            };
            if (pre > 0) take(text, text + pre);
            take(hend, end);
            in.len = (int)(dst - in.bases.data());
            in.bases.resize((size_t)in.len + 2);
        }
    };
    const int T = std::max(1, std::min(threads, nrec / 16));
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(parse);
    parse();
    for (auto &x : th) x.join();
}

// Keeps records [0, first + count) (count < 0: all); MTR_EINVAL / MTR_ERANGE where the reference aborts
int parse_fasta_text(const char *text, int64_t len, int first, int count, int threads, std::vector<ReadInput> &reads)
{
    parse_records(text, len, true, count >= 0 ? first + count : -1, threads, reads);
    for (size_t i = 0; i < reads.size(); i++) {
        if (reads[i].bad == 1) return MTR_EINVAL;           // the reference aborts: "Invalid character"
        if (reads[i].bad == 2) return MTR_ERANGE;
        if (reads[i].len == 0) { reads.resize(i); break; }   // a zero-length read ends the run (handle_one_file.c:283)
    }
    return MTR_OK;
}

// ---------------------------------------------------------------- block reader of handle_one_file
// The file is read in blocks (MTR_READ_BLOCK_MB, default 64), every block is cut at its last record boundary and its
// records are parsed in parallel by parse_records; what is left over joins the next block.  Same results as FastaReader
// (the reference's own line-by-line loop, kept for MTR_SERIAL_READER=1), and the same order of events: the reads in front
// of a bad record are handed out before `fatal` is reported, a zero-length read ends the run.
struct BlockReader {
    FILE *fp = nullptr;
    std::vector<char> buf;
    size_t have = 0, block = 64u << 20;
    bool eof = false, leading = true, done = false;
    int threads = 4;
    std::string fatal, last_id;
    explicit BlockReader(const char *path)
    {
        fp = fopen(path, "r");
        if (!fp) { fprintf(stderr, "fatal error: cannot open %s\n", path); fflush(stderr); exit(EXIT_FAILURE); }
        if (const char *e = getenv("MTR_READ_BLOCK_MB")) block = (size_t)std::max(1, atoi(e)) << 20;
        if (const char *e = getenv("MTR_READ_BLOCK_BYTES")) block = (size_t)std::max(16, atoi(e));      // (tests: boundaries everywhere)
        threads = std::min(8, std::max(1, host_cores() / 4));
        if (const char *e = getenv("MTR_PARSE_THREADS")) threads = std::max(1, atoi(e));
    }
    ~BlockReader() { if (fp) fclose(fp); }
    // the next reads of the file, in order; false: nothing more (see `fatal`)
    bool next(std::vector<ReadInput> &out)
    {
        out.clear();
        while (!done && out.empty()) {
            if (!eof) {
                if (buf.size() < have + block) buf.resize(have + block);
                const size_t got = fread(buf.data() + have, 1, block, fp);
                have += got;
                if (got < block) eof = true;
            }
            size_t cut = have;
            if (!eof) {
                // up to the last '>' that starts a line -- behind the first one, whose record must be whole (at the start of
                // the input the bases in front of the first header belong to it)
                size_t first = have;
                for (size_t p = 0; p < have; p++)
                    if (buf[p] == '>' && (p == 0 || buf[p - 1] == '\n')) { first = p; break; }
                    else if (!leading) break;                      // (later blocks begin with a header)
                cut = 0;
                for (size_t p = have; p-- > first + 1;)
                    if (buf[p] == '>' && buf[p - 1] == '\n') { cut = p; break; }
                if (cut == 0) continue;                            // no whole record yet: read on
            }
            if (cut == 0) { done = true; break; }
            parse_records(buf.data(), (int64_t)cut, leading, -1, threads, out);
            leading = false;
            memmove(buf.data(), buf.data() + cut, have - cut);
            have -= cut;
            if (eof && have == 0) done = true;
            for (size_t i = 0; i < out.size(); i++) {
                const ReadInput &r = out[i];
                if (r.bad == 1) {
                    char m[64];
                    snprintf(m, sizeof m, "Invalid character: %c \n", r.bad_char);
                    fatal = m;
                } else if (r.bad == 2) {
                    // (handle_one_file.c:244-248 prints currentRead->ID, which still holds the read before)
                    char m[2304];
                    snprintf(m, sizeof m, "fatal error: The length %d is tentatively at most %i.\nread ID = %.2000s\nSet MAX_INPUT_LENGTH to a larger value"
                             "cannot allocate space for one of global variables in the heap.\n", kMaxLen, kMaxLen, i ? out[i - 1].id.c_str() : last_id.c_str());
                    fatal = m;
                }
                if (r.bad || r.len == 0) { out.resize(i); done = true; break; }   // a zero-length read ends the run (handle_one_file.c:283)
            }
            if (!out.empty()) last_id = out.back().id;
        }
        return !out.empty();
    }
};

// ---------------------------------------------------------------- groups of reads on one engine context
// A group is one engine run: its reads pass through the engine's read slots (MTR_ENGINE_SLOTS at a time, a slot takes the
// next read as soon as its read has finished), so a group should be several times the slot count; two contexts per GPU
// let the tail of one group (a few reads with long dependent chains) overlap with the bulk of the next.
constexpr int kDefaultContexts = 2;
constexpr int kDefaultGroupReads = 24576;
constexpr long long kDefaultGroupBases = 320LL << 20;
// bases per group when the amount of input is known (a file's size, a parsed text): the smallest multiple of the number
// of contexts that keeps every group under the cap, so that the contexts finish together
long long balanced_group_bases(long long total_bases, int contexts, long long cap)
{
    if (total_bases <= 0 || contexts <= 0) return cap;
    const long long per_round = (long long)contexts * cap;
    const long long rounds = (total_bases + per_round - 1) / per_round;
    const long long groups = rounds * contexts;
    return std::min(cap, (total_bases + groups - 1) / groups + (1LL << 16));
}

struct Resident {                      // what stays on the host for the reads resident in one context
    std::vector<int64_t> stale_off;    // indexed like the resident batch
    std::vector<uint16_t> stale;
};

struct Group {
    std::vector<ReadInput> reads;      // in input order
    long long index = 0;               // position in the output order
    std::string out;
    mtr_pipeline_stats ps = {};
    // resident form (after upload): reads [first, first + reads.size()) of the context's batch
    std::shared_ptr<Resident> res;
    int first = 0;
};

// 2-bit packs the reads of the groups (in this order) and uploads them as ONE batch: afterwards they are resident in the
// context's HBM and every group is a range of that batch
void upload_groups(mtr_ctx *ctx, const std::vector<Group *> &groups)
{
    const double t0 = now_s();
    int n = 0;
    for (Group *g : groups) n += (int)g->reads.size();
    std::vector<int64_t> word_off((size_t)n + 1, 0);
    std::vector<int32_t> lens((size_t)n);
    std::shared_ptr<Resident> res(new Resident());
    res->stale_off.assign((size_t)n + 1, 0);
    int r = 0;
    for (Group *g : groups) {
        g->first = r;
        g->res = res;
        long long bases = 0;
        for (const ReadInput &in : g->reads) {
            lens[r] = in.len;
            const int64_t words = (in.len + 2 + 15) / 16;
            word_off[r + 1] = word_off[r] + ((words + 3) / 4) * 4;
            res->stale_off[r + 1] = res->stale_off[r] + (int64_t)in.stale.size();
            bases += in.len;
            r++;
        }
        memset(&g->ps, 0, sizeof g->ps);
        g->ps.reads = (int64_t)g->reads.size(); g->ps.bases = bases; g->ps.groups = 1;
    }
    std::vector<uint32_t> packed((size_t)word_off[n], 0u);
    res->stale.assign((size_t)res->stale_off[n], 0);
    // 2-bit packing, a contiguous share of the reads per thread (every read owns whole words)
    std::vector<const ReadInput *> flat((size_t)n);
    r = 0;
    for (Group *g : groups)
        for (const ReadInput &in : g->reads) flat[r++] = &in;
    auto pack_range = [&](int a, int b) {
        for (int q = a; q < b; q++) {
            const ReadInput &in = *flat[q];
            uint32_t *dst = packed.data() + word_off[q];
            const uint8_t *bs = in.bases.data();
            const int nb = in.len + 2;
            int i = 0;
            for (; i + 16 <= nb; i += 16) {
                uint32_t w = 0;
                for (int t = 0; t < 16; t++) w |= (uint32_t)bs[i + t] << (2 * t);
                dst[i >> 4] = w;
            }
            for (; i < nb; i++) dst[i >> 4] |= (uint32_t)bs[i] << ((i & 15) * 2);
            if (!in.stale.empty()) memcpy(res->stale.data() + res->stale_off[q], in.stale.data(), in.stale.size() * 2);
        }
    };
    {
        int nt = std::min(8, std::max(1, host_cores() / 2));
        if (const char *e = getenv("MTR_PACK_THREADS")) nt = std::max(1, atoi(e));
        if (n < 64) nt = 1;
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(pack_range, (int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt));
        pack_range(0, (int)((long long)n / nt));
        for (std::thread &x : th) x.join();
    }
    const int rc = mtr_reads_upload(ctx, packed.data(), word_off.data(), lens.data(), n);
    if (rc) die(ctx, "mtr_reads_upload", rc);
    const double ms = (now_s() - t0) * 1e3;
    for (Group *g : groups) {
        g->ps.h2d_bytes = (word_off[g->first + (int)g->reads.size()] - word_off[g->first]) * 4 + (int64_t)(g->reads.size() + 1) * 12;
        g->ps.pack_ms = ms * (double)g->reads.size() / std::max(n, 1);
    }
}

void upload_group(mtr_ctx *ctx, Group &g)
{
    std::vector<Group *> one(1, &g);
    upload_groups(ctx, one);
}

// handle_one_TR for every read of the resident group on the GPU, then chaining + printing on the host
void run_group(mtr_ctx *ctx, Group &g, int print_alignment, int manhattan, float ratio)
{
    const int n = (int)g.reads.size();
    g.out.clear();
    if (n == 0) return;
    const mtr_repeat *reps = nullptr;
    int64_t n_reps = 0;
    const uint8_t *units = nullptr;
    mtr_engine_stats es;
    const int rc = mtr_engine_run_range(ctx, g.first, n, manhattan, ratio, g.res->stale.data(), g.res->stale_off.data(), &reps, &n_reps, &units, &es);
    if (rc) { fflush(stdout); die(ctx, "mtr_engine_run", rc); }
    for (int64_t i = 0; i < es.wrapdp_messages; i++) fprintf(stderr, "You need to increse the value of WrapDPsize.\n");
    const double tc0 = now_s();
    // chaining per read (chaining.cpp:243-363) over the repeats in insertion order
    std::vector<std::vector<Rec>> printed((size_t)n);
    int64_t at = 0;
    std::vector<ChainItem> items;
    long long n_printed = 0;
    for (int r = 0; r < n; r++) {
        items.clear();
        for (; at < n_reps && reps[at].read == g.first + r; at++) {
            const mtr_repeat &q = reps[at];
            ChainItem it;
            it.rec.inputLen = g.reads[r].len; it.rec.rep_start = q.rep_start; it.rec.rep_end = q.rep_end; it.rec.repeat_len = q.repeat_len;
            it.rec.period = q.rep_period; it.rec.units = q.num_freq_unit; it.rec.nm = q.num_matches; it.rec.nx = q.num_mismatches;
            it.rec.ni = q.num_insertions; it.rec.nd = q.num_deletions; it.rec.kmer = q.kmer; it.rec.gain = q.match_gain;
            it.rec.mis = q.mismatch_penalty; it.rec.indel = q.indel_penalty;
            it.rec.unit.assign(units + q.unit_off, units + q.unit_off + q.rep_period);
            it.start = q.rep_start; it.end = q.rep_end; it.score = q.num_matches; it.pred = nullptr;
            items.push_back(std::move(it));
        }
        for (const Rec *p : best_chain(items)) printed[r].push_back(*p);
        n_printed += (long long)printed[r].size();
    }
    g.ps.d2h_bytes += es.d2h_bytes; g.ps.h2d_bytes += es.h2d_bytes;
    if (!print_alignment) {
        for (int r = 0; r < n; r++)
            for (const Rec &q : printed[r]) append_record(g.out, g.reads[r].id, q);
    } else if (n_printed > 0) {
        // pretty_print_alignment (wrap_around_DP.c:57-213) for the printed repeats: one batch of PATH jobs on K3
        std::vector<mtr_wdp_job> jobs;
        std::vector<uint8_t> junits;
        int64_t aux_bytes = 0;
        for (int r = 0; r < n; r++)
            for (const Rec &q : printed[r]) {
                mtr_wdp_job j;
                memset(&j, 0, sizeof j);
                j.read = g.first + r; j.first = q.rep_start - 1; j.rows = q.rep_end - q.rep_start + 1; j.ulen = q.period;
                j.unit_off = (int32_t)junits.size();
                j.gain[0] = (int8_t)q.gain; j.mis[0] = (int8_t)q.mis; j.indel[0] = (int8_t)q.indel;
                j.n_param = 1; j.mode = MTR_TB_PATH;
                aux_bytes = (aux_bytes + 15) & ~15LL;
                j.aux_off = aux_bytes; j.aux_cap = (int64_t)6 * j.rows + 64;
                aux_bytes += j.aux_cap;
                junits.insert(junits.end(), q.unit.begin(), q.unit.end());
                jobs.push_back(j);
            }
        std::vector<mtr_wdp_result> res(jobs.size() * 2);
        std::vector<uint8_t> aux((size_t)aux_bytes);
        const int rc2 = mtr_wdp_run(ctx, jobs.data(), (int)jobs.size(), junits.data(), (int64_t)junits.size(), res.data(), aux.data(), aux_bytes);
        if (rc2) { fflush(stdout); die(ctx, "mtr_wdp_run", rc2); }
        mtr_stats st;
        mtr_get_stats(ctx, &st);
        g.ps.wdp_cells += st.wdp_cells; g.ps.launches += st.launches; g.ps.dp_ms += st.wdp_fill_ms + st.wdp_tb_ms; g.ps.jobs += (int64_t)jobs.size();
        g.ps.h2d_bytes += (int64_t)jobs.size() * sizeof(mtr_wdp_job) + (int64_t)junits.size();
        g.ps.d2h_bytes += (int64_t)res.size() * sizeof(mtr_wdp_result) + aux_bytes;
        size_t k = 0;
        for (int r = 0; r < n; r++)
            for (const Rec &q : printed[r]) {
                append_record(g.out, g.reads[r].id, q);
                append_alignment(g.out, q, g.reads[r].bases.data(), res[2 * k], aux.data() + jobs[k].aux_off);
                k++;
            }
    }
    g.ps.chain_ms = (now_s() - tc0) * 1e3;
    g.ps.candidates += es.candidates; g.ps.waves += es.waves; g.ps.jobs += es.dp_jobs; g.ps.dp_tasks += es.dp_tasks;
    g.ps.wdp_cells += es.dp_cells; g.ps.wdp_slot_cells += es.dp_slot_cells; g.ps.wdp_dir_bytes += es.dp_dir_bytes;
    g.ps.spec_cells += es.spec_cells; g.ps.shared_cells += es.shared_cells; g.ps.wdp_cells_p16 += es.dp_cells_p16; g.ps.tables += es.tables; g.ps.table_positions += es.table_positions; g.ps.walks += es.walks;
    g.ps.repeats += es.repeats; g.ps.launches += es.launches;
    g.ps.dp_ms += es.dp_ms; g.ps.di_kernel_ms += es.di_ms; g.ps.uf_kernel_ms += es.uf_ms; g.ps.engine_wall_ms += es.wall_ms;
}

void add_stats(mtr_pipeline_stats &t, const mtr_pipeline_stats &p)
{
    t.reads += p.reads; t.bases += p.bases; t.candidates += p.candidates; t.waves += p.waves; t.groups += p.groups; t.jobs += p.jobs;
    t.dp_tasks += p.dp_tasks; t.wdp_cells += p.wdp_cells; t.wdp_slot_cells += p.wdp_slot_cells; t.wdp_dir_bytes += p.wdp_dir_bytes;
    t.spec_cells += p.spec_cells; t.shared_cells += p.shared_cells; t.wdp_cells_p16 += p.wdp_cells_p16; t.tables += p.tables; t.table_positions += p.table_positions; t.walks += p.walks; t.repeats += p.repeats;
    t.h2d_bytes += p.h2d_bytes; t.d2h_bytes += p.d2h_bytes; t.launches += p.launches;
    t.dp_ms += p.dp_ms; t.di_kernel_ms += p.di_kernel_ms; t.uf_kernel_ms += p.uf_kernel_ms; t.engine_wall_ms += p.engine_wall_ms;
    t.pack_ms += p.pack_ms; t.chain_ms += p.chain_ms;
}

// ---------------------------------------------------------------- process-wide state behind the C entry points
struct Runtime {
    std::vector<mtr_ctx *> ctxs;       // groups_per_gpu engine contexts on each GPU, GPU-major
    int n_gpu = 1, groups_per_gpu = kDefaultContexts;
    int group_reads = kDefaultGroupReads;            // reads per group (MTR_GROUP_READS) ...
    long long group_bases = kDefaultGroupBases;      // ... or bases per group (MTR_GROUP_MBASES), whichever fills first
    StaleTracker stale;
    std::vector<ReadInput> pending;    // handle_one_read: reads enqueued since the last flush
    int pending_print = 0, pending_manhattan = 1;
    float pending_ratio = 0.6f;
    bool atexit_set = false;
    mtr_pipeline_stats totals = {};    // summed over everything since the last mtr_file_stats

    Runtime()
    {
        if (const char *e = getenv("MTR_GPUS")) n_gpu = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_GROUPS_PER_GPU")) groups_per_gpu = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_GROUP_READS")) group_reads = std::max(1, atoi(e));
        if (const char *e = getenv("MTR_GROUP_MBASES")) group_bases = std::max(1LL, atoll(e)) << 20;
        int base = 0;
        if (const char *e = getenv("MTR_DEVICE")) base = atoi(e);
        for (int g = 0; g < n_gpu; g++)
            for (int i = 0; i < groups_per_gpu; i++) {
                mtr_ctx *c = nullptr;
                const int rc = mtr_cuda_init(base + g, &c);
                if (rc) die(nullptr, "mtr_cuda_init", rc);
                mtr_set_blocking_sync(c, 1);                   // the worker threads sleep while their waves run
                ctxs.push_back(c);
            }
    }
    ~Runtime() { for (mtr_ctx *c : ctxs) mtr_cuda_shutdown(c); }

    void publish(const mtr_pipeline_stats &p)                  // the -c report (main.c:108-121)
    {
        add_stats(totals, p);
        time_range += (float)(p.di_kernel_ms * 1e-3);
        time_period += (float)(p.engine_wall_ms * 1e-3);
        time_initialize_input_string += (float)(p.pack_ms * 1e-3);
        time_count_table += (float)(p.uf_kernel_ms * 1e-3);
        time_wrap_around_DP += (float)(p.dp_ms * 1e-3);
        time_chaining += (float)(p.chain_ms * 1e-3);
        query_counter += (int)p.candidates;
    }
};

Runtime *g_rt = nullptr;
Runtime &runtime()
{
    if (!g_rt) g_rt = new Runtime();
    return *g_rt;
}

// Groups go in in input order (submit), are pulled by whichever context is free and come out of `write` in input order.
class Dispatcher {
public:
    Dispatcher(Runtime &rt, int print_alignment, int manhattan, float ratio) : rt_(rt), print_(print_alignment), manhattan_(manhattan), ratio_(ratio)
    {
        // consecutive groups alternate between the GPUs: worker w serves context w, contexts are GPU-major
        for (size_t w = 0; w < rt.ctxs.size(); w++) workers_.emplace_back([this, w] { work(rt_.ctxs[order(w)]); });
    }
    ~Dispatcher() { finish(); }
    void submit(std::unique_ptr<Group> g)
    {
        std::unique_lock<std::mutex> lk(mu_);
        g->index = submitted_++;
        queue_.push_back(std::move(g));
        ready_.notify_one();
        // the caller is also the writer; it does not run more than two groups per context ahead of what it has written
        for (;;) {
            drain(lk, false);
            if (submitted_ - written_ <= 2 * (long long)rt_.ctxs.size()) break;
            done_cv_.wait(lk, [&] { return done_.count(written_) != 0; });
        }
    }
    void finish()
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (closed_) return;
        drain(lk, true);
        closed_ = true;
        ready_.notify_all();
        lk.unlock();
        for (std::thread &t : workers_) t.join();
    }
private:
    size_t order(size_t w) const
    {
        const size_t per = (size_t)rt_.groups_per_gpu, ng = (size_t)rt_.n_gpu;
        return (w % ng) * per + (w / ng);                      // worker 0 -> GPU 0, worker 1 -> GPU 1, ...
    }
    void work(mtr_ctx *ctx)
    {
        cudaSetDevice(ctx->device);
        for (;;) {
            std::unique_ptr<Group> g;
            {
                std::unique_lock<std::mutex> lk(mu_);
                ready_.wait(lk, [&] { return !queue_.empty() || closed_; });
                if (queue_.empty()) return;
                g = std::move(queue_.front());
                queue_.pop_front();
            }
            upload_group(ctx, *g);
            run_group(ctx, *g, print_, manhattan_, ratio_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                const long long idx = g->index;
                done_[idx] = std::move(g);
            }
            done_cv_.notify_all();
        }
    }
    // writes every finished group that is next in line; all == true: waits until everything submitted is written
    void drain(std::unique_lock<std::mutex> &lk, bool all)
    {
        for (;;) {
            auto it = done_.find(written_);
            if (it == done_.end()) {
                if (!all || written_ == submitted_) return;
                done_cv_.wait(lk, [&] { return done_.count(written_) != 0; });
                continue;
            }
            std::unique_ptr<Group> g = std::move(it->second);
            done_.erase(it);
            written_++;
            fwrite(g->out.data(), 1, g->out.size(), stdout);
            fflush(stdout);
            rt_.publish(g->ps);
        }
    }
    Runtime &rt_;
    int print_, manhattan_;
    float ratio_;
    std::mutex mu_;
    std::condition_variable ready_, done_cv_;
    std::deque<std::unique_ptr<Group>> queue_;
    std::map<long long, std::unique_ptr<Group>> done_;
    long long submitted_ = 0, written_ = 0;
    bool closed_ = false;
    std::vector<std::thread> workers_;
};

struct GlobalsInit {
    GlobalsInit() { orgInputString = (int *)calloc(kMaxLen + 8, sizeof(int)); }
} g_globals_init;

// cuts reads (already visited by the stale tracker) into groups and runs them
void run_reads(Runtime &rt, std::vector<ReadInput> &reads, int print_alignment, int manhattan, float ratio)
{
    Dispatcher disp(rt, print_alignment, manhattan, ratio);
    size_t a = 0;
    while (a < reads.size()) {
        std::unique_ptr<Group> g(new Group());
        long long bases = 0;
        while (a < reads.size() && (int)g->reads.size() < rt.group_reads && bases < rt.group_bases) {
            bases += reads[a].len;
            g->reads.push_back(std::move(reads[a++]));
        }
        disp.submit(std::move(g));
    }
    disp.finish();
}

}  // namespace

// ================================================================ the reference's entry points
extern "C" void mtr_flush(void)
{
    if (!g_rt) return;
    Runtime &rt = *g_rt;
    if (rt.pending.empty()) return;
    run_reads(rt, rt.pending, rt.pending_print, rt.pending_manhattan, rt.pending_ratio);
    rt.pending.clear();
}

// handle_one_read.c:263: the read is orgInputString[0..inputLen).  The reference finishes the read before it returns;
// here the read joins the pending group, which runs when it is full, when a parameter the reference would have applied
// to the earlier reads changes (-a, -p, -m), at mtr_flush(), and at process exit.
extern "C" void handle_one_read(char *readID, int inputLen, int read_cnt, int print_alignment)
{
    (void)read_cnt;
    if (inputLen <= 0) return;
    Runtime &rt = runtime();
    if (!rt.atexit_set) { atexit(mtr_flush); rt.atexit_set = true; }
    if (!rt.pending.empty() && (rt.pending_print != print_alignment || rt.pending_manhattan != Manhattan_Distance || rt.pending_ratio != min_match_ratio)) mtr_flush();
    rt.pending_print = print_alignment; rt.pending_manhattan = Manhattan_Distance; rt.pending_ratio = min_match_ratio;
    ReadInput in;
    in.id = readID ? readID : "";
    in.len = inputLen;
    in.bases.resize(inputLen + 2);
    for (int i = 0; i < inputLen; i++) in.bases[i] = (uint8_t)(orgInputString[i] & 3);
    const int tail[2] = {orgInputString[inputLen], orgInputString[inputLen + 1]};
    rt.pending.push_back(std::move(in));
    rt.stale.visit_batch(rt.pending, rt.pending.size() - 1, rt.pending.size(), 1, tail);
    long long bases = 0;
    for (const ReadInput &r : rt.pending) bases += r.len;
    if ((int)rt.pending.size() >= rt.group_reads * (int)rt.ctxs.size() || bases >= rt.group_bases * (long long)rt.ctxs.size()) mtr_flush();
}

// handle_one_file.c:271: groups of reads go to the engine contexts of all GPUs, output is printed in input order
extern "C" int handle_one_file(char *inputFile, int print_alignment)
{
    Runtime &rt = runtime();
    mtr_flush();
    rt.stale.reset();                                   // the reference allocates its buffers anew for every file (handle_one_file.c:71-136)
    const bool serial = getenv("MTR_SERIAL_READER") != nullptr && atoi(getenv("MTR_SERIAL_READER")) != 0;
    std::unique_ptr<FastaReader> reader(serial ? new FastaReader(inputFile) : nullptr);
    std::unique_ptr<BlockReader> blocks(serial ? nullptr : new BlockReader(inputFile));
    Dispatcher disp(rt, print_alignment, Manhattan_Distance, min_match_ratio);
    int n_reads = 0;
    long long group_bases = rt.group_bases;
    {
        struct stat st;
        if (stat(inputFile, &st) == 0 && st.st_size > 0) group_bases = balanced_group_bases((long long)st.st_size, (int)rt.ctxs.size(), rt.group_bases);
    }
    std::unique_ptr<Group> g(new Group());
    long long bases = 0;
    auto add = [&](ReadInput &&in) {                    // reads join the current group; a full group is sent off
        bases += in.len;
        g->reads.push_back(std::move(in));
        n_reads++;
        if ((int)g->reads.size() < rt.group_reads && bases < group_bases) return;
        rt.stale.visit_batch(g->reads, 0, g->reads.size(), 4, nullptr);
        disp.submit(std::move(g));
        g.reset(new Group());
        bases = 0;
    };
    if (serial) {
        ReadInput in;
        while (reader->next(in)) { add(std::move(in)); in = ReadInput(); }
    } else {
        std::vector<ReadInput> batch;
        while (blocks->next(batch))
            for (ReadInput &in : batch) add(std::move(in));
    }
    if (!g->reads.empty()) {
        rt.stale.visit_batch(g->reads, 0, g->reads.size(), 4, nullptr);
        disp.submit(std::move(g));
    }
    disp.finish();
    const std::string &fatal = serial ? reader->fatal : blocks->fatal;
    if (!fatal.empty()) {                               // every read in front of the bad one has been printed, as in the reference
        fflush(stdout);
        fputs(fatal.c_str(), stderr);
        fflush(stderr);
        exit(EXIT_FAILURE);
    }
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

// ================================================================ the reference's C <-> C++ bridge (mTR.h:146-175, chaining.h:30-56)
// insert_an_alignment_into_set / chaining are chaining.cpp's interface towards handle_one_read.c; pretty_print_alignment
// and print_freq are the two C functions chaining.cpp calls back.  The product's own per-read loop lives on the device and
// never goes through these symbols (run_group chains and prints a whole group); they are exported so that a caller
// that keeps the reference's handle_one_read.c -- or its own per-read code -- above this library finds every symbol of
// the reference's interface, with the reference's semantics: the set is taken in insertion order, the read
// is orgInputString (mTR.h:65), output goes to stdout.
namespace {
std::vector<ChainItem> g_set;              // set_of_alignments, chaining.cpp:201
std::vector<std::string> g_set_ids;
mtr_ctx *g_bridge_ctx = nullptr;           // pretty_print_alignment runs its DP on the GPU (one PATH job of K3)

int base_of_char(char c, const char *who)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: fprintf(stderr, "%s: fatal input char %c\n", who, c); exit(EXIT_FAILURE);
    }
}
}   // namespace

extern "C" void insert_an_alignment_into_set(char *readID, int inputLen, int rep_start, int rep_end, int repeat_len, int rep_period,
                                             int Num_freq_unit, int Num_matches, int Num_mismatches, int Num_insertions,
                                             int Num_deletions, int Kmer, int match_gain, int mismatch_penalty, int indel_penalty,
                                             char *string, int *string_score)
{
    (void)string_score;                    // kept by the reference's Alignment, read by nothing that prints
    ChainItem it;
    it.rec.inputLen = inputLen; it.rec.rep_start = rep_start; it.rec.rep_end = rep_end; it.rec.repeat_len = repeat_len;
    it.rec.period = rep_period; it.rec.units = Num_freq_unit; it.rec.nm = Num_matches; it.rec.nx = Num_mismatches;
    it.rec.ni = Num_insertions; it.rec.nd = Num_deletions; it.rec.kmer = Kmer; it.rec.gain = match_gain;
    it.rec.mis = mismatch_penalty; it.rec.indel = indel_penalty;
    // print_one_TR prints the string with %s (chaining.cpp:127): what follows its first NUL is not part of the unit
    for (int i = 0; string && string[i] && i < rep_period; i++) it.rec.unit.push_back((uint8_t)base_of_char(string[i], "insert_an_alignment_into_set"));
    it.start = rep_start; it.end = rep_end; it.score = Num_matches; it.pred = nullptr;
    g_set.push_back(std::move(it));
    g_set_ids.push_back(readID ? readID : "");
}

extern "C" void pretty_print_alignment(char *unit_string, int unit_len, int rep_start, int rep_end, int match_gain,
                                       int mismatch_penalty, int indel_penalty)
{
    const int rows = rep_end - rep_start + 1;
    if (!orgInputString || !unit_string || unit_len <= 0 || rows <= 0) return;
    if (!g_bridge_ctx) {
        int dev = 0;
        if (const char *e = getenv("MTR_DEVICE")) dev = atoi(e);
        const int rc = mtr_cuda_init(dev, &g_bridge_ctx);
        if (rc) die(nullptr, "mtr_cuda_init", rc);
    }
    mtr_ctx *ctx = g_bridge_ctx;
    // the rows of the DP are orgInputString[rep_start .. rep_end] (wrap_around_DP.c:86-88): a one-read batch of that window
    // plus the two bases behind it, job rows 1..rows = positions 0..rows-1 of the batch (first = -1)
    std::vector<uint8_t> win((size_t)rows + 3, 0);             // win[i] = row i, as append_alignment indexes it
    for (int i = 1; i <= rows + 2; i++) win[i] = (uint8_t)(orgInputString[rep_start + i - 1] & 3);
    const int64_t words = (rows + 2 + 15) / 16;
    const int64_t word_off[2] = {0, ((words + 3) / 4) * 4};
    std::vector<uint32_t> packed((size_t)word_off[1], 0u);
    for (int i = 0; i < rows + 2; i++) packed[i >> 4] |= (uint32_t)win[i + 1] << ((i & 15) * 2);
    const int32_t len = rows;
    int rc = mtr_reads_upload(ctx, packed.data(), word_off, &len, 1);
    if (rc) die(ctx, "mtr_reads_upload", rc);
    Rec r;
    r.rep_start = 1; r.rep_end = rows; r.period = unit_len; r.gain = match_gain; r.mis = mismatch_penalty; r.indel = indel_penalty;
    for (int i = 0; i < unit_len; i++) r.unit.push_back((uint8_t)base_of_char(unit_string[i], "pretty_print_alignment"));
    mtr_wdp_job j;
    memset(&j, 0, sizeof j);
    j.read = 0; j.first = -1; j.rows = rows; j.ulen = unit_len; j.unit_off = 0;
    j.gain[0] = (int8_t)match_gain; j.mis[0] = (int8_t)mismatch_penalty; j.indel[0] = (int8_t)indel_penalty;
    j.n_param = 1; j.mode = MTR_TB_PATH;
    j.aux_off = 0; j.aux_cap = (int64_t)6 * rows + 64;
    std::vector<uint8_t> aux((size_t)j.aux_cap);
    mtr_wdp_result res[2];
    rc = mtr_wdp_run(ctx, &j, 1, r.unit.data(), (int64_t)r.unit.size(), res, aux.data(), j.aux_cap);
    if (rc) { fflush(stdout); die(ctx, "mtr_wdp_run", rc); }
    std::string out;
    append_alignment(out, r, win.data(), res[0], aux.data());
    fwrite(out.data() + 1, 1, out.size() - 1, stdout);         // (the blank line before "match gain" is print_one_TR's, chaining.cpp:165)
}

extern "C" void chaining(int print_alignment)
{
    if (g_set.empty()) return;                                  // chaining.cpp:244
    const std::vector<const Rec *> chain = best_chain(g_set);
    for (const Rec *p : chain) {
        size_t at = 0;
        while (at < g_set.size() && &g_set[at].rec != p) at++;
        std::string out;
        append_record(out, g_set_ids[at], *p);
        fwrite(out.data(), 1, out.size(), stdout);
        if (print_alignment == 1) {
            printf("\n");
            std::string unit;
            for (uint8_t b : p->unit) unit += "ACGT"[b];
            pretty_print_alignment(&unit[0], p->period, p->rep_start, p->rep_end, p->gain, p->mis, p->indel);
        }
        fflush(stdout);
    }
    g_set.clear();
    g_set_ids.clear();
}

// consensus.c:1089-1131 (debugging aid of chaining.cpp, DEBUG_unit_score): one digit per unit position, the number of
// times the k-mer that starts there occurs in orgInputString[rep_start .. rep_end] ('*' from 10 on).  The window is
// coded as init_inputString does (consensus.c:37-57): k-mer codes below min(rep_end, inputLen - k + 1), raw bases above.
extern "C" void print_freq(int rep_start, int rep_end, int rep_period, char *string, int inputLen, int k)
{
    if (!orgInputString || !string || k < 1 || k > 15) return;
    std::map<unsigned, int> count;
    const int coded_end = std::min(rep_end, inputLen - k + 1);
    for (int i = rep_start; i <= rep_end; i++) {
        unsigned code = 0;
        if (i < coded_end) for (int t = 0; t < k; t++) code = 4u * code + (unsigned)(orgInputString[i + t] & 3);
        else if (i < inputLen && (k > 1 || i < rep_end)) code = (unsigned)(orgInputString[i] & 3);   // (k == 1: init_inputString never writes position rep_end -- 0 in a fresh process)
        count[code]++;
    }
    std::vector<int> unit((size_t)std::max(rep_period, 0));
    for (int i = 0; i < rep_period; i++) unit[i] = base_of_char(string[i], "print_freq");
    const unsigned keep = 1u << (2 * (k - 1));                  // pow4[k-1]
    unsigned tmp = 0;
    for (int i = 0; i < k - 1; i++) tmp = 4u * tmp + (unsigned)unit[i % std::max(rep_period, 1)];   // (k - 1 > rep_period: the reference reads uninitialised stack here)
    for (int i = 0; i < rep_period; i++) {
        const unsigned node = 4u * tmp + (unsigned)unit[(i + k - 1) % rep_period];
        tmp = node % keep;
        const std::map<unsigned, int>::const_iterator f = count.find(node);
        const int freq = f == count.end() ? 0 : f->second;
        if (freq < 10) printf("%i", freq); else printf("*");
    }
    printf("\n");
}

// ================================================================ batch-level pipeline ABI (bench, tests, embedding)
struct mtr_pipeline {
    std::vector<mtr_ctx *> ctxs;
    std::vector<std::unique_ptr<Group>> groups;        // in input order; group i is resident in ctxs[i % ctxs.size()]
    StaleTracker *stale = nullptr;
    std::string out;
    mtr_pipeline_stats ps = {};
    int threads = 1;
    int group_reads = kDefaultGroupReads;
    long long group_bases = kDefaultGroupBases;
};

extern "C" int mtr_pipeline_open(int device, int threads, mtr_pipeline **out)
{
    if (!out) return MTR_EINVAL;
    *out = nullptr;
    int k = kDefaultContexts;
    if (const char *e = getenv("MTR_GROUPS_PER_GPU")) k = std::max(1, atoi(e));
    mtr_pipeline *p = new mtr_pipeline();
    for (int i = 0; i < k; i++) {
        mtr_ctx *c = nullptr;
        const int rc = mtr_cuda_init(device, &c);            // fail with a code instead of exiting
        if (rc) { for (mtr_ctx *x : p->ctxs) mtr_cuda_shutdown(x); delete p; return rc; }
        mtr_set_blocking_sync(c, 1);
        p->ctxs.push_back(c);
    }
    p->threads = threads > 0 ? threads : host_cores();
    if (const char *e = getenv("MTR_GROUP_READS")) p->group_reads = std::max(1, atoi(e));
    if (const char *e = getenv("MTR_GROUP_MBASES")) p->group_bases = std::max(1LL, atoll(e)) << 20;
    p->stale = new StaleTracker();
    *out = p;
    return MTR_OK;
}

extern "C" void mtr_pipeline_close(mtr_pipeline *p)
{
    if (!p) return;
    for (mtr_ctx *c : p->ctxs) mtr_cuda_shutdown(c);
    delete p->stale;
    delete p;
}

extern "C" int mtr_pipeline_load_fasta_shard(mtr_pipeline *p, const char *text, int64_t len, int first, int count);

// Parses FASTA text held in host memory (same rules as handle_one_file), reproduces the cross-read stale state, packs the
// reads to 2 bit and uploads them, one group per engine context.  Returns the number of reads now resident.
extern "C" int mtr_pipeline_load_fasta(mtr_pipeline *p, const char *text, int64_t len)
{
    return mtr_pipeline_load_fasta_shard(p, text, len, 0, -1);
}

// Keeps reads [first, first + count) of the text (count < 0: to the end).  The reads before `first` are still visited by
// the stale-state tracker (which starts fresh with every call, like a new process), so a shard behaves exactly as it
// would inside the whole file (H3/H4a).
extern "C" int mtr_pipeline_load_fasta_shard(mtr_pipeline *p, const char *text, int64_t len, int first, int count)
{
    if (!p || (!text && len > 0) || first < 0) return MTR_EINVAL;
    p->groups.clear();
    p->stale->reset();
    std::vector<ReadInput> reads;
    const int rc = parse_fasta_text(text, len, first, count, p->threads, reads);
    if (rc) return rc;
    p->stale->visit_batch(reads, 0, reads.size(), p->threads, nullptr);
    if (first > 0) reads.erase(reads.begin(), reads.begin() + std::min<size_t>((size_t)first, reads.size()));
    const int n = (int)reads.size();
    // the same groups handle_one_file would cut (MTR_GROUP_READS / MTR_GROUP_MBASES), dealt round-robin to the contexts;
    // all groups of a context are uploaded as one resident batch
    long long total_bases = 0;
    for (const ReadInput &r : reads) total_bases += r.len;
    const long long group_bases = balanced_group_bases(total_bases, (int)p->ctxs.size(), p->group_bases);
    for (size_t a = 0; a < reads.size();) {
        std::unique_ptr<Group> g(new Group());
        long long bases = 0;
        while (a < reads.size() && (int)g->reads.size() < p->group_reads && bases < group_bases) {
            bases += reads[a].len;
            g->reads.push_back(std::move(reads[a++]));
        }
        g->index = (long long)p->groups.size();
        p->groups.push_back(std::move(g));
    }
    const size_t k = p->ctxs.size();
    std::vector<std::thread> th;
    for (size_t c = 0; c < k && c < p->groups.size(); c++)
        th.emplace_back([p, c, k] {
            cudaSetDevice(p->ctxs[c]->device);
            std::vector<Group *> mine;
            for (size_t gi = c; gi < p->groups.size(); gi += k) mine.push_back(p->groups[gi].get());
            upload_groups(p->ctxs[c], mine);
        });
    for (std::thread &t : th) t.join();
    return n;
}

// Runs the engine on the resident groups (may be called repeatedly).  *out_text stays valid until the next call.
extern "C" int mtr_pipeline_run(mtr_pipeline *p, int print_alignment, const char **out_text, int64_t *out_len)
{
    if (!p) return MTR_EINVAL;
    const int manhattan = Manhattan_Distance;
    const float ratio = min_match_ratio;
    const double t0 = now_s();
    const size_t k = p->ctxs.size();
    std::vector<std::thread> th;
    for (size_t c = 0; c < k && c < p->groups.size(); c++)
        th.emplace_back([p, c, k, print_alignment, manhattan, ratio] {
            cudaSetDevice(p->ctxs[c]->device);
            for (size_t gi = c; gi < p->groups.size(); gi += k) {
                Group &g = *p->groups[gi];
                const int64_t h2d = g.ps.h2d_bytes, reads = g.ps.reads, bases = g.ps.bases;
                const double pack = g.ps.pack_ms;
                memset(&g.ps, 0, sizeof g.ps);
                g.ps.h2d_bytes = h2d; g.ps.pack_ms = pack; g.ps.reads = reads; g.ps.bases = bases; g.ps.groups = 1;
                run_group(p->ctxs[c], g, print_alignment, manhattan, ratio);
            }
        });
    for (std::thread &t : th) t.join();
    p->out.clear();
    memset(&p->ps, 0, sizeof p->ps);
    for (auto &g : p->groups) { p->out += g->out; add_stats(p->ps, g->ps); }
    p->ps.wall_ms = (now_s() - t0) * 1e3;
    if (out_text) *out_text = p->out.data();
    if (out_len) *out_len = (int64_t)p->out.size();
    return MTR_OK;
}

extern "C" int mtr_pipeline_get_stats(const mtr_pipeline *p, mtr_pipeline_stats *out)
{
    if (!p || !out) return MTR_EINVAL;
    *out = p->ps;
    return MTR_OK;
}

// The first engine context of the pipeline (device queries, mtr_alu_probe).
extern "C" mtr_ctx *mtr_pipeline_ctx(mtr_pipeline *p) { return p && !p->ctxs.empty() ? p->ctxs[0] : nullptr; }
