// cluster.cpp -- optional post-pass over the emitted records: cross-read clustering of tandem repeats
// by (unit length, circular 2-mer vector of the unit), restated from the reference's k_means_clustering.c:136-355.
//
// That file is DEAD in the reference: it is not in the Makefile's OBJS and does not compile (it uses fields that
// repeat_in_read no longer has and the constants MIN_REP_LEN, MIN_NUM_repTR and MH_distance_threshold, defined nowhere), so
// there is no binary to compare with: PARITY UNPINNED.  What is restated is the algorithm as written:
//   k_means_clustering   :251-355  qualify (MIN_REP_LEN < period * units, MIN_MATCH_RATIO < matches / repeat_len, 1 < units),
//                                   sort by <period, 2-mer vector, units>, group, revise, sort by <-freq, id> of the
//                                   representatives, print every qualified record with its representative's id and unit
//   select_repTR_list    :136-167  runs of identical <period, 2-mer vector>; the LAST record of a run represents it
//   TRs_in_neighborhood  :169-180  sum |d 2-mer| <= MH_distance_threshold * period of the representative
//   revise_repTR_list    :182-233  a representative joins the most frequent representative within 10 % of its period
//                                   that is in its neighbourhood (nearest first in both directions, strictly more
//                                   frequent than the best so far); roots are followed, frequencies accumulate
//   cmp_TR               :62-101   the three orders
// Choices where the source leaves none: the constants are parameters (defaults: MIN_MATCH_RATIO 0.6 = mTR.h:32,
// MH_distance_threshold 0.3 = the value SURVEY.md App. A records, MIN_REP_LEN 0, MIN_NUM_repTR 1 -- with any larger value
// the source dereferences the NULL representative of a small run at :238); the source sorts with a random-pivot
// quicksort (rand(), :103-134), which leaves the order of equal keys open -- here equal keys keep their input order.
// The 2-mer vector is freq_2mer_array (handle_one_read.c:63-72) of the record's unit string.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mtr_b200.h"

namespace {

struct TR {
    int id = 0, rep = -1;              // ID (index of the record in the input), index of the representative in `reps`
    int freq = 1;
    int period = 0, units = 0;
    int f2[16];
    size_t line0 = 0, line1 = 0;       // the record's text
    size_t unit0 = 0, unit1 = 0;       // its unit string
    size_t head1 = 0, period0 = 0, period1 = 0;   // text before the period field / the period field
};

struct Rep { int tr; int id; int freq; int period; int f2[16]; int parent; };   // parent: index in reps (itself: a root)

int cmp_tr(const TR &a, const TR &b, int mode)         // k_means_clustering.c:62-101, modes 0 and 1
{
    int d = a.period - b.period;
    if (d) return d;
    for (int i = 0; i < 16; i++) { d = a.f2[i] - b.f2[i]; if (d) return d; }
    return mode == 1 ? a.units - b.units : 0;
}

bool in_neighborhood(const int *a, const Rep &rep, double thr)      // :169-180
{
    int diff = 0;
    for (int i = 0; i < 16; i++) diff += abs(a[i] - rep.f2[i]);
    return !(thr * rep.period < diff);
}

}   // namespace

extern "C" int mtr_cluster_records(const char *tsv, int64_t len, const mtr_cluster_params *prm, char **out_text, int64_t *out_len)
{
    if (!tsv || len < 0 || !out_text || !out_len) return MTR_EINVAL;
    mtr_cluster_params P;
    P.min_match_ratio = 0.6; P.mh_distance_threshold = 0.3; P.min_rep_len = 0; P.min_num_rep = 1;
    if (prm) P = *prm;
    if (P.min_num_rep != 1) return MTR_EINVAL;          // the source is only defined for 1 (see the header of this file)
    *out_text = nullptr; *out_len = 0;

    // ---- records: readID, inputLen, start, end, repeat_len, period, units, matches, ratio, mismatches, ins, del, unit
    std::vector<TR> list;
    int id = 0;
    for (size_t at = 0; at < (size_t)len;) {
        size_t e = at;
        while (e < (size_t)len && tsv[e] != '\n') e++;
        size_t tab[12];
        int nt = 0;
        for (size_t i = at; i < e && nt < 12; i++) if (tsv[i] == '\t') tab[nt++] = i;
        size_t end = e;
        if (end > at && tsv[end - 1] == '\r') end--;
        bool ok = nt == 12 && end > tab[11] + 1;
        for (size_t i = ok ? tab[11] + 1 : end; ok && i < end; i++) ok = tsv[i] == 'A' || tsv[i] == 'C' || tsv[i] == 'G' || tsv[i] == 'T';
        if (ok) {                                                  // (alignment lines of -a output and blank lines are skipped)
            TR t;
            t.id = id++;
            const long repeat_len = atol(tsv + tab[3] + 1), units = atol(tsv + tab[5] + 1), matches = atol(tsv + tab[6] + 1);
            t.period = (int)(end - tab[11] - 1);                   // the unit as printed (= rep_period)
            t.units = (int)units;
            t.line0 = at; t.line1 = end; t.unit0 = tab[11] + 1; t.unit1 = end; t.head1 = tab[4] + 1; t.period0 = tab[4] + 1; t.period1 = tab[5];
            for (int i = 0; i < 16; i++) t.f2[i] = 0;              // freq_2mer_array, handle_one_read.c:63-72
            auto code = [&](size_t i) { const char c = tsv[i]; return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; };
            for (size_t i = t.unit0 + 1; i < t.unit1; i++) t.f2[code(i - 1) * 4 + code(i)]++;
            t.f2[code(t.unit1 - 1) * 4 + code(t.unit0)]++;
            const float ratio = repeat_len > 0 ? (float)matches / (float)repeat_len : 0.0f;
            if ((long)P.min_rep_len < (long)t.period * units && P.min_match_ratio < ratio && 1 < units) list.push_back(t);   // :270-276
        }
        at = e + 1;
    }
    const int n = (int)list.size();

    // ---- sort by <period, 2-mer vector, units> (:303), runs of identical <period, 2-mer vector> (:136-167)
    std::stable_sort(list.begin(), list.end(), [](const TR &a, const TR &b) { return cmp_tr(a, b, 1) < 0; });
    std::vector<Rep> reps;
    for (int i = 0; i < n;) {
        int j = i;
        while (j + 1 < n && cmp_tr(list[j], list[j + 1], 0) == 0) j++;
        Rep r;
        r.tr = j; r.id = list[j].id; r.freq = j - i + 1; r.period = list[j].period; r.parent = (int)reps.size();
        memcpy(r.f2, list[j].f2, sizeof r.f2);
        for (int k = i; k <= j; k++) list[k].rep = (int)reps.size();
        reps.push_back(r);
        i = j + 1;
    }

    // ---- revise (:182-233): every representative looks for a more frequent one nearby
    const int nr = (int)reps.size();
    for (int i = 0; i < nr; i++) {
        const Rep &a = reps[i];
        const int lb = a.period - (int)(a.period * 0.1), ub = a.period + (int)(a.period * 0.1);
        int max_freq = a.freq, max_i = i;
        for (int j = i - 1; 0 <= j && lb <= reps[j].period; j--)
            if (in_neighborhood(a.f2, reps[j], P.mh_distance_threshold) && max_freq < reps[j].freq) { max_freq = reps[j].freq; max_i = j; }
        for (int j = i + 1; j < nr && reps[j].period <= ub; j++)
            if (in_neighborhood(a.f2, reps[j], P.mh_distance_threshold) && max_freq < reps[j].freq) { max_freq = reps[j].freq; max_i = j; }
        reps[i].parent = max_i;
    }
    // roots, in index order, the frequencies accumulating as the source's second loop does (:222-232): a representative
    // that has been re-pointed adds its CURRENT frequency (which may already hold what earlier ones added) to its root
    for (int i = 0; i < nr; i++) {
        if (reps[i].parent == i) continue;
        int r = reps[i].parent;
        while (reps[r].parent != r) r = reps[r].parent;            // (a strictly more frequent parent each step: no cycle)
        reps[i].parent = r;
        reps[r].freq += reps[i].freq;
    }
    // ---- every record takes its root (:236-249), then the order <-freq, id> of the roots (:337)
    std::vector<int> root((size_t)n);
    for (int i = 0; i < n; i++) {
        int r = list[i].rep;
        while (reps[r].parent != r) r = reps[r].parent;
        root[i] = r;
    }
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const Rep &a = reps[root[x]], &b = reps[root[y]];
        if (a.freq != b.freq) return a.freq > b.freq;
        return a.id < b.id;
    });

    // ---- text (:343-348, print_one_TR_with_read without pretty_print): the representative's id, then the record with the
    // representative's unit length and unit string in place of its own
    std::string out;
    char num[32];
    for (int k = 0; k < n; k++) {
        const TR &t = list[order[k]];
        const Rep &r = reps[root[order[k]]];
        const TR &rt = list[r.tr];
        snprintf(num, sizeof num, "%d\t", r.id);
        out += num;
        out.append(tsv + t.line0, t.period0 - t.line0);
        snprintf(num, sizeof num, "%d", rt.period);
        out += num;
        out.append(tsv + t.period1, t.unit0 - t.period1);
        out.append(tsv + rt.unit0, rt.unit1 - rt.unit0);
        out += '\n';
    }
    char *mem = (char *)malloc(out.size() + 1);
    if (!mem) return MTR_ENOMEM;
    memcpy(mem, out.data(), out.size());
    mem[out.size()] = 0;
    *out_text = mem; *out_len = (int64_t)out.size();
    return MTR_OK;
}

extern "C" void mtr_cluster_free(char *text) { free(text); }
