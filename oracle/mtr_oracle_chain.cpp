/*
 * oracle/mtr_oracle_chain.cpp -- restatement of chaining.cpp:43-363.  TEST INFRASTRUCTURE ONLY.
 *
 * Canonical tie-break: the reference keeps its alignments in a std::set ordered by heap
 * address; here they are kept in insertion order, which equals the stock binary on every hazard-free input
 * and equals oracle/_ref/mTR_ref_det everywhere.
 * H2: the "erase, then ++" loop of chaining.cpp:316-328 skips an element and may step past end(); the same
 * std::multimap operations are issued in the same order so that libstdc++ behaves identically.
 */
#include "mtr_oracle.h"
#include <map>
#include <string>
#include <vector>

namespace {
struct Aln {
    mtro_rr rr;
    std::string id;
    int start, end;
    int score;
    Aln *pred;
};
}

struct mtro_chain { std::vector<Aln*> items; };

extern "C" mtro_chain *mtro_chain_new(void) { return new mtro_chain(); }

extern "C" void mtro_chain_free(mtro_chain *ch)
{
    if (!ch) return;
    for (Aln *a : ch->items) delete a;
    delete ch;
}

extern "C" void mtro_chain_insert(mtro_chain *ch, const char *read_id, const mtro_rr *rr)
{
    Aln *a = new Aln();
    a->rr = *rr; a->id = read_id; a->start = rr->rep_start; a->end = rr->rep_end;
    a->score = rr->n_match; a->pred = nullptr;
    ch->items.push_back(a);
}

static void print_record(FILE *out, const Aln *a)        /* print_one_TR, chaining.cpp:125-143 */
{
    const mtro_rr &r = a->rr;
    fprintf(out, "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%f\t%d\t%d\t%d\t%s\n", a->id.c_str(), r.inputLen,
            r.rep_start + 1, r.rep_end + 1, r.repeat_len, r.rep_period, r.n_units, r.n_match,
            (float)r.n_match / r.repeat_len, r.n_mismatch, r.n_ins, r.n_del, r.unit);
}

extern "C" void mtro_chain_run(mtro_chain *ch, FILE *out, int print_alignment, mtro_print_alignment_cb cb, void *user)
{
    if (ch->items.empty()) return;
    typedef std::multimap<int, Aln*> MM;
    MM by_x, by_y;
    for (Aln *a : ch->items) {
        if (a->start + 10 <= a->end) {
            by_x.insert(std::make_pair(a->start, a));
            by_x.insert(std::make_pair(a->end - 10, a));
        }
    }
    for (MM::iterator ev = by_x.begin(); ev != by_x.end(); ev++) {
        Aln *cur = ev->second;
        if (cur->start == ev->first) {                  /* start event, :270-289 */
            if (!by_y.empty()) {
                MM::iterator y, prev;
                const int lim = cur->start + 10;
                for (y = by_y.begin(), prev = y; y != by_y.end(); prev = y, y++) {
                    if (prev->second->end <= lim && y->second->end > lim) {
                        cur->pred = prev->second; cur->score += prev->second->score;
                        break;
                    }
                }
                if (prev->second->end <= lim && y == by_y.end()) {
                    cur->pred = prev->second; cur->score += prev->second->score;
                }
            }
        } else if (by_y.empty()) {                      /* end event, :290-333 */
            by_y.insert(std::make_pair(cur->end, cur));
        } else {
            bool keep = true;
            for (MM::iterator y = by_y.begin(); y != by_y.end(); y++) {
                if (y->second->end <= cur->end && y->second->score > cur->score) keep = false;
                if (y->second->end > cur->end) break;
            }
            if (keep) {
                by_y.insert(std::make_pair(cur->end, cur));
                for (MM::iterator y = by_y.begin(); y != by_y.end(); y++) {
                    if (y->second->end >= cur->end && y->second->score < cur->score)
                        y = by_y.erase(y);              /* followed by the loop's y++ (H2) */
                }
            }
        }
    }
    std::vector<const Aln*> chain;
    for (const Aln *a = by_y.rbegin()->second; a; a = a->pred) chain.push_back(a);
    for (size_t i = chain.size(); i-- > 0; ) {
        print_record(out, chain[i]);
        if (print_alignment == 1) { fputc('\n', out); cb(user, &chain[i]->rr); }
        fflush(out);
    }
    for (Aln *a : ch->items) delete a;
    ch->items.clear();
}
