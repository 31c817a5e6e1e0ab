/*
 * oracle/mtr_oracle.h -- CPU restatement of mTR's per-read tandem-repeat pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from the
 * product (mtr_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement against
 *   (a) the 45 known-answer digests of SURVEY.md App. C (15 shipped FASTA files x 3 modes),
 *   (b) outputs of the unmodified reference compiled by oracle/Makefile (oracle/_ref/mTR_ref_det,
 *       libmtr_ref.so) on seeded synthetic reads, committed as fixtures under tests/golden/.
 *
 * Every function cites the reference file:line (under /root/reference) whose behaviour it restates.
 * The code is written from the semantics in SURVEY.md App. A, without globals: all state that the
 * reference keeps in global arrays lives in mtro_ctx, and deliberately persists from read to read,
 * because the reference's output depends on it (hazards H3/H4 of SURVEY.md 4.3).
 */
#ifndef MTR_ORACLE_H
#define MTR_ORACLE_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MTRO_MAX_INPUT_LENGTH 1000000   /* mTR.h:31 */
#define MTRO_MAX_PERIOD       500       /* mTR.h:34 */
#define MTRO_UNIT_CAP         1024      /* room for the 2x overshoot of consensus.c:965 (H9) */

/* The per-repeat feature record (mTR.h:99-119) without the fields that never reach the output. */
typedef struct {
    int inputLen, rep_start, rep_end, repeat_len, rep_period, n_units;
    int n_match, n_mismatch, n_ins, n_del, kmer, gain, mis_pen, indel_pen;
    char unit[MTRO_UNIT_CAP];
    int  unit_score[MTRO_MAX_PERIOD];
} mtro_rr;

/* Result of one wrap-around DP (wrap_around_DP.c:222-354). */
typedef struct {
    int best;                 /* maximum cell value */
    int max_i, max_j;         /* first row-major argmax; (0,0) when best == 0 */
    int end_i, end_j;         /* cell where the traceback stopped */
    int n_match, n_mismatch, n_ins, n_del, n_scanned;
    long long cells;          /* rows * unit_len */
    int path_len;             /* number of traceback steps (PATH mode: entries written) */
} mtro_dp_result;

enum { MTRO_TB_COUNTS = 0, MTRO_TB_CONSENSUS = 1, MTRO_TB_PATH = 2 };

typedef struct mtro_ctx mtro_ctx;

mtro_ctx *mtro_new(int manhattan, float min_match_ratio);
void      mtro_free(mtro_ctx *c);
void      mtro_set_output(mtro_ctx *c, FILE *out);

/* Whole-file / whole-read drivers (handle_one_file.c:271-293, handle_one_read.c:190-266).
 * mtro_process_file returns the number of reads, like handle_one_file. */
int  mtro_process_file(mtro_ctx *c, const char *path, int print_alignment);
void mtro_process_read(mtro_ctx *c, const char *read_id, const int *bases, int len, int print_alignment);

/* Stage-level entry points used by the kernel parity tests. */

/* Directional index of the read currently loaded with mtro_load_read (fill_directional_index.c:549-602).
 * Outputs have length len: DI (fp64), END, W; entries are -1 where no candidate starts. */
void mtro_load_read(mtro_ctx *c, const int *bases, int len);
void mtro_directional_index(mtro_ctx *c, double *di, int *end, int *w);
/* The padded k-mer coded string S of one k (fill_directional_index.c:137-169), first n entries. */
void mtro_padded_codes(mtro_ctx *c, int k, int *out, int n);
/* Raw per-pass DI stream (fill_directional_index.c:171-295 / 298-450); tmp has N = L + 2r entries. */
void mtro_di_pass(mtro_ctx *c, int k, int w, double *tmp);

/* One wrap-around DP over rows x[1..rows] and unit u[1..ulen] (both 1-origin int arrays, values 0..3).
 *  mode COUNTS    : counts only
 *  mode CONSENSUS : also fills consensus[(ulen+1)*5] and missing[(ulen+1)*4] (consensus.c:919-962)
 *  mode PATH      : also writes path ops, one byte per step in traceback order:
 *                   0 = match, 1 = mismatch, 2 = deletion (unit base, no read base), 3 = insertion
 *  dirs (optional, may be NULL): (rows+1)*(ulen+1) bytes, the traceback decision of EVERY cell:
 *                   0 = diagonal, 1 = left (deletion), 2 = up (insertion), 3 = cell value is 0 (stop)  */
void mtro_wrap_dp(mtro_ctx *c, const int *x, int rows, const int *u, int ulen,
                  int gain, int mis_pen, int indel_pen, int mode,
                  mtro_dp_result *res, int *consensus, int *missing,
                  unsigned char *path, unsigned char *dirs);

/* Unit finder for one candidate range of the loaded read (handle_one_read.c:102-154). */
void mtro_find_tandem_repeat(mtro_ctx *c, int query_start, int query_end, int w, mtro_rr *out);

/* The part of search_De_Bruijn_graph (consensus.c:507-549) that precedes wrap_around_DP, for the loaded read:
 * counts, maximum-frequency node list, first looping forward walk and first looping backward walk.
 * unit[d] / score[d] must hold 500 entries each. */
typedef struct { int max_freq, found_last, found[2], period[2]; } mtro_walk_result;
void mtro_unit_walks(mtro_ctx *c, int query_start, int query_end, int k, mtro_walk_result *res,
                     unsigned char *unit_fwd, int *score_fwd, unsigned char *unit_bwd, int *score_bwd);

/* min_missing table lookup (consensus.c:714-820). */
int mtro_min_missing(int rep_period, double error, int coverage);
int mtro_min_missing_raw(int i, int j, int k);

/* TRs_in_neighborhood / cmp_TR of the dead k_means_clustering.c (:169-180, :62-101), restated for
 * completeness. */
int mtro_trs_in_neighborhood(const int *freq2mer_a, const int *freq2mer_rep, int rep_period);   /* +1 / -1 */
int mtro_cmp_tr(int period_a, const int *f2_a, int units_a, int rep_freq_a, int rep_id_a,
                int period_b, const int *f2_b, int units_b, int rep_freq_b, int rep_id_b, int mode);
void mtro_freq_2mer(const int *unit, int len, int *freq16);

/* Counters for the roofline denominators. */
typedef struct {
    long long dp_calls, dp_cells;          /* wrap_around_DP_sub */
    long long revise_calls, revise_cells;  /* revise_representative_unit_sub */
    long long print_calls, print_cells;    /* pretty_print_alignment */
    long long di_position_passes;          /* sum over passes of (L + r - w - k + 1) */
    long long candidates, searches;        /* query_counter, (range,k) searches */
    long long reads, bases;
} mtro_stats;
void mtro_get_stats(const mtro_ctx *c, mtro_stats *s);

/* Optional hook: called for every wrap_around_DP_sub / revise DP the pipeline executes, so tests and the
 * bench can harvest realistic DP job batches.  kind: 0 = counts DP, 1 = revise (consensus) DP, 2 = print. */
typedef void (*mtro_dp_hook)(void *user, int kind, const int *x, int rows, const int *u, int ulen,
                             int gain, int mis_pen, int indel_pen, const mtro_dp_result *res);
void mtro_set_dp_hook(mtro_ctx *c, mtro_dp_hook hook, void *user);
/* Base of the persistent read buffer (orgInputString); a hook's x pointer minus this is the job's `first`. */
const int *mtro_org(const mtro_ctx *c);
/* Mutable views of the persistent buffers orgInputString (MTRO_MAX_INPUT_LENGTH + 16 ints) and inputString_w_rand
 * (3 * MTRO_MAX_INPUT_LENGTH + 64 ints): lets a test put the cross-read stale state in place
 * explicitly instead of replaying the earlier reads. */
int *mtro_org_mut(mtro_ctx *c);
int *mtro_padded_mut(mtro_ctx *c);

/* ---- chaining (C++ side, mtr_oracle_chain.cpp; chaining.cpp:43-363) ---- */
typedef struct mtro_chain mtro_chain;
mtro_chain *mtro_chain_new(void);
void mtro_chain_free(mtro_chain *ch);
void mtro_chain_insert(mtro_chain *ch, const char *read_id, const mtro_rr *rr);
/* Runs the sweep, prints the best chain to out, empties the set.  print_cb prints one alignment (-a). */
typedef void (*mtro_print_alignment_cb)(void *user, const mtro_rr *rr);
void mtro_chain_run(mtro_chain *ch, FILE *out, int print_alignment, mtro_print_alignment_cb cb, void *user);

#ifdef __cplusplus
}
#endif
#endif
