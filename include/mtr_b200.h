/*
 * include/mtr_b200.h -- C ABI of the B200-native mTR hot path (libmtr_b200.so).
 *
 * Plain C: pointers, sizes and PODs only; no torch / C++ types.  Every entry point returns 0 on success
 * or a negative MTR_E* code (mtr_last_error() gives the text); nothing here falls back to the CPU -- if no
 * CUDA device is usable mtr_cuda_init fails and so does everything after it.
 *
 * Two layers:
 *   (1) the kernel-level ABI (mtr_cuda_*, mtr_reads_*, mtr_wdp_*, mtr_di_*): batched device work, used by
 *       the host pipeline below, by tests/ and by bench.py;
 *   (2) the reference's own entry points for this path, kept verbatim so that mTR's main.c links against
 *       this library unchanged: handle_one_file / handle_one_read and the globals main.c sets and prints
 *       (/root/reference/mTR.h:61-62,126-127,142-143; main.c:53-56,93-121).
 *
 * Each declaration cites the reference interface it replaces (file:line under /root/reference).
 */
#ifndef MTR_B200_H
#define MTR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ errors */
enum {
    MTR_OK = 0,
    MTR_ENODEV = -1,     /* no usable CUDA device / CUDA runtime error at init */
    MTR_ECUDA = -2,      /* a CUDA call failed; see mtr_last_error */
    MTR_EINVAL = -3,     /* bad argument (unit length outside 1..499, rows < 0, ...) */
    MTR_ENOMEM = -4,     /* device or host allocation failed */
    MTR_ERANGE = -5      /* a job exceeds the reference's own caps (WrapDPsize, MAX_INPUT_LENGTH) */
};

typedef struct mtr_ctx mtr_ctx;          /* one per GPU; not thread-safe, one caller thread per ctx */

/* Replaces malloc_global_variables / free_global_variables (handle_one_file.c:71-167): all device and
 * pinned-host buffers live in the context and grow on demand. */
int  mtr_cuda_init(int device, mtr_ctx **out);
void mtr_cuda_shutdown(mtr_ctx *ctx);
const char *mtr_last_error(const mtr_ctx *ctx);   /* ctx may be NULL: error of the last failed init */
int  mtr_device_count(void);
/* on != 0: host threads waiting for this context sleep (cudaEventBlockingSync) instead of spinning; worth it for
 * contexts whose calls take milliseconds (adds tens of microseconds of wake-up latency per call). Default off. */
int  mtr_set_blocking_sync(mtr_ctx *ctx, int on);
/* CUDA stream priority of everything this context launches: level 0 = lowest (bulk work), higher = more urgent (clamped
 * to the device's range).  The pipeline gives its short-job DP lanes a higher level than the long-job lanes and the
 * directional index.  Call before the context has work in flight. */
int  mtr_set_priority(mtr_ctx *ctx, int level);

/* ------------------------------------------------------------------ reads (orgInputString, mTR.h:65) */
/* A batch of reads, 2-bit packed (A,C,G,T = 0..3; base b of a read is bits [2*(b%16), 2*(b%16)+2) of word
 * b/16, counted from that read's first word).  word_off[r] is the index of read r's first word in `packed`
 * (word_off[n_reads] = total words); each read's words must cover len[r] + 2 bases: positions len[r] and
 * len[r]+1 hold the two "stale" bases the reference reads past the end of the read (SURVEY.md 4.3 H4,
 * wrap_around_DP.c:243-245,264).  The batch stays resident until the next upload. */
int mtr_reads_upload(mtr_ctx *ctx, const uint32_t *packed, const int64_t *word_off, const int32_t *len,
                     int n_reads);

/* Makes dst use the read batch resident in src (same device) without copying it; dst does not own the memory,
 * so src must outlive dst's use of the batch.  Lets several contexts run DP batches concurrently on one GPU. */
int mtr_reads_share(mtr_ctx *dst, const mtr_ctx *src);

/* ------------------------------------------------------------------ K3: wrap-around DP */
/* One job = one call of wrap_around_DP_sub (wrap_around_DP.c:222-354), of the DP inside
 * revise_representative_unit_sub (consensus.c:851-962, mode CONSENSUS) or of pretty_print_alignment
 * (wrap_around_DP.c:57-186, mode PATH); with n_param == 2 it is one call of wrap_around_DP
 * (wrap_around_DP.c:357-429): the same window and unit under two penalty sets.
 *
 * Rows i = 1..rows are the read bases x_i = read[first + i]  (so first = query_start for
 * wrap_around_DP_sub, whose window is orgInputString[query_start+1 .. query_end+1], and first = rep_start-1
 * for pretty_print_alignment); columns j = 1..ulen are units[unit_off + j - 1] (one byte per base, 0..3). */
enum { MTR_TB_COUNTS = 0, MTR_TB_CONSENSUS = 1, MTR_TB_PATH = 2 };

typedef struct {
    int32_t read;          /* index into the resident read batch */
    int32_t first;         /* see above; may be -1 (then row 1 is base 0) */
    int32_t rows;          /* rep_len = query_end - query_start + 1 */
    int32_t unit_off;      /* offset of the unit in the `units` byte array */
    int32_t ulen;          /* 1..499 (MAX_PERIOD-1, mTR.h:34) */
    int8_t  gain[2], mis[2], indel[2];   /* MATCH_GAIN, MISMATCH_PENALTY, INDEL_PENALTY per penalty set */
    uint8_t n_param;       /* 1 or 2 */
    uint8_t mode;          /* MTR_TB_*; CONSENSUS and PATH need n_param == 1 */
    int64_t aux_off;       /* CONSENSUS: offset (in int32) of this job's (ulen+1)*9 histogram block in aux;
                              PATH: offset (in bytes) of this job's path block in aux; else ignored */
    int64_t aux_cap;       /* PATH: capacity in bytes of the path block */
} mtr_wdp_job;

/* The fields wrap_around_DP_sub derives its record from (wrap_around_DP.c:337-350):
 * rep_start = query_start + end_i + 1, rep_end = query_start + max_i, repeat_len = max_i - end_i,
 * Num_freq_unit = n_scanned / ulen. */
typedef struct {
    int32_t best;          /* max_wrd: maximum cell value (0 => max_i = max_j = 0, empty traceback) */
    int32_t max_i, max_j;  /* first row-major argmax */
    int32_t end_i, end_j;  /* where the traceback stopped */
    int32_t n_match, n_mismatch, n_ins, n_del, n_scanned;
    int32_t path_len;      /* number of traceback steps (PATH: bytes written, clipped to aux_cap) */
    int32_t flags;         /* bit 0: path clipped */
} mtr_wdp_result;

/* CONSENSUS aux block layout per job: int32 consensus[(ulen+1)][5] then int32 missing[(ulen+1)][4]
 * (the two histograms of consensus.c:919-962, row j = unit position, 1-origin).
 * PATH aux block: one byte per traceback step, in traceback order (from the alignment's end):
 * 0 match, 1 mismatch, 2 deletion (unit base, no read base), 3 insertion (read base, no unit base). */

/* Synchronous: H2D(jobs, units) -> fill kernels -> traceback kernel -> D2H(results[, aux]).
 * results has n_jobs * 2 entries: results[2*j + p] for penalty set p (p = 1 unused when n_param == 1).
 * aux / aux_bytes may be NULL / 0 when no job asks for CONSENSUS or PATH output. */
int mtr_wdp_run(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                mtr_wdp_result *results, void *aux, int64_t aux_bytes);

/* The same in three phases, for pipelining and for timing the kernels with the operands resident in HBM:
 * upload (H2D + job classification), launch (kernels only; may be repeated), download (D2H). */
int mtr_wdp_upload(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                   int64_t aux_bytes);
int mtr_wdp_launch(mtr_ctx *ctx);
int mtr_wdp_download(mtr_ctx *ctx, mtr_wdp_result *results, void *aux, int64_t aux_bytes);
/* on != 0 (default): every fill kernel runs the traceback of a task as soon as that task's fill is done (no second
 * launch, the direction rows are still in cache); on == 0: fill kernels, then one traceback kernel -- lets the fill
 * be timed alone (mtr_stats.wdp_fill_ms / wdp_tb_ms; fused: wdp_fill_ms covers both, wdp_tb_ms ~ 0). */
int mtr_wdp_set_fused_traceback(mtr_ctx *ctx, int on);

/* ------------------------------------------------------------------ K1/K2: directional index */
/* Replaces fill_directional_index_with_end (fill_directional_index.c:549-602) for every read of the
 * resident batch: k-mer coding with the MT19937 flanks (:137-169), the sliding three-window distance
 * (Manhattan :171-295, or Pearson :298-450 when manhattan == 0), the local max/min merge (:467-503), the
 * unshift (:587-597) and remove_redundant_ranges (:505-546).
 *
 * stale / stale_off: per read, the values the reference would find in inputString_w_rand beyond the area it
 * re-initialises for this read (index >= len + 4r); stale_off[r]..stale_off[r+1] delimit
 * read r's uint16 values, missing values read as 0 (fresh process).  May be NULL (all zero).
 *
 * Output, per read r, for positions [0, len[r]): di (fp64), end, w -- exactly the reference's
 * directional_index / directional_index_end / directional_index_w after the call; pos_off[r] is the offset
 * of read r in the three arrays (pos_off[n] = total).  di may be NULL: a range is live iff end >= 0 (the pipeline
 * only needs end and w). */
int mtr_di_run(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off,
               const int64_t *pos_off, double *di, int32_t *end, int32_t *w);
/* The same for reads [first, first + count) of the resident batch only; every array is indexed as for the whole batch.
 * Lets the host sweep a batch in slices and start on a slice while the next one is on the GPU. */
int mtr_di_run_range(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off,
                     const int64_t *pos_off, double *di, int32_t *end, int32_t *w, int first, int count);

/* ------------------------------------------------------------------ K4: unit finder */
/* One task = search_De_Bruijn_graph (consensus.c:507-549) for one candidate range [qs, qe] of one resident read
 * and one k, up to the point where it would call wrap_around_DP: k-mer counts of the window (init_inputString
 * :37-60, generate_freqNode_return_list_maxNodes :132-229), then the greedy forward walk over the maximum-
 * frequency nodes until the first loop, then the backward walk likewise (:269-505). */
typedef struct {
    int32_t read, qs, qe, k;   /* k in 2..15 */
} mtr_uf_task;

typedef struct {
    int32_t max_freq;          /* maxFreq; the walks run only if it exceeds MIN_NUM_FREQ_UNIT = 5 (:532) */
    int32_t found_last;        /* foundLoop of the LAST walk attempted: search_De_Bruijn_graph's return value (Q4) */
    int32_t found[2];          /* forward, backward: a loop was found */
    int32_t period[2];         /* its length (1..499) */
    int64_t unit_off[2];       /* offset of the unit (bytes 0..3) in `units` and of its node counts
                                  (string_score, :350-352) in `scores`; -1 if not found */
} mtr_uf_result;

/* Synchronous.  units / scores must hold out_cap entries with out_cap >= sum over tasks of 2*min(500,(qe-qs)/5);
 * *out_used receives the number of entries written. */
int mtr_uf_run(mtr_ctx *ctx, const mtr_uf_task *tasks, int n_tasks, mtr_uf_result *results,
               uint8_t *units, int32_t *scores, int64_t out_cap, int64_t *out_used);

/* ------------------------------------------------------------------ resident engine: the per-read path on the device */
/* handle_one_TR (handle_one_read.c:190-261) for EVERY read of the resident batch, without the host in the loop:
 * directional index (fill_directional_index_with_end), the candidate loop with its pruning (:227-246),
 * find_tandem_repeat over the k range (:102-154), search_De_Bruijn_graph (consensus.c:507-582: k-mer counts, greedy
 * walks, wrap_around_DP under both penalty sets), polish_repeat and revise_representative_unit (:584-704, 1048-1087).
 * The reads advance in waves of kernels; the host only launches waves until every read is finished.
 *
 * The result is what the reference hands to insert_an_alignment_into_set (mTR.h:146-165, handle_one_read.c:156-176),
 * one record per accepted repeat, ordered by (read, order of insertion); chaining and printing (chaining.cpp) remain
 * with the caller.  Arrays returned through the out-pointers belong to the context and stay valid until the next
 * call on it.  stale / stale_off: as for mtr_di_run. */
typedef struct {
    int32_t read;              /* index into the resident batch */
    int32_t seq;               /* order of insertion within the read */
    int32_t rep_start, rep_end, repeat_len, rep_period, num_freq_unit;
    int32_t num_matches, num_mismatches, num_insertions, num_deletions;
    int32_t kmer, match_gain, mismatch_penalty, indel_penalty;
    int32_t pad_;
    int64_t unit_off;          /* offset of the unit (rep_period bytes, values 0..3) in the units array */
} mtr_repeat;

typedef struct {
    int64_t waves;             /* kernel waves launched for the group */
    int64_t candidates;        /* query_counter: candidate ranges the reference would have visited */
    int64_t dp_jobs, dp_tasks; /* wrap-around DP calls (a two-penalty call is one job) / fill tasks launched */
    int64_t dp_cells;          /* sum rows * ulen over every DP launched (both penalty sets) */
    int64_t dp_slot_cells, dp_dir_bytes;
    int64_t spec_cells;        /* part of dp_cells run for look-ahead candidates that were pruned after all */
    int64_t dp_cells_p16;      /* part of dp_cells run by the paired int16x2 kernels */
    int64_t shared_cells;      /* cells of DPs the reference runs that were NOT launched: another k (or walk direction) of the same
                                  candidate had found the identical unit, and the identical DP was run once (not in dp_cells) */
    int64_t tables, table_positions, walks;   /* k-mer count tables built, positions counted, greedy walks run */
    int64_t repeats;
    int64_t wrapdp_messages;   /* "You need to increse the value of WrapDPsize." lines (handle_one_read.c:89-91) */
    int64_t launches;          /* CUDA kernels launched */
    int64_t h2d_bytes, d2h_bytes;
    double  di_ms;             /* CUDA-event time of the directional-index kernels */
    double  dp_ms;             /* CUDA-event time of the K3 fill + traceback kernels, summed over the waves */
    double  uf_ms;             /* ... of the scheduler / unit-finder / polish kernels */
    double  wall_ms;           /* host wall clock of the whole call */
} mtr_engine_stats;

int mtr_engine_run(mtr_ctx *ctx, int manhattan, float min_match_ratio, const uint16_t *stale, const int64_t *stale_off,
                   const mtr_repeat **repeats, int64_t *n_repeats, const uint8_t **units, mtr_engine_stats *stats);
/* The same for reads [first, first + count) of the resident batch only (mtr_repeat::read stays an index into the whole
 * batch; stale_off is indexed like the whole batch): several groups of reads can be resident in one context and run one
 * after the other. */
int mtr_engine_run_range(mtr_ctx *ctx, int first, int count, int manhattan, float min_match_ratio, const uint16_t *stale,
                         const int64_t *stale_off, const mtr_repeat **repeats, int64_t *n_repeats, const uint8_t **units,
                         mtr_engine_stats *stats);
/* MTR_SPECULATE-style knobs of the engine (defaults: 8 look-ahead candidates). */
int mtr_engine_set_speculate(mtr_ctx *ctx, int depth);
/* Union, over every engine call of the process since the last reset, of the time intervals in which K3 kernels of
 * ANY context were running on `device` (milliseconds): the in-pipeline kernel time of the DP roofline when several
 * groups share one GPU.  reset != 0 clears the record after reading it. */
double mtr_engine_dp_busy_ms(int device, int reset);

/* ------------------------------------------------------------------ counters for the roofline */
typedef struct {
    double   wdp_fill_ms, wdp_tb_ms, di_ms;      /* CUDA-event time of the last launch of each stage */
    int64_t  wdp_cells;                          /* sum over jobs and penalty sets of rows * ulen */
    int64_t  wdp_slot_cells;                     /* cells actually computed incl. padding lanes */
    int64_t  wdp_dir_bytes;                      /* direction bytes written by the last launch */
    int64_t  di_position_passes;                 /* sum over passes of (L + r - w - k + 1) */
    int64_t  di_bytes_in, di_bytes_out;
    int32_t  launches;                           /* kernels launched by the last call */
    int32_t  n_sm;
    double   uf_ms;                              /* CUDA-event time of the last unit-finder launch */
    int64_t  uf_tasks, uf_table_bytes;
} mtr_stats;
int mtr_get_stats(const mtr_ctx *ctx, mtr_stats *out);

/* Integer-ALU issue microbenchmark: giga lane-operations per second of VIADDMNMX.RELU (kind 0), LOP3/IADD
 * (kind 1) and VIADDMNMX.S16x2 (kind 2) with every SM busy -- the denominator of the DP roofline. */
int mtr_alu_probe(mtr_ctx *ctx, int kind, double *gops);

/* ------------------------------------------------------------------ batch-level pipeline */
/* The whole per-read path of handle_one_file (handle_one_file.c:271-293 -> handle_one_read.c:190-261) on FASTA
 * text that is already in host memory, split so that the host<->device boundary can be timed separately:
 *   load_fasta : parse + stale-state tracking + 2-bit pack + H2D, one group of reads per engine context
 *                (MTR_GROUPS_PER_GPU contexts; afterwards the batch is resident in HBM)
 *   run        : mtr_engine_run on every group concurrently, then chaining and formatting on the host;
 *                returns the text the reference would print for these reads (TSV records, alignments with -a)
 * Uses the globals Manhattan_Distance and min_match_ratio like the reference. */
typedef struct mtr_pipeline mtr_pipeline;
typedef struct {
    int64_t reads, bases, groups;
    int64_t candidates;                                    /* query_counter */
    int64_t waves;                                         /* engine waves, summed over the groups */
    int64_t jobs, dp_tasks;                                /* wrap-around DP calls / K3 fill tasks */
    int64_t wdp_cells, wdp_slot_cells, wdp_dir_bytes;      /* cells = sum rows * ulen over every DP launched */
    int64_t spec_cells;                                    /* part of wdp_cells run for look-ahead candidates that were pruned
                                                              after all (MTR_SPECULATE, default 8): not algorithmic cells */
    int64_t wdp_cells_p16;                                 /* part of wdp_cells run by the paired int16x2 kernels */
    int64_t shared_cells;                                  /* algorithmic cells NOT launched: identical search DPs of one candidate
                                                              (same window, same unit found for another k) run once */
    int64_t tables, table_positions, walks, repeats;
    int64_t h2d_bytes, d2h_bytes;
    int64_t launches;                                      /* CUDA kernels launched */
    double  dp_ms, di_kernel_ms, uf_kernel_ms;             /* CUDA-event time on the launching streams, summed over groups */
    double  engine_wall_ms;                                /* wall clock inside mtr_engine_run, summed over groups */
    double  pack_ms, chain_ms;                             /* host: 2-bit pack + upload / chaining + formatting, summed */
    double  wall_ms;                                       /* mtr_pipeline_run: wall clock of the call */
} mtr_pipeline_stats;
int  mtr_pipeline_open(int device, int threads, mtr_pipeline **out);
void mtr_pipeline_close(mtr_pipeline *p);
int  mtr_pipeline_load_fasta(mtr_pipeline *p, const char *text, int64_t len);
/* Sharding: keep reads [first, first+count) only (count < 0: to the end); earlier reads are parsed just to carry the
 * reference's cross-read stale state, so concatenating the shards' outputs equals the whole file's output. */
int  mtr_pipeline_load_fasta_shard(mtr_pipeline *p, const char *text, int64_t len, int first, int count);
int  mtr_pipeline_run(mtr_pipeline *p, int print_alignment, const char **out_text, int64_t *out_len);
int  mtr_pipeline_get_stats(const mtr_pipeline *p, mtr_pipeline_stats *out);
mtr_ctx *mtr_pipeline_ctx(mtr_pipeline *p);

/* ------------------------------------------------------------------ the reference's entry points */
/* mTR.h:126  handle_one_file(char *inputFile, int print_alignment): parses the FASTA, runs the pipeline on
 * the GPUs named by MTR_GPUS (default: device 0), prints the records to stdout in input order; returns the
 * number of reads.  Reads the globals below, like the reference. */
int  handle_one_file(char *inputFile, int print_alignment);
/* mTR.h:127  handle_one_read: the read is orgInputString[0..inputLen) (int per base), as in the reference
 * (handle_one_file.c:284-286).  Here it enqueues a copy; mtr_flush() completes and prints everything
 * enqueued so far, in order. */
void handle_one_read(char *readID, int inputLen, int read_cnt, int print_alignment);
void mtr_flush(void);
/* Counters (reads, cells, kernel launches, host<->device bytes ...) summed over everything the three entry points above
 * have processed since the previous call; resets them.  The reference has only the -c timers (mTR.h:142-143). */
int  mtr_file_stats(mtr_pipeline_stats *out);

/* ------------------------------------------------------------------ optional post-pass: cross-read clustering */
/* the reference's k_means_clustering.c:136-355 restated over the emitted TSV records (that file is dead
 * code in the reference -- not built, does not compile -- so this has no binary to be compared with; cluster.cpp states what
 * is restated and which choices the source leaves open).  tsv: records as bin/mTR prints them (other lines are skipped).
 * Output (malloc'ed, free with mtr_cluster_free): one line per qualified record, "<id of the representative>\t<record>" with
 * the representative's unit length and unit string in place of the record's own, clusters by falling size.
 * params == NULL: min_match_ratio 0.6 (MIN_MATCH_RATIO, mTR.h:32), mh_distance_threshold 0.3, min_rep_len 0, min_num_rep 1. */
typedef struct {
    double min_match_ratio;        /* MIN_MATCH_RATIO < matches / repeat_len */
    double mh_distance_threshold;  /* sum |d 2-mer| <= threshold * unit length of the representative (TRs_in_neighborhood) */
    int    min_rep_len;            /* MIN_REP_LEN < unit length * number of units */
    int    min_num_rep;            /* MIN_NUM_repTR; only 1 is defined by the source */
} mtr_cluster_params;
int  mtr_cluster_records(const char *tsv, int64_t len, const mtr_cluster_params *params, char **out_text, int64_t *out_len);
void mtr_cluster_free(char *text);

/* ------------------------------------------------------------------ the reference's C <-> C++ bridge */
/* mTR.h:146-175 / chaining.h:30-56: the interface between handle_one_read.c and chaining.cpp, for a caller that keeps
 * its own per-read code above this library.  The library's own per-read loop runs on the device and never goes
 * through these four symbols.
 *   insert_an_alignment_into_set  chaining.cpp:203-241: adds one repeat of the current read to the set (the set is kept
 *                                 in insertion order: the canonical order of the reference's pointer-ordered std::set)
 *   chaining                      chaining.cpp:243-363: best chain over the set by the reference's own sweep, prints its
 *                                 records to stdout (print_one_TR), print_alignment == 1: a blank line and
 *                                 pretty_print_alignment after each; empties the set
 *   pretty_print_alignment        wrap_around_DP.c:57-213: rows = orgInputString[rep_start..rep_end]; the DP and its
 *                                 traceback run on the GPU (one K3 job, mode MTR_TB_PATH), the text is the reference's
 *   print_freq                    consensus.c:1089-1131 (debugging aid): k-mer counts of the unit's rotations in the
 *                                 window, one digit per position                                                   */
void insert_an_alignment_into_set(char *readID, int inputLen, int rep_start, int rep_end, int repeat_len, int rep_period,
                                  int Num_freq_unit, int Num_matches, int Num_mismatches, int Num_insertions,
                                  int Num_deletions, int Kmer, int match_gain, int mismatch_penalty, int indel_penalty,
                                  char *string, int *string_score);
void chaining(int print_alignment);
void pretty_print_alignment(char *unit_string, int unit_len, int rep_start, int rep_end, int match_gain,
                            int mismatch_penalty, int indel_penalty);
void print_freq(int rep_start, int rep_end, int rep_period, char *string, int inputLen, int k);

extern int   Manhattan_Distance;      /* mTR.h:61, set by main.c:56,77 */
extern float min_match_ratio;         /* mTR.h:62, set by main.c:54,68 */
extern int  *orgInputString;          /* mTR.h:65 */
extern float time_all, time_memory, time_range, time_period, time_initialize_input_string,
             time_wrap_around_DP, time_count_table, time_chaining;      /* mTR.h:142 */
extern int   query_counter;           /* mTR.h:143 */

#ifdef __cplusplus
}
#endif
#endif
