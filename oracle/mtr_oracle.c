/*
 * oracle/mtr_oracle.c -- CPU restatement of mTR's per-read pipeline.  TEST INFRASTRUCTURE ONLY
 * (see mtr_oracle.h).  Parity: pinned against the reference binary and the App. C digests.
 *
 * Written from the semantics condensed in SURVEY.md App. A.  Reference citations are file:line under
 * /root/reference.  State the reference keeps in never-cleared global arrays is kept in mtro_ctx and
 * is never cleared between reads either, because the output depends on it (H3/H4).
 */
#define _POSIX_C_SOURCE 200809L
#include "mtr_oracle.h"
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXLEN   MTRO_MAX_INPUT_LENGTH
#define MAXP     MTRO_MAX_PERIOD
#define WRAPCAP  200000000            /* WrapDPsize, mTR.h:51 */
#define HISTBINS 16384                /* 4*BLK, handle_one_file.c:107-112 */
#define MAXTIES  1024                 /* MAX_tiebreaks, mTR.h:46 */
#define MT_TABLE_LEN 1300000          /* >= MAXLEN + 2 * MAXLEN/10 */

struct mtro_ctx {
    int manhattan;
    float min_match_ratio;
    FILE *out;
    /* persistent, never cleared between reads (mTR.h:65-67) */
    int *org;        /* orgInputString */
    int *kstr;       /* inputString: k-mer codes of the current search window */
    int *padded;     /* inputString_w_rand */
    int cur_len;
    /* directional index */
    double *di_tmp, *di;
    int *di_end, *di_w;
    int *h0, *h1, *h2;
    unsigned char *mt_base;           /* genrand_int32() % 4 stream after init_genrand(0) */
    /* k-mer counts */
    int *count;                       /* direct table, 4^6 */
    int *hkey, *hval; int hcap;       /* exact counts for k > 6 (layout unobservable) */
    int count_k;
    /* DP matrix, persistent like WrapDP (row 0 is only partly re-initialised, wrap_around_DP.c:250) */
    int *W; size_t W_cap;
    mtro_chain *chain;
    const char *cur_id;
    mtro_stats st;
    mtro_dp_hook hook; void *hook_user;
    int pow4[16];
};

/* ---------------------------------------------------------------- MT19937 (MT.h:65-78, 110-145) */
static void mt_fill_table(unsigned char *dst, int n)
{
    uint32_t mt[624];
    int idx = 624;
    mt[0] = 0u;                                        /* init_genrand(0) */
    for (int i = 1; i < 624; i++)
        mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    for (int t = 0; t < n; t++) {
        if (idx >= 624) {
            for (int kk = 0; kk < 624; kk++) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        dst[t] = (unsigned char)(y % 4u);
    }
}

/* ---------------------------------------------------------------- context */
static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "mtr_oracle: out of memory\n"); exit(EXIT_FAILURE); }
    return p;
}

mtro_ctx *mtro_new(int manhattan, float min_match_ratio)
{
    mtro_ctx *c = xcalloc(1, sizeof *c);
    c->manhattan = manhattan;
    c->min_match_ratio = min_match_ratio;
    c->out = stdout;
    /* The reference mallocs MAX_INPUT_LENGTH ints and reads past the end for long reads; fresh malloc'ed
     * pages are zero, so zero-filled, generously sized buffers reproduce a fresh process. */
    c->org    = xcalloc(MAXLEN + 16, sizeof(int));
    c->kstr   = xcalloc(MAXLEN + 16, sizeof(int));
    c->padded = xcalloc(3 * (size_t)MAXLEN + 64, sizeof(int));
    c->di_tmp = xcalloc(2 * (size_t)MAXLEN, sizeof(double));
    c->di     = xcalloc(2 * (size_t)MAXLEN, sizeof(double));
    c->di_end = xcalloc(2 * (size_t)MAXLEN, sizeof(int));
    c->di_w   = xcalloc(2 * (size_t)MAXLEN, sizeof(int));
    c->h0 = xcalloc(HISTBINS, sizeof(int));
    c->h1 = xcalloc(HISTBINS, sizeof(int));
    c->h2 = xcalloc(HISTBINS, sizeof(int));
    c->mt_base = xcalloc(MT_TABLE_LEN, 1);
    mt_fill_table(c->mt_base, MT_TABLE_LEN);
    c->count = xcalloc(4096, sizeof(int));
    c->hcap = 0;
    c->W = NULL; c->W_cap = 0;
    c->chain = mtro_chain_new();
    c->pow4[0] = 1;
    for (int i = 1; i < 16; i++) c->pow4[i] = c->pow4[i - 1] * 4;
    return c;
}

void mtro_free(mtro_ctx *c)
{
    if (!c) return;
    free(c->org); free(c->kstr); free(c->padded); free(c->di_tmp); free(c->di); free(c->di_end);
    free(c->di_w); free(c->h0); free(c->h1); free(c->h2); free(c->mt_base); free(c->count);
    free(c->hkey); free(c->hval); free(c->W);
    mtro_chain_free(c->chain);
    free(c);
}

void mtro_set_output(mtro_ctx *c, FILE *out) { c->out = out; }
void mtro_get_stats(const mtro_ctx *c, mtro_stats *s) { *s = c->st; }
void mtro_set_dp_hook(mtro_ctx *c, mtro_dp_hook hook, void *user) { c->hook = hook; c->hook_user = user; }
const int *mtro_org(const mtro_ctx *c) { return c->org; }
int *mtro_org_mut(mtro_ctx *c) { return c->org; }
int *mtro_padded_mut(mtro_ctx *c) { return c->padded; }

static void rr_clear(mtro_rr *r)            /* clear_rr, fill_directional_index.c:40-60 */
{
    r->inputLen = r->rep_start = r->rep_end = r->repeat_len = r->rep_period = r->n_units = -1;
    r->n_match = r->n_mismatch = r->n_ins = r->n_del = r->kmer = -1;
    r->gain = r->mis_pen = r->indel_pen = -1;
    r->unit[0] = '\0';
    for (int i = 0; i < MAXP; i++) r->unit_score[i] = -1;
}

static float rr_ratio(const mtro_rr *r)     /* the (float)M/(M+X+I+D) idiom, e.g. wrap_around_DP.c:398 */
{
    return (float)r->n_match / (r->n_match + r->n_mismatch + r->n_ins + r->n_del);
}

static int base_of_char(char ch)
{
    switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    fprintf(stderr, "mtr_oracle: fatal unit char %c\n", ch);
    exit(EXIT_FAILURE);
}

/* ================================================================ directional index */

static int random_len_of(int len)           /* handle_one_read.c:194-202 */
{
    return len < 1000 ? 100 : len / 10;
}

/* init_inputString_surrounded_by_random_seq, fill_directional_index.c:137-169 */
static void build_padded(mtro_ctx *c, int k)
{
    const int L = c->cur_len, r = random_len_of(L);
    int *S = c->padded;
    int cpos = 0;
    for (int i = 0; i < L + 4 * r && i < MAXLEN; i++) S[i] = c->mt_base[cpos++];
    for (int i = 0; i < r; i++) S[i] = c->mt_base[cpos++];
    for (int i = 0; i < L; i++) S[r + i] = c->org[i];
    for (int i = 0; i < r; i++) S[r + L + i] = c->mt_base[cpos++];
    int carry = 0;
    for (int i = 0; i < k - 1; i++) carry = 4 * carry + S[i];
    for (int i = 0; i < L + 2 * r - k + 1; i++) {
        S[i] = 4 * carry + S[i + k - 1];
        carry = S[i] % c->pow4[k - 1];
    }
}

void mtro_load_read(mtro_ctx *c, const int *bases, int len)
{
    for (int i = 0; i < len; i++) c->org[i] = bases[i];     /* handle_one_file.c:284-285 */
    c->cur_len = len;
}

void mtro_padded_codes(mtro_ctx *c, int k, int *out, int n)
{
    build_padded(c, k);
    for (int i = 0; i < n; i++) out[i] = c->padded[i];
}

/* One (k, w) pass: fill_directional_index_Manhattan :171-295 / _PCC :298-450.  The reference slides three
 * adjacent w-wide histograms and maintains the distance terms incrementally; restated here with one helper
 * that moves a single count and keeps every sum exact (all quantities are integers). */
typedef struct { long long d01, d12, q0, q1, q2, ip01, ip12; } di_sums;

static inline long long iabs64(long long v) { return v < 0 ? -v : v; }

static void bump(mtro_ctx *c, di_sums *s, int which, int bin, int delta)
{
    int *h0 = c->h0, *h1 = c->h1, *h2 = c->h2;
    s->d01 -= iabs64(h0[bin] - h1[bin]);  s->d12 -= iabs64(h1[bin] - h2[bin]);
    s->ip01 -= (long long)h0[bin] * h1[bin];  s->ip12 -= (long long)h1[bin] * h2[bin];
    s->q0 -= (long long)h0[bin] * h0[bin]; s->q1 -= (long long)h1[bin] * h1[bin]; s->q2 -= (long long)h2[bin] * h2[bin];
    if (which == 0) h0[bin] += delta; else if (which == 1) h1[bin] += delta; else h2[bin] += delta;
    s->d01 += iabs64(h0[bin] - h1[bin]);  s->d12 += iabs64(h1[bin] - h2[bin]);
    s->ip01 += (long long)h0[bin] * h1[bin];  s->ip12 += (long long)h1[bin] * h2[bin];
    s->q0 += (long long)h0[bin] * h0[bin]; s->q1 += (long long)h1[bin] * h1[bin]; s->q2 += (long long)h2[bin] * h2[bin];
}

static void di_pass(mtro_ctx *c, int k, int w)
{
    const int L = c->cur_len, r = random_len_of(L), N = L + 2 * r;
    const int *S = c->padded;
    double *tmp = c->di_tmp;
    const int nbins = c->pow4[k];
    for (int i = 0; i < N; i++) tmp[i] = -1;
    memset(c->h0, 0, HISTBINS * sizeof(int));
    memset(c->h1, 0, HISTBINS * sizeof(int));
    memset(c->h2, 0, HISTBINS * sizeof(int));
    for (int i = 0; i < w; i++) { c->h0[S[i]]++; c->h1[S[i + w]]++; c->h2[S[i + 2 * w]]++; }
    di_sums s; memset(&s, 0, sizeof s);
    for (int b = 0; b < nbins; b++) {                       /* :191-198 / :323-332: only bins < 4^k */
        s.d01 += iabs64(c->h0[b] - c->h1[b]);  s.d12 += iabs64(c->h1[b] - c->h2[b]);
        s.q0 += (long long)c->h0[b] * c->h0[b]; s.q1 += (long long)c->h1[b] * c->h1[b]; s.q2 += (long long)c->h2[b] * c->h2[b];
        s.ip01 += (long long)c->h0[b] * c->h1[b]; s.ip12 += (long long)c->h1[b] * c->h2[b];
    }
    const double n = (double)nbins, sw = (double)w;
    const int steps = N - w - r - k + 1;
    for (int i = 0; i < steps; i++) {
        double DI;
        if (c->manhattan) {
            DI = ((double)s.d01 - (double)s.d12) / (2 * sw);                    /* :211 */
        } else {
            double sd0 = sqrt((double)s.q0 * n - sw * sw);                      /* :340-357 */
            double sd1 = sqrt((double)s.q1 * n - sw * sw);
            double sd2 = sqrt((double)s.q2 * n - sw * sw);
            double P01 = 0, P12 = 0;
            if (sd0 * sd1 > 0) P01 = ((double)s.ip01 * n - sw * sw) / (sd0 * sd1);
            if (sd1 * sd2 > 0) P12 = ((double)s.ip12 * n - sw * sw) / (sd1 * sd2);
            DI = P12 - P01;
        }
        tmp[i + w] = DI;
        const int a = S[i], b = S[i + w], cc = S[i + 2 * w], d = S[i + 3 * w];
        bump(c, &s, 0, a, -1); bump(c, &s, 0, b, +1);
        bump(c, &s, 1, b, -1); bump(c, &s, 1, cc, +1);
        bump(c, &s, 2, cc, -1); bump(c, &s, 2, d, +1);
    }
    c->st.di_position_passes += steps;
}

void mtro_di_pass(mtro_ctx *c, int k, int w, double *tmp)
{
    const int L = c->cur_len, N = L + 2 * random_len_of(L);
    build_padded(c, k);
    di_pass(c, k, w);
    memcpy(tmp, c->di_tmp, (size_t)N * sizeof(double));
}

/* put_local_maximum_into_directional_index, fill_directional_index.c:467-503 */
static void merge_pass(mtro_ctx *c, int N, int w)
{
    const double *tmp = c->di_tmp;
    double local_max = -1;
    int local_max_i = -1;
    for (int i = 0; i < N; i++) {
        if (local_max < tmp[i]) { local_max = tmp[i]; local_max_i = i; }
        /* local_max_i == -1 implies local_max == -1, so the reference's read of DI[-1] (H7) is inert */
        if (local_max_i >= 0 && local_max_i + w < i && c->di[local_max_i] < local_max && 0 < local_max) {
            double local_min = 1;
            int local_min_j = local_max_i;
            for (int j = local_max_i; j < N; j++) {
                if (local_min > tmp[j]) { local_min = tmp[j]; local_min_j = j; }
                if (local_min_j + w < j) {
                    c->di[local_max_i] = local_max;
                    c->di_w[local_max_i] = w;
                    c->di_end[local_max_i] = local_min_j + w;
                    i = local_min_j + w;        /* may move backwards; the for's i++ follows */
                    break;
                }
            }
            local_max = -1;                     /* local_max_i keeps its value (Q2) */
        }
    }
}

/* remove_redundant_ranges, fill_directional_index.c:505-546 */
static void prune_ranges(mtro_ctx *c, int L)
{
    double *di = c->di; int *en = c->di_end;
    for (int i = 0; i < L; i++) {
        const int ie = en[i];
        const double idi = di[i];
        if (!(0 < idi)) continue;
        for (int j = i + 1; j <= ie; j++) {
            const int je = en[j];
            const double jdi = di[j];
            if (!(0 < jdi)) continue;
            int lo_end = ie < je ? ie : je, hi_end = ie > je ? ie : je;
            double jac = (double)(lo_end - j) / (hi_end - i);        /* max(i,j) = j, min(i,j) = i */
            if (0.98 < jac) {
                if (idi < jdi) { di[i] = -1; en[i] = -1; break; }
                di[j] = -1; en[j] = -1;
            } else {
                if (i >= j && ie <= je && idi < jdi) { di[i] = -1; en[i] = -1; break; }   /* never: i < j */
                if (i <= j && ie >= je && idi > jdi) { di[j] = -1; en[j] = -1; }
            }
        }
    }
}

/* fill_directional_index_with_end, fill_directional_index.c:549-602 */
static void directional_index(mtro_ctx *c)
{
    const int L = c->cur_len, r = random_len_of(L), N = L + 2 * r;
    for (int i = 0; i < N; i++) { c->di[i] = -1; c->di_end[i] = -1; c->di_w[i] = -1; }
    for (int k = 1; k <= 5; k += 2) {
        const int max_w = k == 1 ? 20 : (k == 3 ? 80 : 10240);
        build_padded(c, k);
        for (int w = 5; w <= max_w && w < L / 2; w *= 2) {
            di_pass(c, k, w);
            merge_pass(c, N, w);
        }
    }
    for (int i = 0; i < L; i++) {
        c->di[i] = c->di[i + r];
        c->di_end[i] = c->di_end[i + r] - r;
        c->di_w[i] = c->di_w[i + r];
    }
    for (int i = L; i < N; i++) { c->di[i] = -1; c->di_end[i] = -1; c->di_w[i] = -1; }
    prune_ranges(c, L);
}

void mtro_directional_index(mtro_ctx *c, double *di, int *end, int *w)
{
    directional_index(c);
    for (int i = 0; i < c->cur_len; i++) { di[i] = c->di[i]; end[i] = c->di_end[i]; w[i] = c->di_w[i]; }
}

/* ================================================================ wrap-around DP */

static void ensure_W(mtro_ctx *c, size_t need)
{
    if (need <= c->W_cap) return;
    size_t cap = c->W_cap ? c->W_cap : (1u << 20);
    while (cap < need) cap *= 2;
    int *nw = realloc(c->W, cap * sizeof(int));
    if (!nw) { fprintf(stderr, "mtr_oracle: out of memory (DP matrix)\n"); exit(EXIT_FAILURE); }
    memset(nw + c->W_cap, 0, (cap - c->W_cap) * sizeof(int));   /* untouched WrapDP pages read as zero */
    c->W = nw; c->W_cap = cap;
}

/* wrap_around_DP_sub fill + traceback (wrap_around_DP.c:236-333); the same nest is in
 * pretty_print_alignment (:68-186) and revise_representative_unit_sub (consensus.c:875-962). */
void mtro_wrap_dp(mtro_ctx *c, const int *x, int rows, const int *u, int ulen,
                  int G, int MM, int IN, int mode, mtro_dp_result *res,
                  int *consensus, int *missing, unsigned char *path, unsigned char *dirs)
{
    const int next = ulen + 1;
    size_t need = (size_t)next * (size_t)(rows + 1) + 1;
    if (need < (size_t)rows + 2) need = (size_t)rows + 2;
    ensure_W(c, need);
    int *W = c->W;
    for (int j = 0; j <= rows; j++) W[j] = 0;               /* :250 -- linear, not per column */
    int best = 0, max_i = 0, max_j = 0;
    for (int i = 1; i <= rows; i++) {
        int *cur = W + (size_t)next * i, *prev = cur - next;
        for (int j = 1; j <= ulen; j++) {
            int v;
            if (x[i] == u[j]) {
                v = prev[j - 1] + G;
            } else {
                int vm = prev[j - 1] - MM, vi = prev[j] - IN;
                v = vm > vi ? vm : vi;
                if (j > 1) { int vd = cur[j - 1] - IN; if (vd > v) v = vd; }
                if (v < 0) v = 0;
            }
            cur[j] = v;
            if (best < v) { best = v; max_i = i; max_j = j; }
        }
        cur[0] = cur[ulen];                                 /* wrap around, :284 */
    }
    if (dirs) {
        memset(dirs, 3, (size_t)(rows + 1) * next);
        for (int i = 1; i <= rows; i++) {
            const int *cur = W + (size_t)next * i, *prev = cur - next;
            for (int j = 1; j <= ulen; j++) {
                int v = cur[j], d;
                if (v <= 0) d = 3;
                else if (x[i] == u[j]) d = 0;
                else if (v == prev[j - 1] - MM) d = 0;
                else if (v == cur[j - 1] - IN) d = 1;       /* j == 1 sees the wrap copy (Q6) */
                else d = 2;
                dirs[(size_t)i * next + j] = (unsigned char)d;
            }
        }
    }
    if (mode == MTRO_TB_CONSENSUS) {
        memset(consensus, 0, (size_t)(ulen + 1) * 5 * sizeof(int));
        memset(missing, 0, (size_t)(ulen + 1) * 4 * sizeof(int));
    }
    int nm = 0, nx = 0, ni = 0, nd = 0, scanned = 0, steps = 0;
    int i = max_i, j = max_j, run = best;
    if (j == 0) j = ulen;
    while (i > 0 && W[(size_t)next * i + j] > 0) {
        const int *cur = W + (size_t)next * i, *prev = cur - next;
        int op;
        if (run == prev[j - 1] + G && x[i] == u[j]) op = 0;
        else if (run == prev[j - 1] - MM && x[i] != u[j]) op = 1;
        else if (run == cur[j - 1] - IN) op = 2;
        else if (run == prev[j] - IN) op = 3;
        else if (run == 0) break;
        else { fprintf(stderr, "fatal error in wrap-around DP max_wrd = %i\n", run); exit(EXIT_FAILURE); }
        if (mode == MTRO_TB_PATH) path[steps] = (unsigned char)op;
        steps++;
        switch (op) {
        case 0: if (mode == MTRO_TB_CONSENSUS) consensus[j * 5 + x[i]]++; run -= G;  i--; j--; nm++; scanned++; break;
        case 1: if (mode == MTRO_TB_CONSENSUS) consensus[j * 5 + x[i]]++; run += MM; i--; j--; nx++; scanned++; break;
        case 2: if (mode == MTRO_TB_CONSENSUS) consensus[j * 5 + 4]++;    run += IN; j--;      nd++; scanned++; break;
        default: if (mode == MTRO_TB_CONSENSUS) missing[j * 4 + x[i]]++;  run += IN; i--;      ni++; break;
        }
        if (j == 0) j = ulen;
    }
    res->best = best; res->max_i = max_i; res->max_j = max_j; res->end_i = i; res->end_j = j;
    res->n_match = nm; res->n_mismatch = nx; res->n_ins = ni; res->n_del = nd; res->n_scanned = scanned;
    res->cells = (long long)rows * ulen; res->path_len = steps;
}

static void unit_to_ints(const mtro_rr *rr, int *u)       /* 1-origin, wrap_around_DP.c:231-242 */
{
    for (int i = 0; i < rr->rep_period; i++) u[i + 1] = base_of_char(rr->unit[i]);
}

/* wrap_around_DP_sub, wrap_around_DP.c:222-354 */
static void dp_sub(mtro_ctx *c, int qs, int qe, mtro_rr *rr, int G, int MM, int IN)
{
    int u[MTRO_UNIT_CAP + 1];
    const int ulen = rr->rep_period;
    unit_to_ints(rr, u);
    if (ulen <= 0) { fprintf(stderr, "mtr_oracle: unit of length %d reached the DP (H9)\n", ulen); exit(EXIT_FAILURE); }
    const int rows = qe - qs + 1;
    mtro_dp_result res;
    mtro_wrap_dp(c, c->org + qs, rows, u, ulen, G, MM, IN, MTRO_TB_COUNTS, &res, NULL, NULL, NULL, NULL);
    c->st.dp_calls++; c->st.dp_cells += res.cells;
    if (c->hook) c->hook(c->hook_user, 0, c->org + qs, rows, u, ulen, G, MM, IN, &res);
    rr->rep_start = qs + res.end_i + 1;                     /* :337-350 */
    rr->rep_end = qs + res.max_i;
    rr->repeat_len = res.max_i - res.end_i;
    rr->n_units = res.n_scanned / ulen;
    rr->n_match = res.n_match; rr->n_mismatch = res.n_mismatch; rr->n_ins = res.n_ins; rr->n_del = res.n_del;
    rr->gain = G; rr->mis_pen = MM; rr->indel_pen = IN;
}

/* wrap_around_DP, wrap_around_DP.c:357-429 */
static void dp_both(mtro_ctx *c, int qs, int qe, mtro_rr *rr)
{
    static const int params[2][3] = { {1, 1, 3}, {1, 3, 1} };
    mtro_rr tmp, best;
    rr_clear(&best);
    float best_ratio = -1;
    for (int p = 0; p < 2; p++) {
        tmp = *rr;
        dp_sub(c, qs, qe, &tmp, params[p][0], params[p][1], params[p][2]);
        float ratio = rr_ratio(&tmp);
        if (best_ratio < ratio) { best = tmp; best_ratio = ratio; }
    }
    *rr = best;
}

/* pretty_print_alignment, wrap_around_DP.c:57-213: window is org[rep_start .. rep_end] */
static void print_alignment(void *user, const mtro_rr *rr)
{
    mtro_ctx *c = user;
    int u[MTRO_UNIT_CAP + 1];
    unit_to_ints(rr, u);
    const int rows = rr->rep_end - rr->rep_start + 1;
    const int *x = c->org + rr->rep_start - 1;
    unsigned char *path = malloc((size_t)rows + (size_t)rows * rr->rep_period + 16);
    mtro_dp_result res;
    mtro_wrap_dp(c, x, rows, u, rr->rep_period, rr->gain, rr->mis_pen, rr->indel_pen, MTRO_TB_PATH,
                 &res, NULL, NULL, path, NULL);
    c->st.print_calls++; c->st.print_cells += res.cells;
    if (c->hook) c->hook(c->hook_user, 2, x, rows, u, rr->rep_period, rr->gain, rr->mis_pen, rr->indel_pen, &res);
    const int n = res.path_len;
    char *a = malloc(n + 1), *m = malloc(n + 1), *b = malloc(n + 1);
    int i = res.max_i, j = res.max_j;
    if (j == 0) j = rr->rep_period;
    static const char ch[4] = { 'A', 'C', 'G', 'T' };
    for (int t = 0; t < n; t++) {
        switch (path[t]) {
        case 0: a[t] = ch[x[i]]; m[t] = '|'; b[t] = ch[u[j]]; i--; j--; break;
        case 1: a[t] = ch[x[i]]; m[t] = ' '; b[t] = ch[u[j]]; i--; j--; break;
        case 2: a[t] = '-';      m[t] = ' '; b[t] = ch[u[j]]; j--; break;
        default: a[t] = ch[x[i]]; m[t] = ' '; b[t] = '-'; i--; break;
        }
        if (j == 0) j = rr->rep_period;
    }
    fprintf(c->out, "match gain = %i, mismatch penalty = %i, indel penalty = %i\n\n", rr->gain, rr->mis_pen, rr->indel_pen);
    for (int s = n - 1; s >= 0; s -= 50) {                  /* :188-212 */
        int e = s - 50 >= -1 ? s - 50 : -1;
        for (int t = s; t > e; t--) fputc(a[t], c->out);
        fputc('\n', c->out);
        for (int t = s; t > e; t--) fputc(m[t], c->out);
        fputc('\n', c->out);
        for (int t = s; t > e; t--) fputc(b[t], c->out);
        fputs("\n\n", c->out);
    }
    free(a); free(m); free(b); free(path);
}

/* ================================================================ k-mer counts (consensus.c:37-253) */

static void kmer_window(mtro_ctx *c, int k, int qs, int qe)          /* init_inputString :37-60 */
{
    const int L = c->cur_len;
    for (int i = qs; i < qe + k - 1 && i < L; i++) c->kstr[i] = c->org[i];
    int carry = 0;
    for (int i = qs; i < qs + k - 1; i++) carry = 4 * carry + c->kstr[i];
    for (int i = qs; i < qe && i < L - k + 1; i++) {
        c->kstr[i] = 4 * carry + c->kstr[i + k - 1];
        carry = c->kstr[i] % c->pow4[k - 1];
    }
}

static void hash_reset(mtro_ctx *c, int width)
{
    int cap = 1024;
    while (cap < 4 * (width + 2)) cap *= 2;
    if (cap > c->hcap) {
        free(c->hkey); free(c->hval);
        c->hkey = xcalloc(cap, sizeof(int)); c->hval = xcalloc(cap, sizeof(int));
        c->hcap = cap;
    }
    memset(c->hkey, 0xff, (size_t)c->hcap * sizeof(int));
    memset(c->hval, 0, (size_t)c->hcap * sizeof(int));
}

static inline int hash_slot(const mtro_ctx *c, int node)
{
    unsigned h = ((unsigned)node * 2654435761u) & (unsigned)(c->hcap - 1);
    while (c->hkey[h] != -1 && c->hkey[h] != node) h = (h + 1) & (unsigned)(c->hcap - 1);
    return (int)h;
}

static inline int *count_ref(mtro_ctx *c, int node, int create)
{
    if (c->count_k <= 6) {
        /* k <= 6 uses a direct table (consensus.c:142-163).  A stale code >= 4^k can only come from
         * kstr[L] (H4b); the reference then indexes out of bounds, here it is ignored. */
        if (node < 0 || node >= c->pow4[c->count_k]) return NULL;
        return &c->count[node];
    }
    int h = hash_slot(c, node);
    if (c->hkey[h] == -1) { if (!create) return NULL; c->hkey[h] = node; }
    return &c->hval[h];
}

static inline int node_count(mtro_ctx *c, int node)                  /* freq_node :231-253 */
{
    int *p = count_ref(c, node, 0);
    return p ? *p : 0;
}

static int count_window(mtro_ctx *c, int k, int qs, int qe)          /* returns maxFreq */
{
    c->count_k = k;
    if (k <= 6) memset(c->count, 0, (size_t)c->pow4[k] * sizeof(int));
    else hash_reset(c, qe - qs + 1);
    for (int i = qs; i <= qe; i++) { int *p = count_ref(c, c->kstr[i], 1); if (p) (*p)++; }
    int maxf = -1;
    for (int i = qs; i <= qe; i++) { int v = node_count(c, c->kstr[i]); if (maxf < v) maxf = v; }
    return maxf;
}

/* generate_freqNode_return_list_maxNodes :132-229 (listing decrements the listed node's count, Q8) */
static int list_max_nodes(mtro_ctx *c, int k, int qs, int qe, int *list, int cap, int *maxfreq)
{
    int maxf = count_window(c, k, qs, qe), n = 0;
    for (int i = qs; i <= qe; i++) {
        int *p = count_ref(c, c->kstr[i], 0);
        if (p && *p == maxf) {
            list[n++] = c->kstr[i];
            (*p)--;
            if (cap <= n) break;
        }
    }
    *maxfreq = maxf;
    return n;
}

/* generate_freqNode_return_maxNode :71-130 (first node, in window order, to reach the final maximum) */
static int first_max_node(mtro_ctx *c, int k, int qs, int qe)
{
    int node = 0;
    if (k <= 6) {
        int maxf = count_window(c, k, qs, qe);
        for (int i = qs; i <= qe; i++) if (node_count(c, c->kstr[i]) == maxf) { node = c->kstr[i]; break; }
    } else {
        /* hash branch :105-123: running maximum while counting */
        c->count_k = k;
        hash_reset(c, qe - qs + 1);
        int maxf = -1;
        for (int i = qs; i <= qe; i++) {
            int *p = count_ref(c, c->kstr[i], 1);
            (*p)++;
            if (maxf < *p) { maxf = *p; node = c->kstr[i]; }
        }
    }
    return node;
}

/* ================================================================ de Bruijn walks (consensus.c:269-505) */

static int walk(mtro_ctx *c, int qs, int qe, int start, int k, int backward, mtro_rr *rr)
{
    const int *p4 = c->pow4;
    int ustr[MAXP], uscore[MAXP];
    int ties[MAXTIES], ties_new[MAXTIES];
    int node = start, period = 0;
    const int limit = (qe - qs) / 5;
    for (int l = 0; l < MAXP && l < limit; l++) {
        if (!backward) { ustr[l] = node / p4[k - 1]; uscore[l] = node_count(c, node); }
        int m, best_digit = 0, nties = 1;
        ties[0] = 0;
        const int lookahead = l < 10 ? 1 : k;
        for (m = 1; m <= lookahead; m++) {
            int best = -1, nnew = 0;
            best_digit = 0;
            for (int t = 0; t < nties; t++) {
                for (int jb = 0; jb < 4; jb++) {
                    int digit, cand;
                    if (!backward) {
                        digit = 4 * ties[t] + jb;
                        cand = p4[m] * (node % p4[k - m]) + digit;
                    } else {
                        digit = jb * p4[m - 1] + ties[t];
                        cand = digit * p4[k - m] + node / p4[m];
                    }
                    int cnt = node_count(c, cand);
                    if (best < cnt) { best = cnt; best_digit = digit; nnew = 0; ties_new[nnew++] = digit; }
                    else if (best == cnt && nnew < MAXTIES) ties_new[nnew++] = digit;
                }
            }
            if (backward ? nnew <= 1 : nnew == 1) break;
            memcpy(ties, ties_new, (size_t)nnew * sizeof(int));
            nties = nnew;
        }
        if (!backward) {
            node = 4 * (node % p4[k - 1]) + best_digit / p4[m - 1];     /* m == lookahead+1 -> 'A' (Q9) */
        } else {
            node = (best_digit % 4) * p4[k - 1] + node / 4;
            ustr[l] = node / p4[k - 1]; uscore[l] = node_count(c, node);
        }
        if (node == start) { period = l + 1; if (MAXP <= period) period = 0; break; }
    }
    if (period == 0) return 0;
    rr->rep_period = period;
    for (int i = 0; i < period; i++) {
        int src = backward ? period - 1 - i : i;
        rr->unit[i] = "ACGT"[ustr[src]];
        rr->unit_score[i] = uscore[src];
    }
    rr->unit[period] = '\0';
    return 1;
}

/* search_De_Bruijn_graph, consensus.c:507-582 */
static int search_unit(mtro_ctx *c, int qs, int qe, mtro_rr *rr)
{
    const int k = rr->kmer;
    kmer_window(c, k, qs, qe);
    int nodes[100], maxfreq;
    int nn = list_max_nodes(c, k, qs, qe, nodes, 100, &maxfreq);
    c->st.searches++;
    mtro_rr tmp, best;
    rr_clear(&best);
    float best_ratio = -1;
    int found = 0;
    if (5 < maxfreq) {
        for (int dir = 0; dir < 2; dir++) {
            for (int i = 0; i < nn; i++) {
                tmp = *rr;
                found = walk(c, qs, qe, nodes[i], k, dir, &tmp);
                if (!found) continue;
                dp_both(c, qs, qe, &tmp);
                float ratio = rr_ratio(&tmp);
                if (best_ratio < ratio && c->min_match_ratio <= ratio && 5 < tmp.n_units &&
                    2 <= tmp.rep_period && tmp.rep_period < MAXP) {
                    best_ratio = ratio; best = tmp;
                }
                break;
            }
        }
    }
    *rr = best;
    return found;                                   /* flag of the LAST attempt (Q4) */
}

void mtro_unit_walks(mtro_ctx *c, int qs, int qe, int k, mtro_walk_result *res,
                     unsigned char *unit_fwd, int *score_fwd, unsigned char *unit_bwd, int *score_bwd)
{
    kmer_window(c, k, qs, qe);
    int nodes[100], maxfreq;
    int nn = list_max_nodes(c, k, qs, qe, nodes, 100, &maxfreq);
    memset(res, 0, sizeof *res);
    res->max_freq = maxfreq;
    if (!(5 < maxfreq)) return;
    mtro_rr tmp;
    for (int dir = 0; dir < 2; dir++) {
        for (int i = 0; i < nn; i++) {
            rr_clear(&tmp);
            int found = walk(c, qs, qe, nodes[i], k, dir, &tmp);
            res->found_last = found;
            if (!found) continue;
            res->found[dir] = 1; res->period[dir] = tmp.rep_period;
            unsigned char *u = dir ? unit_bwd : unit_fwd;
            int *sc = dir ? score_bwd : score_fwd;
            for (int j = 0; j < tmp.rep_period; j++) { u[j] = (unsigned char)base_of_char(tmp.unit[j]); sc[j] = tmp.unit_score[j]; }
            break;
        }
    }
}

/* ================================================================ polish + revise (consensus.c:584-1087) */

static int align_score(mtro_ctx *c, int start, int k, int node, int period, const int *unit)  /* :584-596 */
{
    int sum = 0;
    for (int j = start; 0 <= j && start - k < j; j--) {
        node = unit[j % period] * c->pow4[k - 1] + node / 4;
        sum += node_count(c, node);
    }
    return sum;
}

static int suspicious(const mtro_rr *rr, int j)                                                /* :598-608 */
{
    int cnt = 0;
    for (int i = 0; i < rr->kmer - 1 && 0 <= j - i; i++) if (rr->unit_score[j - i] < 2) cnt++;
    return (rr->kmer - 1) * 0.8 < (double)cnt;
}

static void polish_unit(mtro_ctx *c, mtro_rr *rr)                                               /* :610-704 */
{
    const int k = rr->kmer, period = rr->rep_period;
    const int *p4 = c->pow4;
    if (period <= k) return;
    kmer_window(c, k, rr->rep_start, rr->rep_end);
    (void)first_max_node(c, k, rr->rep_start, rr->rep_end);         /* builds the count table */
    int unit[MAXP], revised[MAXP];
    for (int i = 0; i < period; i++) unit[i] = base_of_char(rr->unit[i]);
    int jr = MAXP - 1;
    int best = 0;
    for (int i = 0; i < k; i++) best = unit[i] * p4[k - 1 - i] + best;
    for (int j = period - 1; 0 <= j; ) {
        int ref = unit[j] * p4[k - 1] + best / 4;
        int best_freq = node_count(c, ref);
        best = ref;
        if (rr->unit_score[j] == 1 && suspicious(rr, j)) {
            for (int l = 0; l < 4; l++) {
                int alt = (ref + (l - unit[j]) * p4[k - 1]) % p4[k];
                if (best_freq < node_count(c, alt)) { best_freq = node_count(c, alt); best = alt; }
            }
            if (best == ref) {
                revised[jr--] = unit[j--];
            } else {
                int s_del = align_score(c, j, k, best, period, unit);
                int s_sub = align_score(c, j - 1, k, best, period, unit);
                int s_ins = -1;
                /* j == 0 reads unit[-1] in the reference; the outcome cannot depend on it (s_del >= 0) */
                if (j >= 1 && best / p4[k - 1] == unit[(j - 1) % period])
                    s_ins = align_score(c, j - 2, k, best, period, unit);
                revised[jr--] = best / p4[k - 1];
                int mx = s_del > s_sub ? s_del : s_sub; if (s_ins > mx) mx = s_ins;
                if (mx == s_del) { /* keep j */ } else if (mx == s_sub) j -= 1; else j -= 2;
            }
        } else {
            revised[jr--] = unit[j--];
        }
        if (jr < 0) return;                                         /* fails to revise */
    }
    rr->rep_period = (MAXP - 1) - jr;
    for (int i = 0; i < rr->rep_period; i++) rr->unit[i] = "ACGT"[revised[i + jr + 1]];
    rr->unit[rr->rep_period] = '\0';
}

/* min_missing_bases[10][10][20] (consensus.c:714-785): every row starts at 1 and never steps by more
 * than 1, so it is stored as the 19 step bits of each row. */
static const unsigned min_missing_steps[10][10] = {
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

int mtro_min_missing_raw(int i, int j, int k)
{
    unsigned m = min_missing_steps[i][j] & ((1u << k) - 1u);
    int v = 1;
    while (m) { v += (int)(m & 1u); m >>= 1; }
    return v;
}

int mtro_min_missing(int rep_period, double error, int coverage)    /* consensus.c:787-820 */
{
    static const int plim[9] = { 200, 150, 100, 75, 50, 30, 20, 10, 5 };
    static const double elim[9] = { 0.25, 0.225, 0.2, 0.175, 0.15, 0.125, 0.1, 0.075, 0.05 };
    const float r = 1;
    int i = 9, j = 9, k;
    for (int t = 0; t < 9; t++) if (rep_period > r * plim[t]) { i = t; break; }
    for (int t = 0; t < 9; t++) if (error > r * elim[t]) { j = t; break; }
    k = coverage <= 1 ? 0 : (coverage >= 20 ? 19 : coverage - 1);
    return mtro_min_missing_raw(i, j, k);
}

/* revise_representative_unit_sub, consensus.c:851-1046 */
static void revise_sub(mtro_ctx *c, mtro_rr *rr, int G, int MM, int IN)
{
    const int ulen = rr->rep_period, qs = rr->rep_start, qe = rr->rep_end;
    rr->gain = G; rr->mis_pen = MM; rr->indel_pen = IN;
    int u[MTRO_UNIT_CAP + 1];
    unit_to_ints(rr, u);
    const int rows = qe - qs + 1;
    int *consensus = xcalloc((size_t)(ulen + 1) * 5, sizeof(int));
    int *missing = xcalloc((size_t)(ulen + 1) * 4, sizeof(int));
    mtro_dp_result res;
    mtro_wrap_dp(c, c->org + qs, rows, u, ulen, G, MM, IN, MTRO_TB_CONSENSUS, &res, consensus, missing, NULL, NULL);
    c->st.revise_calls++; c->st.revise_cells += res.cells;
    if (c->hook) c->hook(c->hook_user, 1, c->org + qs, rows, u, ulen, G, MM, IN, &res);
    int revised[2 * MAXP + 8], n = 0;
    for (int j = 1; j <= ulen; j++) {
        int mv = -1, mb = -1;
        for (int q = 0; q < 5; q++) if (mv < consensus[j * 5 + q]) { mv = consensus[j * 5 + q]; mb = q; }
        if (mb < 4) revised[n++] = mb;
        mv = -1; int mm = -1;
        for (int q = 0; q < 4; q++) if (mv < missing[j * 4 + q]) { mv = missing[j * 4 + q]; mm = q; }
        int coverage = rr->repeat_len / rr->rep_period;
        if (5 <= coverage && coverage <= 20) {
            double mismatch_ratio = (double)(rr->n_mismatch + rr->n_ins + rr->n_del) / rr->repeat_len;
            if (mtro_min_missing(rr->rep_period, mismatch_ratio, coverage) <= mv && 0 <= mm && mm <= 3)
                revised[n++] = mm;
        }
    }
    rr->rep_period = n;
    for (int i = 0; i < n; i++) rr->unit[i] = "ACGT"[revised[i]];
    rr->unit[n] = '\0';
    free(consensus); free(missing);
}

/* revise_representative_unit, consensus.c:1048-1087 */
static void revise_unit(mtro_ctx *c, mtro_rr *rr)
{
    static const int params[2][3] = { {5, 1, 1}, {1, 1, 3} };
    polish_unit(c, rr);
    const float ratio0 = rr_ratio(rr);              /* never refreshed (Q10) */
    mtro_rr tmp;
    for (int p = 0; p < 2; p++) {
        tmp = *rr;
        revise_sub(c, &tmp, params[p][0], params[p][1], params[p][2]);
        if (tmp.rep_period < MAXP) {
            dp_sub(c, tmp.rep_start, tmp.rep_end, &tmp, params[p][0], params[p][1], params[p][2]);
            if (ratio0 < rr_ratio(&tmp)) *rr = tmp;
        }
    }
}

/* find_tandem_repeat_sub, handle_one_read.c:77-100 */
static void find_sub(mtro_ctx *c, int qs, int qe, mtro_rr *rr)
{
    int found = search_unit(c, qs, qe, rr);
    if (!found) { rr_clear(rr); return; }
    if ((long long)rr->rep_period * (qe - qs + 1) > WRAPCAP) {
        fprintf(stderr, "You need to increse the value of WrapDPsize.\n");
        rr_clear(rr);
        return;
    }
    int coverage = rr->repeat_len / rr->rep_period;
    if (5 <= coverage && coverage <= 20 && 5 < rr->rep_period) revise_unit(c, rr);
}

/* find_tandem_repeat, handle_one_read.c:102-154 */
void mtro_find_tandem_repeat(mtro_ctx *c, int qs, int qe, int w, mtro_rr *out)
{
    int min_k, max_k;
    if (w < 100) { min_k = 2; max_k = 10; }
    else if (w < 1000) { min_k = 2; max_k = 12; }
    else { min_k = 5; max_k = 15; }
    float best_ratio = -1;
    mtro_rr tmp;
    for (int k = min_k; k <= max_k; k++) {
        rr_clear(&tmp);
        tmp.inputLen = c->cur_len;
        tmp.kmer = k;
        find_sub(c, qs, qe, &tmp);
        float ratio = rr_ratio(&tmp);
        if (best_ratio < ratio && c->min_match_ratio <= ratio && 5 < tmp.n_units && 2 <= tmp.rep_period) {
            best_ratio = ratio;
            *out = tmp;
        }
    }
}

/* handle_one_TR, handle_one_read.c:190-261 */
void mtro_process_read(mtro_ctx *c, const char *read_id, const int *bases, int len, int print_aln)
{
    mtro_load_read(c, bases, len);
    c->cur_id = read_id;
    c->st.reads++; c->st.bases += len;
    directional_index(c);
    mtro_rr rr;
    for (int qs = 0; qs < len; qs++) {
        int qe = c->di_end[qs];
        if (!(-1 < qe && qe < len)) continue;
        rr_clear(&rr);
        mtro_find_tandem_repeat(c, qs, qe, c->di_w[qs], &rr);
        c->st.candidates++;
        if (rr.repeat_len > 0 && rr.rep_start + 10 < rr.rep_end) {
            mtro_chain_insert(c->chain, read_id, &rr);
            for (int i = rr.rep_start; i < rr.rep_end; i++)         /* :178-188 */
                if (c->di[i] != -1 && c->di_end[i] < rr.rep_end) { c->di[i] = -1; c->di_end[i] = -1; c->di_w[i] = -1; }
        }
    }
    mtro_chain_run(c->chain, c->out, print_aln, print_alignment, c);
}

/* handle_one_file + return_one_read, handle_one_file.c:201-293 */
int mtro_process_file(mtro_ctx *c, const char *path, int print_aln)
{
    FILE *fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, "fatal error: cannot open %s\n", path); exit(EXIT_FAILURE); }
    char *line = malloc(4096);
    int *bases = malloc(sizeof(int) * MAXLEN);
    char id[4096] = "", next_id[4096] = "";
    int n = 0, nreads = 0, seen_header = 0, any = 0, stop = 0;
    /* The reference consumes 4095-byte chunks and treats a chunk that starts with '>' as a header, whether
     * or not it is at the start of a line; restated as the same chunked loop. */
    while (!stop && fgets(line, 4096, fp)) {
        any = 1;
        if (line[0] == '>') {
            int i;
            if (seen_header) {
                strcpy(id, next_id);
                for (i = 1; line[i] && line[i] != '\n' && line[i] != '\r'; i++) next_id[i - 1] = line[i];
                next_id[i - 1] = '\0';
                if (n == 0) { stop = 1; break; }                    /* zero-length read ends the run (H8) */
                nreads++;
                mtro_process_read(c, id, bases, n, print_aln);
                n = 0;
            } else {
                seen_header = 1;
                for (i = 1; line[i] && line[i] != '\n' && line[i] != '\r'; i++) next_id[i - 1] = line[i];
                next_id[i - 1] = '\0';
            }
        } else {
            for (int i = 0; line[i] && line[i] != '\n' && line[i] != '\r'; i++) {
                int b;
                switch (line[i]) {
                case 'A': case 'a': b = 0; break;
                case 'C': case 'c': b = 1; break;
                case 'G': case 'g': b = 2; break;
                case 'T': case 't': b = 3; break;
                default: fprintf(stderr, "Invalid character: %c \n", line[i]); exit(EXIT_FAILURE);
                }
                bases[n++] = b;
                if (MAXLEN <= n) {
                    fprintf(stderr, "fatal error: The length %d is tentatively at most %i.\nread ID = %s\nSet MAX_INPUT_LENGTH to a larger value", n, MAXLEN, id);
                    fprintf(stderr, "cannot allocate space for one of global variables in the heap.\n");
                    exit(EXIT_FAILURE);
                }
            }
        }
    }
    if (!stop && any && n > 0) {
        strcpy(id, next_id);
        nreads++;
        mtro_process_read(c, id, bases, n, print_aln);
    }
    fclose(fp);
    free(line); free(bases);
    return nreads;
}

/* ================================================================ dead-code restatement (A7) */

void mtro_freq_2mer(const int *unit, int len, int *f)               /* handle_one_read.c:63-72 */
{
    for (int i = 0; i < 16; i++) f[i] = 0;
    for (int i = 1; i < len; i++) f[unit[i - 1] * 4 + unit[i]]++;
    f[unit[len - 1] * 4 + unit[0]]++;
}

int mtro_trs_in_neighborhood(const int *a, const int *rep, int rep_period)   /* k_means_clustering.c:169-180 */
{
    /* +1 if aTR lies in the neighbourhood of the representative, -1 otherwise: sum_i |a[i] - rep[i]| over the 16 2-mer
     * frequencies against MH_distance_threshold * repTR.actual_rep_period.  The reference never defines
     * MH_distance_threshold (the file is dead code and does not compile); 0.3 is the value SURVEY.md App. A records. */
    int diff = 0;
    for (int i = 0; i < 16; i++) diff += a[i] > rep[i] ? a[i] - rep[i] : rep[i] - a[i];
    return (0.3 * rep_period < diff) ? -1 : 1;
}

int mtro_cmp_tr(int period_a, const int *f2_a, int units_a, int rep_freq_a, int rep_id_a,
                int period_b, const int *f2_b, int units_b, int rep_freq_b, int rep_id_b, int mode)   /* k_means_clustering.c:62-101 */
{
    if (mode == 0 || mode == 1) {
        int diff = period_a - period_b;                        /* <actual_rep_period, freq_2mer [, Num_freq_unit]> */
        if (diff != 0) return diff;
        for (int i = 0; i < 16; i++) {
            diff = f2_a[i] - f2_b[i];
            if (diff != 0) return diff;
        }
        return mode == 1 ? units_a - units_b : 0;
    }
    if (mode == 2) {                                            /* <frequency descending, identifier> of the representatives */
        const int diff = -(rep_freq_a - rep_freq_b);
        return diff == 0 ? rep_id_a - rep_id_b : diff;
    }
    return 0;
}
