// tests/hostsim/sim_device.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the kernel-level layer of include/mtr_b200.h (mtr_cuda_init, mtr_reads_upload, mtr_di_run,
// mtr_wdp_run, ...) that answers every device call with the CPU oracle (oracle/mtr_oracle.c).  It is linked, together
// with the UNMODIFIED product source mtr_b200/csrc/pipeline.cpp, into tests/hostsim/_build/ only, so that the host
// side of the drop-in -- FASTA reader, stale-state tracker, group dispatcher, chaining, TSV / alignment formatting,
// the handle_one_file / handle_one_read entry points -- and, through sim_engine.cpp, the resident engine's own code
// (eng_core.h) can be checked against the golden digests on a machine without a GPU
// (`-m "not gpu"`), and profiled there.
//
// Nothing here is built by mtr_b200/csrc/Makefile, shipped in libmtr_b200.so or reachable from the product: the
// product has no CPU path (tests/test_abi_cpu.py checks that).  The stale state handed to mtr_di_run by the host's
// StaleTracker is what the simulated device sees (written into the oracle's persistent buffers read by read), so a
// wrong tracker shows up as a digest mismatch exactly as it would on the GPU.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <vector>
#include "../../mtr_b200/csrc/mtr_internal.h"
#include "../../oracle/mtr_oracle.h"

// ---- the few CUDA runtime entry points pipeline.cpp touches (pinned buffers, device selection)
extern "C" cudaError_t cudaSetDevice(int) { return cudaSuccess; }
extern "C" cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
extern "C" cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
extern "C" cudaError_t cudaMalloc(void **p, size_t) { *p = nullptr; return cudaErrorMemoryAllocation; }
extern "C" cudaError_t cudaFree(void *) { return cudaSuccess; }

struct SimReads {
    std::vector<uint32_t> packed;
    std::vector<int64_t> word_off;
    std::vector<int32_t> len;
    int base(int r, long long b) const { return (packed[(size_t)(word_off[r] + (b >> 4))] >> ((b & 15) * 2)) & 3; }
};

struct DiState {                       // hangs off mtr_ctx::di
    std::shared_ptr<SimReads> reads;
    mtro_ctx *oracle = nullptr;
    int oracle_manhattan = -1;
    mtro_ctx *get(int manhattan)
    {
        if (!oracle || oracle_manhattan != manhattan) { if (oracle) mtro_free(oracle); oracle = mtro_new(manhattan, 0.6f); oracle_manhattan = manhattan; }
        return oracle;
    }
};

static std::string g_init_error;
void mtr_set_error(mtr_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
}
long long wdp_dir_bytes(int ulen, int rows) { return (long long)rows * ((ulen + 15) / 16 * 4); }
void di_state_free(mtr_ctx *ctx) { if (ctx->di) { if (ctx->di->oracle) mtro_free(ctx->di->oracle); delete ctx->di; ctx->di = nullptr; } }
void uf_state_free(mtr_ctx *) {}
// the engine twin (sim_engine.cpp) reads the resident reads and shares the context's oracle
const uint32_t *sim_packed(mtr_ctx *ctx) { return ctx->di->reads->packed.data(); }
mtro_ctx *sim_oracle(mtr_ctx *ctx, int manhattan) { return ctx->di->get(manhattan ? 1 : 0); }

extern "C" {

const char *mtr_last_error(const mtr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }
int mtr_device_count(void) { const char *e = getenv("MTR_HOSTSIM_DEVICES"); return e ? atoi(e) : 2; }

int mtr_cuda_init(int device, mtr_ctx **out)
{
    if (!out) return MTR_EINVAL;
    *out = nullptr;
    if (device < 0 || device >= mtr_device_count()) { mtr_set_error(nullptr, "hostsim: device %d out of range", device); return MTR_EINVAL; }
    mtr_ctx *ctx = new mtr_ctx();
    ctx->device = device;
    ctx->n_sm = 148;
    ctx->stats.n_sm = 148;
    ctx->di = new DiState();
    *out = ctx;
    return MTR_OK;
}

void mtr_cuda_shutdown(mtr_ctx *ctx)
{
    if (!ctx) return;
    di_state_free(ctx);
    eng_state_free(ctx);
    delete ctx;
}

int mtr_set_priority(mtr_ctx *ctx, int) { return ctx ? MTR_OK : MTR_EINVAL; }
int mtr_set_blocking_sync(mtr_ctx *ctx, int on) { if (!ctx) return MTR_EINVAL; ctx->blocking_sync = on != 0; return MTR_OK; }

int mtr_reads_upload(mtr_ctx *ctx, const uint32_t *packed, const int64_t *word_off, const int32_t *len, int n_reads)
{
    if (!ctx) return MTR_EINVAL;
    if (n_reads < 0 || (n_reads > 0 && (!packed || !word_off || !len))) { mtr_set_error(ctx, "reads_upload: null argument"); return MTR_EINVAL; }
    for (int r = 0; r < n_reads; r++) {
        if (len[r] < 0 || len[r] >= 1000000) { mtr_set_error(ctx, "reads_upload: read %d has length %d", r, len[r]); return MTR_ERANGE; }
        if (word_off[r + 1] - word_off[r] < (len[r] + 2 + 15) / 16) { mtr_set_error(ctx, "reads_upload: read %d: words do not cover len + 2 bases", r); return MTR_EINVAL; }
    }
    auto rd = std::make_shared<SimReads>();
    const int64_t nw = n_reads ? word_off[n_reads] : 0;
    rd->packed.assign(packed, packed + nw);
    rd->packed.resize((size_t)nw + 16, 0u);
    rd->word_off.assign(word_off, word_off + (n_reads ? n_reads + 1 : 0));
    rd->len.assign(len, len + n_reads);
    ctx->di->reads = rd;
    ctx->word_off = rd->word_off; ctx->len = rd->len; ctx->n_reads = n_reads; ctx->n_words = nw;
    return MTR_OK;
}

int mtr_reads_share(mtr_ctx *dst, const mtr_ctx *src)
{
    if (!dst || !src || dst == src) return MTR_EINVAL;
    if (dst->device != src->device) { mtr_set_error(dst, "reads_share: contexts are on different devices"); return MTR_EINVAL; }
    dst->di->reads = src->di->reads;
    dst->word_off = src->word_off; dst->len = src->len; dst->n_reads = src->n_reads; dst->n_words = src->n_words;
    return MTR_OK;
}

// fill_directional_index_with_end for every resident read, each with exactly the stale state the host passed in
int mtr_di_run_range(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off,
                     double *di, int32_t *end, int32_t *w, int first, int count);
int mtr_di_run(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off,
               double *di, int32_t *end, int32_t *w)
{
    if (!ctx) return MTR_EINVAL;
    return mtr_di_run_range(ctx, manhattan, stale, stale_off, pos_off, di, end, w, 0, ctx->n_reads);
}

int mtr_di_run_range(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off,
                     double *di, int32_t *end, int32_t *w, int first, int count)
{
    if (!ctx || first < 0 || count < 0 || first + count > ctx->n_reads) return MTR_EINVAL;
    if (count == 0) return MTR_OK;
    if (!pos_off || !end || !w) return MTR_EINVAL;
    const SimReads &rd = *ctx->di->reads;
    mtro_ctx *o = ctx->di->get(manhattan ? 1 : 0);
    int *org = mtro_org_mut(o), *padded = mtro_padded_mut(o);
    std::vector<int> bases;
    std::vector<double> dtmp;
    long long passes = 0;
    for (int r = first; r < first + count; r++) {
        const int L = rd.len[r];
        bases.resize(L + 2);
        for (int i = 0; i < L + 2; i++) bases[i] = rd.base(r, i);
        mtro_load_read(o, bases.data(), L);
        org[L] = bases[L]; org[L + 1] = bases[L + 1];
        const int rl = L < 1000 ? 100 : L / 10;
        const long long ext = std::min<long long>((long long)L + 4 * rl, 1000000);
        const long long ns = stale && stale_off ? stale_off[r + 1] - stale_off[r] : 0;
        const long long hi = std::min<long long>(3000000LL, ext + std::max<long long>(ns, 0) + 65536);
        for (long long x = ext; x < hi; x++) padded[x] = 0;
        for (long long i = 0; i < ns && ext + i < 3000000LL; i++) padded[ext + i] = stale[stale_off[r] + i];
        dtmp.resize((size_t)L + 1);
        mtro_directional_index(o, di ? di + pos_off[r] : dtmp.data(), end + pos_off[r], w + pos_off[r]);
        mtro_stats st;
        mtro_get_stats(o, &st);
        passes = st.di_position_passes;
    }
    ctx->stats.di_ms = 0; ctx->stats.di_position_passes = passes; ctx->stats.launches = 0;
    ctx->stats.di_bytes_in = 0; ctx->stats.di_bytes_out = 0;
    return MTR_OK;
}

int mtr_wdp_run(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                mtr_wdp_result *results, void *aux, int64_t aux_bytes)
{
    if (!ctx || n_jobs < 0 || (n_jobs > 0 && (!jobs || !units || !results))) return MTR_EINVAL;
    const SimReads &rd = *ctx->di->reads;
    mtro_ctx *o = ctx->di->get(ctx->di->oracle_manhattan < 0 ? 1 : ctx->di->oracle_manhattan);
    std::vector<int> x, u;
    std::vector<unsigned char> path;
    long long cells = 0;
    for (int j = 0; j < n_jobs; j++) {
        const mtr_wdp_job &q = jobs[j];
        if (q.read < 0 || q.read >= ctx->n_reads || q.ulen < 1 || q.ulen > 499 || q.rows < 0 || q.n_param < 1 || q.n_param > 2 ||
            q.unit_off < 0 || q.unit_off + q.ulen > units_len || (q.mode != MTR_TB_COUNTS && q.n_param != 1)) {
            mtr_set_error(ctx, "hostsim wdp_run: malformed job %d", j);
            return MTR_EINVAL;
        }
        x.assign(q.rows + 2, 0);
        for (int i = 1; i <= q.rows; i++) x[i] = rd.base(q.read, (long long)q.first + i);
        u.assign(q.ulen + 2, 0);
        for (int i = 1; i <= q.ulen; i++) u[i] = units[q.unit_off + i - 1];
        for (int s = 0; s < q.n_param; s++) {
            mtro_dp_result r;
            int *cons = nullptr, *miss = nullptr;
            unsigned char *pp = nullptr;
            if (q.mode == MTR_TB_CONSENSUS) {
                if (!aux || (q.aux_off + (int64_t)(q.ulen + 1) * 9) * 4 > aux_bytes) return MTR_EINVAL;
                cons = (int *)aux + q.aux_off; miss = cons + (size_t)(q.ulen + 1) * 5;
                memset(cons, 0, (size_t)(q.ulen + 1) * 9 * 4);
            } else if (q.mode == MTR_TB_PATH) {
                path.assign((size_t)q.rows * (q.ulen + 1) + q.rows + q.ulen + 64, 0);
                pp = path.data();
            }
            mtro_wrap_dp(o, x.data(), q.rows, u.data(), q.ulen, q.gain[s], q.mis[s], q.indel[s], q.mode, &r, cons, miss, pp, nullptr);
            mtr_wdp_result &d = results[2 * j + s];
            d.best = r.best; d.max_i = r.max_i; d.max_j = r.max_j; d.end_i = r.end_i; d.end_j = r.end_j;
            d.n_match = r.n_match; d.n_mismatch = r.n_mismatch; d.n_ins = r.n_ins; d.n_del = r.n_del; d.n_scanned = r.n_scanned;
            d.path_len = r.path_len; d.flags = 0;
            if (q.mode == MTR_TB_PATH) {
                if (!aux || q.aux_off + q.aux_cap > aux_bytes) return MTR_EINVAL;
                int n = r.path_len;
                if (n > q.aux_cap) { n = (int)q.aux_cap; d.flags |= 1; d.path_len = n; }
                memcpy((unsigned char *)aux + q.aux_off, pp, (size_t)n);
            }
            cells += (long long)q.rows * q.ulen;
        }
    }
    ctx->stats.wdp_fill_ms = 0; ctx->stats.wdp_tb_ms = 0; ctx->stats.wdp_cells = cells; ctx->stats.wdp_slot_cells = cells;
    ctx->stats.wdp_dir_bytes = 0; ctx->stats.launches = 0;
    return MTR_OK;
}

int mtr_uf_run(mtr_ctx *ctx, const mtr_uf_task *, int, mtr_uf_result *, uint8_t *, int32_t *, int64_t, int64_t *)
{
    mtr_set_error(ctx, "hostsim: the K4 unit finder is not simulated");
    return MTR_EINVAL;
}

// entry points the host pipeline never calls; present so that the ctypes binding (mtr_b200/capi.py) loads this library
int mtr_wdp_upload(mtr_ctx *ctx, const mtr_wdp_job *, int, const uint8_t *, int64_t, int64_t) { mtr_set_error(ctx, "hostsim: not simulated"); return MTR_EINVAL; }
int mtr_wdp_launch(mtr_ctx *ctx) { mtr_set_error(ctx, "hostsim: not simulated"); return MTR_EINVAL; }
int mtr_wdp_download(mtr_ctx *ctx, mtr_wdp_result *, void *, int64_t) { mtr_set_error(ctx, "hostsim: not simulated"); return MTR_EINVAL; }
int mtr_wdp_set_fused_traceback(mtr_ctx *ctx, int) { return ctx ? MTR_OK : MTR_EINVAL; }
int mtr_alu_probe(mtr_ctx *ctx, int, double *gops) { if (!ctx || !gops) return MTR_EINVAL; *gops = 1.0; return MTR_OK; }

int mtr_get_stats(const mtr_ctx *ctx, mtr_stats *out) { if (!ctx || !out) return MTR_EINVAL; *out = ctx->stats; return MTR_OK; }

}  // extern "C"
