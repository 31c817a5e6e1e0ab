// ctx.cu -- context, resident read batch and the kernel-level C ABI of libmtr_b200.so.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <time.h>
#include "mtr_internal.h"

static std::string g_init_error;

static double wall_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

void mtr_set_error(mtr_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
}

extern "C" const char *mtr_last_error(const mtr_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int mtr_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int mtr_cuda_init(int device, mtr_ctx **out)
{
    if (!out) return MTR_EINVAL;
    *out = nullptr;
    // many engine contexts share one GPU, each with its own streams: ask for the maximum number of hardware work queues
    // (the default of 8 makes kernels of unrelated streams wait for each other); only effective before CUDA starts
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        mtr_set_error(nullptr, "mtr_cuda_init: no CUDA device (%s); this library has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return MTR_ENODEV;
    }
    if (device < 0 || device >= n) { mtr_set_error(nullptr, "mtr_cuda_init: device %d out of range (%d devices)", device, n); return MTR_EINVAL; }
    mtr_ctx *ctx = new mtr_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        mtr_set_error(nullptr, "mtr_cuda_init: %s", cudaGetErrorString(e));
        delete ctx;
        return MTR_ENODEV;
    }
    if (prop.major < 10) {
        mtr_set_error(nullptr, "mtr_cuda_init: device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major, prop.minor);
        delete ctx;
        return MTR_ENODEV;
    }
    ctx->n_sm = prop.multiProcessorCount;
    ctx->stats.n_sm = ctx->n_sm;
    bool ok = cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < WDP_NCLASS; k++) {
        ok = cudaStreamCreateWithFlags(&ctx->stream[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->class_done[k], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int k = 0; ok && k < 8; k++) ok = cudaEventCreate(&ctx->ev[k]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        mtr_set_error(nullptr, "mtr_cuda_init: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return MTR_ECUDA;
    }
    *out = ctx;
    return MTR_OK;
}

extern "C" void mtr_cuda_shutdown(mtr_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->prof_n)
        fprintf(stderr, "[mtr profile] ctx %p: %lld wdp_run calls, host ms per call: upload %.3f (classify+sort %.3f, buffers %.3f, stage+enqueue %.3f) launch %.3f download+wait %.3f\n", (void *)ctx,
                ctx->prof_n, ctx->prof_s[0] / ctx->prof_n * 1e3, ctx->prof_up[0] / ctx->prof_n * 1e3, ctx->prof_up[1] / ctx->prof_n * 1e3,
                ctx->prof_up[2] / ctx->prof_n * 1e3, ctx->prof_s[1] / ctx->prof_n * 1e3, ctx->prof_s[2] / ctx->prof_n * 1e3);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    di_state_free(ctx);
    uf_state_free(ctx);
    eng_state_free(ctx);
    WdpState &w = ctx->wdp;
    w.d_tasks.release(); w.d_dirs.release(); w.d_results.release(); w.d_aux.release();
    w.d_counters.release(); w.h_tasks.release(); w.h_results.release();
    ctx->d_packed.release(); ctx->d_word_off.release(); ctx->d_len.release();
    for (int k = 0; k < WDP_NCLASS; k++) {
        if (ctx->stream[k]) cudaStreamDestroy(ctx->stream[k]);
        if (ctx->class_done[k]) cudaEventDestroy(ctx->class_done[k]);
    }
    for (int k = 0; k < 8; k++) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
    if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    delete ctx;
}

extern "C" int mtr_reads_upload(mtr_ctx *ctx, const uint32_t *packed, const int64_t *word_off, const int32_t *len, int n_reads)
{
    if (!ctx) return MTR_EINVAL;
    if (n_reads < 0 || (n_reads > 0 && (!packed || !word_off || !len))) { mtr_set_error(ctx, "reads_upload: null argument"); return MTR_EINVAL; }
    for (int r = 0; r < n_reads; r++) {
        if (len[r] < 0 || len[r] >= 1000000) { mtr_set_error(ctx, "reads_upload: read %d has length %d (MAX_INPUT_LENGTH is 1000000)", r, len[r]); return MTR_ERANGE; }
        if (word_off[r + 1] - word_off[r] < (len[r] + 2 + 15) / 16) { mtr_set_error(ctx, "reads_upload: read %d: words do not cover len + 2 bases", r); return MTR_EINVAL; }
    }
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nw = n_reads ? word_off[n_reads] : 0;
    // 16 spare words so that vectorised (uint4) loads of the last read never leave the allocation
    MTR_CUDA(ctx, ctx->d_packed.reserve((size_t)(nw + 16) * 4));
    MTR_CUDA(ctx, ctx->d_word_off.reserve((size_t)(n_reads + 1) * 8));
    MTR_CUDA(ctx, ctx->d_len.reserve((size_t)(n_reads + 1) * 4));
    if (n_reads > 0) {
        MTR_CUDA(ctx, cudaMemcpyAsync(ctx->d_packed.p, packed, (size_t)nw * 4, cudaMemcpyHostToDevice, ctx->main_stream));
        MTR_CUDA(ctx, cudaMemsetAsync((char *)ctx->d_packed.p + (size_t)nw * 4, 0, 64, ctx->main_stream));
        MTR_CUDA(ctx, cudaMemcpyAsync(ctx->d_word_off.p, word_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->main_stream));
        MTR_CUDA(ctx, cudaMemcpyAsync(ctx->d_len.p, len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->main_stream));
        MTR_CUDA(ctx, mtr_sync(ctx));
    }
    ctx->word_off.assign(word_off, word_off + (n_reads ? n_reads + 1 : 0));
    ctx->len.assign(len, len + n_reads);
    ctx->n_reads = n_reads;
    ctx->n_words = nw;
    ctx->wdp.uploaded = false;
    return MTR_OK;
}

// Recreates the context's streams with a CUDA stream priority: level 0 = lowest (bulk work: directional index, long DP
// batches), higher levels = more urgent, clamped to what the device offers.  Pending thread blocks of a higher-priority
// stream are placed before those of lower-priority streams, so a 0.1 ms latency-class batch does not queue behind the
// thousands of blocks of a bulk kernel.  Call it before the context has work in flight.
extern "C" int mtr_set_priority(mtr_ctx *ctx, int level)
{
    if (!ctx) return MTR_EINVAL;
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    int least = 0, greatest = 0;
    MTR_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));     // numerically lower = higher priority
    int prio = least - std::max(0, level);
    if (prio < greatest) prio = greatest;
    MTR_CUDA(ctx, cudaStreamSynchronize(ctx->main_stream));
    MTR_CUDA(ctx, cudaStreamDestroy(ctx->main_stream));
    MTR_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->main_stream, cudaStreamNonBlocking, prio));
    for (int k = 0; k < WDP_NCLASS; k++) {
        MTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream[k]));
        MTR_CUDA(ctx, cudaStreamDestroy(ctx->stream[k]));
        MTR_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream[k], cudaStreamNonBlocking, prio));
    }
    return MTR_OK;
}

extern "C" int mtr_set_blocking_sync(mtr_ctx *ctx, int on)
{
    if (!ctx) return MTR_EINVAL;
    ctx->blocking_sync = on != 0;
    return MTR_OK;
}

extern "C" int mtr_reads_share(mtr_ctx *dst, const mtr_ctx *src)
{
    if (!dst || !src || dst == src) return MTR_EINVAL;
    if (dst->device != src->device) { mtr_set_error(dst, "reads_share: contexts are on different devices"); return MTR_EINVAL; }
    dst->d_packed.borrow(src->d_packed.p, src->d_packed.cap);
    dst->d_word_off.borrow(src->d_word_off.p, src->d_word_off.cap);
    dst->d_len.borrow(src->d_len.p, src->d_len.cap);
    dst->word_off = src->word_off;
    dst->len = src->len;
    dst->n_reads = src->n_reads;
    dst->n_words = src->n_words;
    dst->wdp.uploaded = false;
    return MTR_OK;
}

extern "C" int mtr_wdp_upload(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len, int64_t aux_bytes)
{
    if (!ctx) return MTR_EINVAL;
    return wdp_upload_impl(ctx, jobs, n_jobs, units, units_len, aux_bytes, true);
}

extern "C" int mtr_wdp_launch(mtr_ctx *ctx)
{
    if (!ctx) return MTR_EINVAL;
    return wdp_launch_impl(ctx, true);
}

extern "C" int mtr_wdp_download(mtr_ctx *ctx, mtr_wdp_result *results, void *aux, int64_t aux_bytes)
{
    if (!ctx) return MTR_EINVAL;
    return wdp_download_impl(ctx, results, aux, aux_bytes, false);
}

extern "C" int mtr_wdp_run(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                           mtr_wdp_result *results, void *aux, int64_t aux_bytes)
{
    if (!ctx) return MTR_EINVAL;
    // everything is enqueued on one stream; the host waits once, in the download
    static const bool prof = getenv("MTR_PROFILE") != nullptr;
    const double t0 = prof ? wall_s() : 0;
    int rc = wdp_upload_impl(ctx, jobs, n_jobs, units, units_len, aux_bytes, false);
    if (rc) return rc;
    const double t1 = prof ? wall_s() : 0;
    rc = wdp_launch_impl(ctx, false);
    if (rc) return rc;
    const double t2 = prof ? wall_s() : 0;
    rc = wdp_download_impl(ctx, results, aux, aux_bytes, true);
    if (prof) { ctx->prof_s[0] += t1 - t0; ctx->prof_s[1] += t2 - t1; ctx->prof_s[2] += wall_s() - t2; ctx->prof_n++; }
    return rc;
}

extern "C" int mtr_wdp_set_fused_traceback(mtr_ctx *ctx, int on)
{
    if (!ctx) return MTR_EINVAL;
    ctx->wdp.fused_tb = on != 0;
    return MTR_OK;
}

extern "C" int mtr_get_stats(const mtr_ctx *ctx, mtr_stats *out)
{
    if (!ctx || !out) return MTR_EINVAL;
    *out = ctx->stats;
    return MTR_OK;
}
