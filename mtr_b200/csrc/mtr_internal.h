// mtr_internal.h -- shared declarations of libmtr_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/mtr_b200.h"

// ---------------------------------------------------------------- error plumbing
void mtr_set_error(mtr_ctx *ctx, const char *fmt, ...);
#define MTR_CUDA(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            mtr_set_error((ctx), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                          cudaGetErrorString(e_));                                              \
            return MTR_ECUDA;                                                                   \
        }                                                                                       \
    } while (0)

// ---------------------------------------------------------------- growable device / pinned buffers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;      // points into another context's allocation
    void borrow(void *q, size_t c) { release(); p = q; cap = c; borrowed = true; }
    cudaError_t reserve(size_t bytes)
    {
        if (borrowed) { p = nullptr; cap = 0; borrowed = false; }
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 4 + 4096;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    cudaError_t reserve_exact(size_t bytes)     // no extra headroom: the caller chose the size
    {
        if (borrowed) { p = nullptr; cap = 0; borrowed = false; }
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 4 + 4096;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ---------------------------------------------------------------- wrap-around DP (wdp.cu)
// One fill task = one (job, penalty set) pair, or -- in the paired int16x2 kernels -- one job with both sets.
struct WdpTask {
    long long base0;      // x_i = base (base0 + i) of the packed read buffer
    long long dir_off;    // byte offset of this task's direction matrix (second set: + dir_bytes)
    long long dir_bytes;  // bytes of one direction matrix = rows * dir_stride
    long long aux_off;    // CONSENSUS: int32 offset, PATH: byte offset
    long long aux_cap;
    long long unit_off;   // offset of the unit (one byte per base) in the units array
    int rows, ulen;
    int dir_stride;       // bytes per row of the direction matrix (= slots / 4)
    int result_idx;       // results[result_idx (+1 for the second set of a paired task)]
    short gain[2], mis[2], indel[2];
    unsigned char n_param, mode;
    unsigned char cls;    // resident engine: fill class 0..9 (int32) / 10..19 (paired int16x2), see eng_core.h
    unsigned char pad_;
};

constexpr int WDP_NCLASS = 26;
constexpr int kWdpOwnerShift = 3;     // engine: result slot r belongs to owner r >> kWdpOwnerShift (eight result slots per chain)
constexpr int kWdpRowBuckets = 96;   // quarter-octave buckets of a task's row count in the engine's sorted task lists (eng_core.h kRowBuckets)
struct WdpClass { int G, C, paired; };          // lanes per job, cells per lane, int16x2 pairing

struct WdpState {
    DevBuf d_tasks, d_dirs, d_results, d_aux, d_counters;      // d_tasks holds [WdpTask x n | unit bases]
    PinBuf h_tasks, h_results;
    std::vector<WdpTask> tasks;                 // sorted by class, then by rows descending
    int class_begin[WDP_NCLASS + 1] = {0};
    int n_jobs = 0, n_results = 0;
    long long dir_total = 0, aux_bytes = 0;
    long long cells = 0, slot_cells = 0;
    bool uploaded = false;
    size_t units_dev_off = 0;                   // byte offset of the unit bases inside d_tasks
    int latency = 0;                            // the uploaded batch uses the latency classes (one fused launch)
    bool fused_tb = true;                       // traceback inside the fill kernels (mtr_wdp_set_fused_traceback)
    unsigned counter_gen = 0;
    std::vector<int> unused_results;            // result slots no task writes (second slot of one-set jobs)
};

// ---------------------------------------------------------------- context
struct mtr_ctx {
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream[WDP_NCLASS] = {};
    cudaStream_t main_stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t class_done[WDP_NCLASS] = {};
    cudaEvent_t sync_ev = nullptr;          // blocking-sync event: waiting threads sleep instead of spinning
    bool blocking_sync = false;             // mtr_set_blocking_sync
    std::string err;
    // resident reads
    DevBuf d_packed, d_word_off, d_len;
    std::vector<int64_t> word_off;
    std::vector<int32_t> len;
    int n_reads = 0;
    int64_t n_words = 0;
    WdpState wdp;
    struct DiState *di = nullptr;
    struct UfState *uf = nullptr;
    struct EngState *eng = nullptr;
    mtr_stats stats = {};
    double prof_s[3] = {0, 0, 0};           // MTR_PROFILE: host seconds in upload / launch / download+wait of mtr_wdp_run
    long long prof_n = 0;
    double prof_up[3] = {0, 0, 0};          // ... inside upload: classify + sort / buffer growth / staging copy + enqueue
};

// sync == false: only enqueue on main_stream (mtr_wdp_run waits once, at the end)
int  wdp_upload_impl(mtr_ctx *ctx, const mtr_wdp_job *jobs, int n_jobs, const uint8_t *units, int64_t units_len,
                     int64_t aux_bytes, bool sync);
int  wdp_launch_impl(mtr_ctx *ctx, bool sync);
int  wdp_download_impl(mtr_ctx *ctx, mtr_wdp_result *results, void *aux, int64_t aux_bytes, bool read_times);
// waits for everything queued on main_stream without burning a host core (the dispatcher threads of the pipeline
// would otherwise spin next to the host workers)
inline cudaError_t mtr_sync(mtr_ctx *ctx)
{
    if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->main_stream);
    cudaError_t e = cudaEventRecord(ctx->sync_ev, ctx->main_stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->sync_ev);
}
// launch description of the K3 kernels over a task list that lives in device memory (resident engine, wdp.cu)
struct WdpDevLaunch {
    const WdpTask *tasks;
    const int *class_begin;       // [WDP_NCLASS + 1]: the last entry is the number of tasks
    const int *seg_task, *seg_slot;   // prefix sums over the 2 * nseg_family segments (+1) of the sorted task list
    int nseg_family;
    int *counters;                // [WDP_NCLASS] slot-queue heads, zero at launch
    const unsigned long long *class_share;   // [20] estimated work per (family, fill class): SM s starts with the class its share of the SMs covers
    const uint32_t *packed;
    const uint8_t *units;
    uint8_t *dirs;
    mtr_wdp_result *results;
    void *aux;
    char *pending0;               // &owner[0].pending; owner of result r is r / 4 (nullptr: nobody counts)
    int pending_stride;           // sizeof(owner)
    int *pending_total;
    int blocks;                   // persistent grid of every kernel ...
    int blocks_i32, blocks_p16;   // ... unless the caller knows how many warp slots each fill family has (< 0: use `blocks`)
    int fused;                    // tracebacks inside the fill kernels (no traceback kernel)
    int fill_smem;                // dynamic shared memory asked for by the fill kernels (bounds their blocks per SM)
    unsigned long long *prof;     // MTR_PROFILE: 10 counters (see WdpTb), nullptr = off
    cudaStream_t side[8];
    cudaEvent_t fork, join[8];
    int n_side;
};
cudaError_t wdp_launch_dev(const WdpDevLaunch &L, cudaStream_t s);
int  di_compute(mtr_ctx *ctx, int manhattan, const uint16_t *stale, const int64_t *stale_off, const int64_t *pos_off, int first, int count);
void di_device_outputs(mtr_ctx *ctx, double **di, int **end, int **w);
void di_state_free(mtr_ctx *ctx);
void uf_state_free(mtr_ctx *ctx);
void eng_state_free(mtr_ctx *ctx);
long long wdp_dir_bytes(int ulen, int rows);   // bytes of one direction matrix (per penalty set)
