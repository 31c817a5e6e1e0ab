// alu_probe.cu -- measures the integer-ALU issue ceiling the wrap-around DP is rooflined against
// (ceiling = lane-ops/s of the DPX / integer pipe divided by the instructions per cell).
#include "mtr_internal.h"

template <int KIND>
__global__ void __launch_bounds__(256) alu_probe_kernel(int iters, int seed, int *out)
{
    int a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int b = seed | 1, c = seed + 3;
    for (int i = 0; i < iters; i++) {
        if (KIND == 0) {            // VIADDMNMX.RELU, the DP's work-horse
            a0 = __viaddmax_s32_relu(a0, b, c); a1 = __viaddmax_s32_relu(a1, b, c); a2 = __viaddmax_s32_relu(a2, b, c);
            a3 = __viaddmax_s32_relu(a3, b, c); a4 = __viaddmax_s32_relu(a4, b, c); a5 = __viaddmax_s32_relu(a5, b, c);
            a6 = __viaddmax_s32_relu(a6, b, c); a7 = __viaddmax_s32_relu(a7, b, c);
        } else if (KIND == 1) {     // LOP3 (plain integer ALU op)
            a0 = (a0 ^ b) & ~c; a1 = (a1 ^ b) & ~c; a2 = (a2 ^ b) & ~c; a3 = (a3 ^ b) & ~c;
            a4 = (a4 ^ b) & ~c; a5 = (a5 ^ b) & ~c; a6 = (a6 ^ b) & ~c; a7 = (a7 ^ b) & ~c;
            a0 += i; a1 += i; a2 += i; a3 += i; a4 += i; a5 += i; a6 += i; a7 += i;      // keeps the chain live (counted below)
        } else {                    // VIADDMNMX.S16x2
            a0 = __viaddmax_s16x2_relu(a0, b, c); a1 = __viaddmax_s16x2_relu(a1, b, c); a2 = __viaddmax_s16x2_relu(a2, b, c);
            a3 = __viaddmax_s16x2_relu(a3, b, c); a4 = __viaddmax_s16x2_relu(a4, b, c); a5 = __viaddmax_s16x2_relu(a5, b, c);
            a6 = __viaddmax_s16x2_relu(a6, b, c); a7 = __viaddmax_s16x2_relu(a7, b, c);
        }
    }
    const int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x7fffffff) out[0] = r;
}

// Returns giga lane-operations per second for the chosen instruction kind (0 DPX s32, 1 LOP3+IADD, 2 DPX s16x2).
extern "C" int mtr_alu_probe(mtr_ctx *ctx, int kind, double *gops)
{
    if (!ctx || !gops || kind < 0 || kind > 2) return MTR_EINVAL;
    MTR_CUDA(ctx, cudaSetDevice(ctx->device));
    MTR_CUDA(ctx, ctx->wdp.d_counters.reserve(sizeof(int) * WDP_NCLASS));
    const int iters = 8192, blocks = ctx->n_sm * 8, threads = 256;
    cudaStream_t s = ctx->main_stream;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        MTR_CUDA(ctx, cudaEventRecord(ctx->ev[5], s));
        if (kind == 0) alu_probe_kernel<0><<<blocks, threads, 0, s>>>(iters, rep + 1, (int *)ctx->wdp.d_counters.p);
        else if (kind == 1) alu_probe_kernel<1><<<blocks, threads, 0, s>>>(iters, rep + 1, (int *)ctx->wdp.d_counters.p);
        else alu_probe_kernel<2><<<blocks, threads, 0, s>>>(iters, rep + 1, (int *)ctx->wdp.d_counters.p);
        MTR_CUDA(ctx, cudaGetLastError());
        MTR_CUDA(ctx, cudaEventRecord(ctx->ev[6], s));
        MTR_CUDA(ctx, mtr_sync(ctx));
        float ms = 0;
        MTR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
        if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * iters * (kind == 1 ? 16.0 : 8.0);
    *gops = ops / (best * 1e-3) / 1e9;
    return MTR_OK;
}
