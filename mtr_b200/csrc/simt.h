// simt.h -- the few warp-level primitives the resident engine (eng_core.h) is written against.
//
// The engine's device code is "warp-uniform SPMD": every lane of a warp executes the same scalar control flow on the
// same values; only the data-parallel loops stride by lane (`for (i = lo + lane(); i < hi; i += NL)`), and whatever
// they compute reaches the control flow again through one of the reductions below.  Compiled by nvcc the primitives
// are the sm_100a warp intrinsics; compiled by a plain C++ compiler (tests/hostsim: the host-logic tests that run
// without a GPU) a warp has ONE lane and every primitive degenerates to the identity, so the very same engine code
// is checked against the reference digests on the CPU, with the DP and the directional index answered by the oracle.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MTR_DEV __device__ __forceinline__
#define MTR_DEV_NOINLINE __device__ __noinline__
#define MTR_CONST __device__ const
namespace simt {
constexpr int NL = 32;
constexpr unsigned FULL = 0xffffffffu;
MTR_DEV int lane() { return (int)(threadIdx.x & 31u); }
MTR_DEV void wsync() { __syncwarp(); }
MTR_DEV unsigned wballot(bool p) { return __ballot_sync(FULL, p); }
MTR_DEV bool wany(bool p) { return __any_sync(FULL, p) != 0; }
MTR_DEV unsigned lanemask_lt() { return (1u << lane()) - 1u; }
MTR_DEV int popc(unsigned v) { return __popc(v); }
MTR_DEV int ffs(unsigned v) { return __ffs((int)v); }                 // 1-based, 0 if none
MTR_DEV int clz(unsigned v) { return __clz((int)v); }
MTR_DEV int bcast(int v, int src) { return __shfl_sync(FULL, v, src); }
MTR_DEV unsigned bcast(unsigned v, int src) { return __shfl_sync(FULL, v, src); }
MTR_DEV long long bcast(long long v, int src) { return __shfl_sync(FULL, v, src); }
MTR_DEV unsigned match_any(unsigned v) { return __match_any_sync(FULL, v); }
MTR_DEV int wmax(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
MTR_DEV int wmin(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
MTR_DEV int wsum(int v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
MTR_DEV long long wsum(long long v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
MTR_DEV int wscan_excl(int v)                   // exclusive prefix sum over the lanes
{
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane() >= o) x += y; }
    return x - v;
}
MTR_DEV int atomic_cas(int *p, int cmp, int v) { return atomicCAS(p, cmp, v); }
MTR_DEV unsigned atomic_cas(unsigned *p, unsigned cmp, unsigned v) { return atomicCAS(p, cmp, v); }
MTR_DEV unsigned atomic_add(unsigned *p, unsigned v) { return atomicAdd(p, v); }
MTR_DEV unsigned ldv(const unsigned *p) { return *(const volatile unsigned *)p; }
MTR_DEV void block_sync() { __syncthreads(); }
MTR_DEV int block_tid() { return (int)threadIdx.x; }
MTR_DEV int block_size() { return (int)blockDim.x; }
MTR_DEV long long clock_now() { return clock64(); }
MTR_DEV int atomic_add(int *p, int v) { return atomicAdd(p, v); }
MTR_DEV unsigned long long atomic_add(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
MTR_DEV unsigned long long atomic_cas(unsigned long long *p, unsigned long long cmp, unsigned long long v) { return atomicCAS(p, cmp, v); }
MTR_DEV int atomic_max(int *p, int v) { return atomicMax(p, v); }
MTR_DEV int atomic_exch(int *p, int v) { return atomicExch(p, v); }
MTR_DEV unsigned long long ldv(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
MTR_DEV int ldv(const int *p) { return *(const volatile int *)p; }
MTR_DEV void fence() { __threadfence(); }
}  // namespace simt
#else
#define MTR_DEV inline
#define MTR_DEV_NOINLINE inline
#define MTR_CONST static const
namespace simt {
constexpr int NL = 1;
inline int lane() { return 0; }
inline void wsync() {}
inline unsigned wballot(bool p) { return p ? 1u : 0u; }
inline bool wany(bool p) { return p; }
inline unsigned lanemask_lt() { return 0u; }
inline int popc(unsigned v) { return __builtin_popcount(v); }
inline int ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int bcast(int v, int) { return v; }
inline unsigned bcast(unsigned v, int) { return v; }
inline long long bcast(long long v, int) { return v; }
inline unsigned match_any(unsigned) { return 1u; }
inline int wmax(int v) { return v; }
inline int wmin(int v) { return v; }
inline int wsum(int v) { return v; }
inline long long wsum(long long v) { return v; }
inline int wscan_excl(int) { return 0; }
inline int atomic_cas(int *p, int cmp, int v) { const int o = *p; if (o == cmp) *p = v; return o; }
inline unsigned atomic_cas(unsigned *p, unsigned cmp, unsigned v) { const unsigned o = *p; if (o == cmp) *p = v; return o; }
inline unsigned atomic_add(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline unsigned ldv(const unsigned *p) { return *p; }
inline void block_sync() {}
inline int block_tid() { return 0; }
inline int block_size() { return 1; }
inline long long clock_now() { return 0; }
inline int atomic_add(int *p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned long long atomic_add(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomic_cas(unsigned long long *p, unsigned long long cmp, unsigned long long v) { const unsigned long long o = *p; if (o == cmp) *p = v; return o; }
inline int atomic_max(int *p, int v) { const int o = *p; if (v > o) *p = v; return o; }
inline int atomic_exch(int *p, int v) { const int o = *p; *p = v; return o; }
inline unsigned long long ldv(const unsigned long long *p) { return *p; }
inline int ldv(const int *p) { return *p; }
inline void fence() {}
}  // namespace simt
#endif
