// cosmomap2_b200 -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cosmomap2_b200.h"

namespace cm2 {

// ---- host side --------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define CM2_REQUIRE(cond, msg)                                                     \
    do {                                                                           \
        if (!(cond)) return cm2::set_error(CM2_ERR_ARG, "%s: %s", __func__, msg);  \
    } while (0)

#define CM2_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return cm2::set_error(CM2_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
    } while (0)

#define CM2_LAUNCHED()                                                                        \
    do {                                                                                      \
        cm2::count_launch();                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return cm2::set_error(CM2_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
    } while (0)

static inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline cudaStream_t as_stream(cm2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a persistent grid-stride kernel: a multiple of the SM count
template <class Kernel>
static inline int persistent_grid(Kernel k, int block, size_t smem, int64_t work_blocks, int max_per_sm = 8) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, block, smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    int64_t g = (int64_t)sm_count() * per_sm;
    if (work_blocks < g) g = work_blocks < 1 ? 1 : work_blocks;
    return (int)g;
}

// ---- device side: streaming loads/stores (TOD is read once: keep it out of the way of the
// L2-resident map) --------------------------------------------------------------------------
#ifdef __CUDACC__
struct alignas(32) I8 { int v[8]; };
struct alignas(32) D4 { double v[4]; };

// 256-bit global loads (LDG.E.256, sm_100+), L1 no-allocate, L2 evict-first
__device__ __forceinline__ I8 ld_stream_i8(const int32_t *p) {
    I8 r;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ D4 ld_stream_d4(const double *p) {
    D4 r;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(p));
    return r;
}
// same, but leave the line in L2 for a second pass of the same CTA (fused filter kernel)
__device__ __forceinline__ I8 ld_keep_i8(const int32_t *p) {
    I8 r;
    asm volatile("ld.global.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ D4 ld_keep_d4(const double *p) {
    D4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_d4(double *p, const D4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.v[0]), "d"(v.v[1]),
                 "d"(v.v[2]), "d"(v.v[3])
                 : "memory");
}
// scalar tail loads: the .L2::evict_first priority form only exists for the 256-bit vectors
__device__ __forceinline__ int ld_stream_i1(const int32_t *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream_d1(const double *p) { return __ldcs(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum: fixed shuffle tree + fixed order over warps.  `red` = 32 doubles smem.
__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;  // valid in warp 0
}
#endif

}  // namespace cm2
