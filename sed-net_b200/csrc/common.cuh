// Shared device helpers for the sm_100a kernels of the SED-Net hot path.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define SED_OK 0
#define SED_ERR_ARG (-1)
#define SED_ERR_UNSUPPORTED (-2)
#define SED_ERR_GUARD (-3)
#define SED_ERR_CUDA_BASE (-1000)  // -(1000 + cudaError_t)

// counts kernel launches issued by the library (sed_launch_count); atomic: the entry points may be called from
// several host threads
extern std::atomic<long long> g_sed_launches;

#define SED_CHECK_LAUNCH()                                   \
    do {                                                     \
        ++g_sed_launches;                                    \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return SED_ERR_CUDA_BASE - (int)e__; \
    } while (0)

// SEDNET_B200_DEBUG_SYNC=1: synchronise the device after every step and name the first one that faults (debugging aid)
inline bool sed_debug_sync() {
    static const bool on = getenv("SEDNET_B200_DEBUG_SYNC") != nullptr;
    return on;
}

#define SED_TRY(call)                                                                         \
    do {                                                                                      \
        int rc__ = (call);                                                                    \
        if (rc__ != SED_OK) return rc__;                                                      \
        if (sed_debug_sync()) {                                                               \
            cudaError_t se__ = cudaDeviceSynchronize();                                       \
            if (se__ != cudaSuccess) {                                                        \
                fprintf(stderr, "[sednet_b200] %s:%d %s -> %s\n", __FILE__, __LINE__, #call, \
                        cudaGetErrorString(se__));                                            \
                return SED_ERR_CUDA_BASE - (int)se__;                                         \
            }                                                                                 \
        }                                                                                     \
    } while (0)

#define SED_CUDA(call)                                       \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return SED_ERR_CUDA_BASE - (int)e__; \
    } while (0)

namespace sed {

constexpr int kNumSMs = 148;  // B200

// Order-preserving map float -> uint32 (larger float <=> larger uint).
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((unsigned long long)hi << 32) | lo;
}

// Block-wide sum of NV doubles per thread; result valid in thread 0 (and broadcast through smem `out`).
// scratch must hold NV * (blockDim.x / 32) doubles.
template <int NV>
__device__ __forceinline__ void block_sum_d(double (&v)[NV], double* scratch, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum_d(v[i]);
        if (lane == 0) scratch[warp * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += scratch[w * NV + threadIdx.x];  // fixed order: deterministic
        out[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = out[i];
    __syncthreads();
}

}  // namespace sed
