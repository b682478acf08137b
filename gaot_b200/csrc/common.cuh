// Shared device/host helpers for the gaot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define GAOT_API extern "C" __attribute__((visibility("default")))

// Every entry point is asynchronous on `stream` and returns a cudaError_t-compatible int (0 = ok).
// Process-wide count of kernels this library has launched (diagnostics only: bench.py reports it as
// `gpu_launches`).  GAOT_RETURN_LAST counts the one kernel most entry points launch; entry points that
// launch more add the rest with GAOT_NOTE_LAUNCHES.
extern "C" long long gaot_launch_counter_add(long long n);
#define GAOT_NOTE_LAUNCHES(n) gaot_launch_counter_add(n)
#define GAOT_RETURN_LAST() do { gaot_launch_counter_add(1); return (int)cudaGetLastError(); } while (0)
#define GAOT_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

static inline int gaot_blocks(int64_t work, int threads) { return (int)((work + threads - 1) / threads); }

// SM count of the current device (148 on B200), queried once per process
static inline int gaot_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// Grid size for grid-stride kernels: a multiple of the SM count, capped by the work.
static inline int gaot_grid(int64_t work, int threads, int ctas_per_sm = 8) {
    int64_t need = (work + threads - 1) / threads;
    int64_t cap = (int64_t)gaot_sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load that does not allocate in L1 (data read once per CTA)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}
