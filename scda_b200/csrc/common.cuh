// Shared device/host helpers for libscda_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scda_b200.h"

#define SCDA_API extern "C" __attribute__((visibility("default")))

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Status convention of include/scda_b200.h.
static inline int scda_launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Streaming 128-bit store: outputs that are written once and not re-read by
// this kernel should not displace the L2-resident feature map.
__device__ __forceinline__ void st_stream_f4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_i4(int *p, int4 v)
{
    asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f32(float *p, float v)
{
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_s32(int *p, int v)
{
    asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream_i4(const int *p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
// Fire-and-forget fp32 add resolved in L2.
__device__ __forceinline__ void red_add_f32(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
