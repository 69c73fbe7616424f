// Companions of the fp32-parity precision mode ("bf16x3", scda_b200/tc.py: set_precision).
//
// The reference computes every convolution / linear layer in fp32 (cuDNN / cuBLAS of torch
// 0.4.1, no TF32: models/faster_rcnn/vgg_adver_expansion_cluster.py:46-60,101-114,
// models/head.py:13-18, common_net.py).  The tensor cores have no fp32 operand type; the
// parity mode keeps activations and gradients in fp32 NHWC and feeds the SAME tcgen05 kernels
// (conv_halo.cu, gemm_tc.cu, kind::f16, fp32 accumulation in TMEM) with operands split into two
// bf16 halves, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), concatenated along the
// reduction dimension:
//     A = [ a_hi | a_lo | a_hi ]      B = [ b_hi | b_hi | b_lo ]
//     sum_k A B = a_hi b_hi + a_lo b_hi + a_hi b_lo          (a_lo b_lo ~ 2^-16 |a b| is dropped)
// i.e. three MMAs per product and a relative error of ~2^-16 per product — 32 x below TF32's
// 2^-11 — at one third of the bf16 rate.  This file holds the HBM-bound passes around those
// MMAs: the operand split of activations (split3) and of weights (split_weights: the K-major
// form [rows, 3K] for the forward GEMMs and the row-stacked form [3 rows, K] that the data
// gradients read as an MN-major operand), and fp32 variants of the layout / pooling / column-sum
// kernels of nhwc_ops.cu.  All: 16-byte accesses, one pass over each tensor.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo)
{
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat162 h = __halves2bfloat162(ah, bh);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __bfloat162float(ah), b - __bfloat162float(bh));
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

__device__ __forceinline__ void split8(const float *p, uint4 &hi, uint4 &lo)
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    split_pair(a.x, a.y, hi.x, lo.x);
    split_pair(a.z, a.w, hi.y, lo.y);
    split_pair(b.x, b.y, hi.z, lo.z);
    split_pair(b.z, b.w, hi.w, lo.w);
}

// x fp32 [rows, C] (row stride ldx) -> y bf16 [rows, 3C] = [hi | lo | hi]; thread = (row, 8 channels)
__global__ void __launch_bounds__(256)
split3_kernel(const float *__restrict__ x, long long ldx, long long rows, int C, __nv_bfloat16 *__restrict__ y)
{
    const int cv = C >> 3;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cv;
        const int c = (int)(i - r * cv) * 8;
        uint4 hi, lo;
        split8(x + r * ldx + c, hi, lo);
        __nv_bfloat16 *o = y + r * 3 * C + c;
        *reinterpret_cast<uint4 *>(o) = hi;
        *reinterpret_cast<uint4 *>(o + C) = lo;
        *reinterpret_cast<uint4 *>(o + 2 * C) = hi;
    }
}

// w fp32 [rows, K] (row stride ldw) -> fwd bf16 [rows, 3K] = [hi | hi | lo] (either may be null)
//                                   -> stk bf16 [3 rows, K] = [hi ; hi ; lo]
__global__ void __launch_bounds__(256)
split_weights_kernel(const float *__restrict__ w, long long ldw, long long rows, int K,
                     __nv_bfloat16 *__restrict__ fwd, __nv_bfloat16 *__restrict__ stk)
{
    const int kv = K >> 3;
    const long long total = rows * kv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / kv;
        const int k = (int)(i - r * kv) * 8;
        uint4 hi, lo;
        split8(w + r * ldw + k, hi, lo);
        if (fwd) {
            __nv_bfloat16 *o = fwd + r * 3 * K + k;
            *reinterpret_cast<uint4 *>(o) = hi;
            *reinterpret_cast<uint4 *>(o + K) = hi;
            *reinterpret_cast<uint4 *>(o + 2 * K) = lo;
        }
        if (stk) {
            *reinterpret_cast<uint4 *>(stk + r * K + k) = hi;
            *reinterpret_cast<uint4 *>(stk + (rows + r) * K + k) = hi;
            *reinterpret_cast<uint4 *>(stk + (2 * rows + r) * K + k) = lo;
        }
    }
}

__device__ __forceinline__ float4 max4(float4 a, float4 b)
{
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// 2x2 max-pool, fp32 NHWC; thread = (output pixel, 4 channels)
__global__ void __launch_bounds__(256)
maxpool_f32_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int NB, int H, int W, int C)
{
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 2;
    const long long total = (long long)NB * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % cv);
        long long t = i / cv;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float *p = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + c4 * 4;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p + C)),
                     c = __ldg(reinterpret_cast<const float4 *>(p + (long long)W * C)),
                     d = __ldg(reinterpret_cast<const float4 *>(p + (long long)W * C + C));
        *reinterpret_cast<float4 *>(y + i * 4) = max4(max4(a, b), max4(c, d));
    }
}

// gradient of the above at x: dy goes to the FIRST maximum of the window in (h, w) scan order;
// with relu_mask also through the ReLU that produced x (same rule as nhwc_ops.cu: maxpool_bwd_kernel)
__global__ void __launch_bounds__(256)
maxpool_f32_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dx, int NB,
                       int H, int W, int C, int relu_mask)
{
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 2;
    const long long total = (long long)NB * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % cv);
        long long t = i / cv;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * C + c4 * 4;
        const long long off[4] = {0, C, (long long)W * C, (long long)W * C + C};
        float xv[4][4], ov[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + base + off[k]));
            xv[k][0] = v.x; xv[k][1] = v.y; xv[k][2] = v.z; xv[k][3] = v.w;
        }
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(dy + i * 4));
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int best = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (xv[k][j] > xv[best][j]) best = k;
            const bool live = !relu_mask || xv[best][j] > 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) ov[k][j] = (k == best && live) ? g[j] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4 *>(dx + base + off[k]) = make_float4(ov[k][0], ov[k][1], ov[k][2], ov[k][3]);
    }
}

// NCHW fp32 -> NHWC fp32, channels >= C written as zero up to Cpad (32 x 32 tile through shared memory)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_f32_kernel(const float *__restrict__ x, float *__restrict__ y, int C, long long HW, int Cpad)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        tile[r][tx] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        if (p < HW && c < Cpad) y[((long long)n * HW + p) * Cpad + c] = tile[tx][r];
    }
}

// out[N] += column sums of x fp32 [M, ld]: 8 row lanes x 32 columns per block, one red.add per (block, column)
__global__ void __launch_bounds__(256)
colsum_f32_ld_kernel(const float *__restrict__ x, long long ld, long long M, int N, float *__restrict__ out)
{
    __shared__ float part[8][33];
    const int lane_c = threadIdx.x & 31, lane_r = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane_c;
    const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float acc = 0.f;
    if (c < N)
        for (long long r = r0 + lane_r; r < r1; r += 8) acc += __ldg(x + r * ld + c);
    part[lane_r][lane_c] = acc;
    __syncthreads();
    if (lane_r == 0 && c < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) acc += part[k][lane_c];
        red_add_f32(out + c, acc);
    }
}

int grid_for(long long work_items, int threads)
{
    long long want = (work_items + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

SCDA_API int scda_split3_f32_bf16(long long rows, int C, const float *x, long long ldx, void *y, cudaStream_t stream)
{
    if (rows <= 0 || C <= 0 || !x || !y || C % 8 || ldx % 4 || ldx < C) return 0;
    if (((uintptr_t)x | (uintptr_t)y) % 16) return 0;
    split3_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, stream>>>(x, ldx, rows, C, (__nv_bfloat16 *)y);
    return scda_launch_status();
}

SCDA_API int scda_split_weights_f32_bf16(long long rows, int K, const float *w, long long ldw, void *fwd, void *stk,
                                         cudaStream_t stream)
{
    if (rows <= 0 || K <= 0 || !w || (!fwd && !stk) || K % 8 || ldw % 4 || ldw < K) return 0;
    if (((uintptr_t)w | (uintptr_t)fwd | (uintptr_t)stk) % 16) return 0;
    split_weights_kernel<<<grid_for(rows * (K / 8), 256), 256, 0, stream>>>(w, ldw, rows, K, (__nv_bfloat16 *)fwd,
                                                                            (__nv_bfloat16 *)stk);
    return scda_launch_status();
}

SCDA_API int scda_maxpool2x2_nhwc_f32(int NB, int H, int W, int C, const float *x, float *y, cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || C <= 0 || !x || !y) return 0;
    if ((H | W) & 1 || C % 4 || ((uintptr_t)x | (uintptr_t)y) % 16) return 0;
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 4);
    maxpool_f32_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, y, NB, H, W, C);
    return scda_launch_status();
}

SCDA_API int scda_maxpool2x2_bwd_nhwc_f32(int NB, int H, int W, int C, const float *x, const float *dy, float *dx,
                                          int relu_mask, cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || C <= 0 || !x || !dy || !dx) return 0;
    if ((H | W) & 1 || C % 4 || ((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) % 16) return 0;
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 4);
    maxpool_f32_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, dy, dx, NB, H, W, C, relu_mask);
    return scda_launch_status();
}

SCDA_API int scda_nchw_f32_to_nhwc_f32(int NB, int C, int H, int W, int Cpad, const float *x, float *y,
                                       cudaStream_t stream)
{
    if (NB <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad < C || !x || !y) return 0;
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), NB);
    nchw_to_nhwc_f32_kernel<<<grid, 256, 0, stream>>>(x, y, C, HW, Cpad);
    return scda_launch_status();
}

SCDA_API int scda_colsum_f32_ld(long long M, int N, const float *x, long long ld, float *out, cudaStream_t stream)
{
    if (M <= 0 || N <= 0 || !x || !out || ld < N) return 0;
    const int gx = ceil_div(N, 32);
    long long gy = (M + 255) / 256;
    const long long cap = (long long)kNumSMs * 8 / gx;
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    colsum_f32_ld_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(x, ld, M, N, out);
    return scda_launch_status();
}
