// Memory-bound companions of the tensor-core kernels on the detector's bf16 NHWC path:
// layout conversion at the two ends of the backbone, 2x2 max-pool forward / backward
// (backward fused with the ReLU gradient of the layer below), the split-K slab reduction
// of the weight gradients and the bias gradient (column sums).
//
// Replaces nn.MaxPool2d(2, 2) / nn.ReLU of the VGG stack
// (models/faster_rcnn/vgg_adver_expansion_cluster.py:101-114 of the reference) and the
// cuDNN layout transposes around every convolution (nchwToNhwcKernel / nhwcToNchwKernel:
// 525 launches and 13.5 % of the step in profiles/r1_launches_a_step_torchconv_summary.txt).
// All kernels are HBM bound: 16-byte vector accesses, one pass over each tensor.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b)
{
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&b));
    return *reinterpret_cast<uint32_t *>(&r);
}

// ------------------------------------------------------------------ max-pool forward
// thread = (output pixel, 8 channels)
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ y, int NB, int H, int W,
                   int C)
{
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 3;
    const long long total = (long long)NB * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long t = i / cv;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const __nv_bfloat16 *p = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + c8 * 8;
        const uint4 a = ldg16(p), b = ldg16(p + C), c = ldg16(p + (long long)W * C),
                    d = ldg16(p + (long long)W * C + C);
        uint4 m;
        m.x = bf2_max(bf2_max(a.x, b.x), bf2_max(c.x, d.x));
        m.y = bf2_max(bf2_max(a.y, b.y), bf2_max(c.y, d.y));
        m.z = bf2_max(bf2_max(a.z, b.z), bf2_max(c.z, d.z));
        m.w = bf2_max(bf2_max(a.w, b.w), bf2_max(c.w, d.w));
        *reinterpret_cast<uint4 *>(y + i * 8) = m;
    }
}

// ------------------------------------------------------------------ max-pool backward (+ ReLU gradient)
// dx[window] = dy at the FIRST maximum of the window in (h, w) scan order (PyTorch's tie
// rule), 0 elsewhere; with relu_mask the gradient is also 0 where x <= 0 (x is the output
// of the ReLU below the pool, so this is that ReLU's backward).
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ dy,
                   __nv_bfloat16 *__restrict__ dx, int NB, int H, int W, int C, int relu_mask)
{
    const int Ho = H >> 1, Wo = W >> 1, cv = C >> 3;
    const long long total = (long long)NB * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long t = i / cv;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * C + c8 * 8;
        const long long off[4] = {0, C, (long long)W * C, (long long)W * C + C};
        uint4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xv[k] = ldg16(x + base + off[k]);
        const uint4 g = ldg16(dy + i * 8);
        const __nv_bfloat16 *gp = reinterpret_cast<const __nv_bfloat16 *>(&g);
        uint4 ov[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(&xv[k])[j]);
            int best = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k] > v[best]) best = k;
            const bool live = !relu_mask || v[best] > 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                reinterpret_cast<__nv_bfloat16 *>(&ov[k])[j] = (k == best && live) ? gp[j] : __float2bfloat16_rn(0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4 *>(dx + base + off[k]) = ov[k];
    }
}

// ------------------------------------------------------------------ NCHW fp32 -> NHWC bf16 (channel padded)
// 32 x 32 (channel x pixel) tile through shared memory: coalesced fp32 reads along pixels,
// coalesced bf16 writes along channels.  Channels >= C are written as zero up to Cpad.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, int C, long long HW, int Cpad)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        tile[r][tx] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        if (p < HW && c < Cpad) y[((long long)n * HW + p) * Cpad + c] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// NHWC bf16 -> NCHW fp32 (the feature map handed to the fp32 RoI operators)
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, int C, long long HW)
{
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        tile[r][tx] = (p < HW && c < C) ? __bfloat162float(x[((long long)n * HW + p) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[tx][r];
    }
}

// ------------------------------------------------------------------ split-K slab reduction
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const float *__restrict__ slabs, long long slab_stride, int n_slabs, float *__restrict__ dst,
                    long long n, int accumulate)
{
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
         i += (long long)gridDim.x * blockDim.x * 4) {
        float4 acc = accumulate ? *reinterpret_cast<const float4 *>(dst + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < n_slabs; ++s) {
            const float4 v = ld_stream_f4(slabs + (long long)s * slab_stride + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4 *>(dst + i) = acc;
    }
}

// Small outputs, many slabs (conv1_2: 36 864 elements x 49 slabs): with one thread per four elements only 36 blocks
// run and each thread walks 49 dependent-latency loads (15 us).  Here kLanes threads share an element group:
// lane l adds slabs l, l + kLanes, ... (independent loads), the lane sums are added in lane order — a fixed order,
// so the result is still deterministic.  Block = (256 / kLanes) element groups x kLanes slab lanes.
template <int kLanes>
__global__ void __launch_bounds__(256)
reduce_slabs_lanes_kernel(const float *__restrict__ slabs, long long slab_stride, int n_slabs,
                          float *__restrict__ dst, long long n, int accumulate)
{
    constexpr int kElems = 256 / kLanes;
    __shared__ float4 part[kLanes][kElems];
    const int e = threadIdx.x % kElems, l = threadIdx.x / kElems;
    for (long long i0 = (long long)blockIdx.x * kElems * 4; i0 < n; i0 += (long long)gridDim.x * kElems * 4) {
        const long long i = i0 + e * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            int sidx = l;
            for (; sidx + 3 * kLanes < n_slabs; sidx += 4 * kLanes) {
                const float4 v0 = ld_stream_f4(slabs + (long long)sidx * slab_stride + i);
                const float4 v1 = ld_stream_f4(slabs + (long long)(sidx + kLanes) * slab_stride + i);
                const float4 v2 = ld_stream_f4(slabs + (long long)(sidx + 2 * kLanes) * slab_stride + i);
                const float4 v3 = ld_stream_f4(slabs + (long long)(sidx + 3 * kLanes) * slab_stride + i);
                acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
            }
            for (; sidx < n_slabs; sidx += kLanes) {
                const float4 v = ld_stream_f4(slabs + (long long)sidx * slab_stride + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        part[l][e] = acc;
        __syncthreads();
        if (l == 0 && i < n) {
            float4 r = accumulate ? *reinterpret_cast<const float4 *>(dst + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < kLanes; ++k) {
                const float4 v = part[k][e];
                r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
            }
            *reinterpret_cast<float4 *>(dst + i) = r;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ column sums (bias gradient)
// x bf16 [M, ld] -> out fp32 [N] (+=).  Block = 256 threads = 8 row lanes x 32 column
// pairs; grid.x covers column pairs, grid.y cuts the rows; partial sums meet in out[]
// through one red.add per (block, column): the sum order across blocks is not fixed, the
// bias gradient is therefore reproducible to fp32 rounding only (as cuDNN's is).
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16 *__restrict__ x, long long ld, long long M, int N, float *__restrict__ out)
{
    __shared__ float2 part[8][32];
    const int lane_c = threadIdx.x & 31, lane_r = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane_c) * 2;
    const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float2 acc = make_float2(0.f, 0.f);
    if (c < N) {
        for (long long r = r0 + lane_r; r < r1; r += 8) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(x + r * ld + c);
            acc.x += __low2float(v);
            acc.y += __high2float(v);
        }
    }
    part[lane_r][lane_c] = acc;
    __syncthreads();
    if (lane_r == 0 && c < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            acc.x += part[k][lane_c].x;
            acc.y += part[k][lane_c].y;
        }
        red_add_f32(out + c, acc.x);
        if (c + 1 < N) red_add_f32(out + c + 1, acc.y);
    }
}

// The same with 16-byte loads (N, ld multiples of 8, x 16-byte aligned): thread t of a block owns the 8-column
// group t % groups (groups = columns of this block / 8 <= 256) and every lanes_r-th row of the block's row range,
// four loads in flight.  The 4-byte form above keeps ~1 KB per block in flight and streamed conv1_2's 67 MB
// gradient at ~2 TB/s (35 us inside the iteration); this one is bound by HBM.
__global__ void __launch_bounds__(256)
colsum8_kernel(const __nv_bfloat16 *__restrict__ x, long long ld, long long M, int N, float *__restrict__ out)
{
    __shared__ float part[256][9];                   // (+1: the final column walk is conflict free)
    const int c_base = blockIdx.y * 2048;
    const int groups = min(256, (N - c_base) >> 3);
    const int lanes_r = 256 / groups;
    const int g = threadIdx.x % groups, lr = threadIdx.x / groups;
    const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    auto add = [&](const uint4 &q) {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[2 * j] += __low2float(h[j]);
            acc[2 * j + 1] += __high2float(h[j]);
        }
    };
    if (lr < lanes_r) {
        const __nv_bfloat16 *col = x + c_base + g * 8;
        long long r = r0 + lr;
        for (; r + 3ll * lanes_r < r1; r += 4ll * lanes_r) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(col + r * ld));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(col + (r + lanes_r) * ld));
            const uint4 q2 = __ldg(reinterpret_cast<const uint4 *>(col + (r + 2ll * lanes_r) * ld));
            const uint4 q3 = __ldg(reinterpret_cast<const uint4 *>(col + (r + 3ll * lanes_r) * ld));
            add(q0); add(q1); add(q2); add(q3);
        }
        for (; r < r1; r += lanes_r) add(__ldg(reinterpret_cast<const uint4 *>(col + r * ld)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[threadIdx.x][j] = acc[j];
    __syncthreads();
    // thread t < groups * 8: column t of the block, the row lanes added in order
    for (int t = threadIdx.x; t < groups * 8; t += 256) {
        const int gg = t >> 3, j = t & 7;
        float s = 0.f;
        for (int k = 0; k < lanes_r; ++k) s += part[k * groups + gg][j];
        red_add_f32(out + c_base + t, s);
    }
}

// Column sums of a contiguous fp32 [M, N] matrix (N % 4 == 0, N / 4 a power of two <= 256): the
// bias gradient of a channels-last convolution, dB[c] = sum over pixels of dY[pixel, c].  Thread
// t of a block owns column group t % (N/4) and every (256 / (N/4))-th row of the block's slab:
// 16-byte loads, a shared-memory tree over the row lanes, one red.add per (block, column).
// (torch's sum over (0, 2, 3) of a channels_last tensor runs at ~1.5 TB/s; this streams at HBM rate.)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float *__restrict__ x, long long M, int N, float *__restrict__ out)
{
    __shared__ float4 part[256];
    const int groups = N >> 2;                       // float4 column groups
    const int lanes_r = 256 / groups;
    const int g = threadIdx.x % groups, lr = threadIdx.x / groups;
    const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (long long r = r0 + lr; r < r1; r += lanes_r) {
        const float4 v = __ldg(x4 + r * groups + g);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = lanes_r >> 1; s > 0; s >>= 1) {
        if (lr < s) {
            const float4 o = part[threadIdx.x + s * groups];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            part[threadIdx.x] = acc;
        }
        __syncthreads();
    }
    if (lr == 0) {
        red_add_f32(out + 4 * g, acc.x);
        red_add_f32(out + 4 * g + 1, acc.y);
        red_add_f32(out + 4 * g + 2, acc.z);
        red_add_f32(out + 4 * g + 3, acc.w);
    }
}

int grid_for(long long work_items, int threads)
{
    long long want = (work_items + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

// dst[c][r] = src[r][c] for a bf16 matrix (leading dimensions in elements): 32 x 32 tiles through shared memory,
// coalesced both ways.  (torch's strided copy took 90 us for the 512 x 4096 gradient blocks whose transposes
// feed the weight-gradient GEMMs of fc6 / fc7 — on the chain that the backbone's backward waits for.)
namespace {
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ src, long long lds, __nv_bfloat16 *__restrict__ dst,
                      long long ldd, int rows, int cols)
{
    __shared__ __nv_bfloat16 tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * k][tx] = src[(long long)r * lds + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < rows && c < cols) dst[(long long)c * ldd + r] = tile[tx][ty + 8 * k];
    }
}
}  // namespace

SCDA_API int scda_transpose_bf16(int rows, int cols, const void *src, long long lds, void *dst, long long ldd,
                                 cudaStream_t stream)
{
    if (rows <= 0 || cols <= 0 || !src || !dst || lds < cols || ldd < rows) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_bf16_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)src, lds, (__nv_bfloat16 *)dst, ldd, rows,
                                                   cols);
    return scda_launch_status();
}

SCDA_API int scda_maxpool2x2_nhwc_bf16(int NB, int H, int W, int C, const void *x, void *y, cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || C <= 0 || !x || !y) return 0;
    if ((H | W) & 1 || C % 8) return 0;
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 8);
    maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, NB,
                                                                 H, W, C);
    return scda_launch_status();
}

SCDA_API int scda_maxpool2x2_bwd_nhwc_bf16(int NB, int H, int W, int C, const void *x, const void *dy, void *dx,
                                           int relu_mask, cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || C <= 0 || !x || !dy || !dx) return 0;
    if ((H | W) & 1 || C % 8) return 0;
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 8);
    maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __nv_bfloat16 *)x,
                                                                 (const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dx,
                                                                 NB, H, W, C, relu_mask);
    return scda_launch_status();
}

SCDA_API int scda_nchw_f32_to_nhwc_bf16(int NB, int C, int H, int W, int Cpad, const float *x, void *y,
                                        cudaStream_t stream)
{
    if (NB <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad < C || !x || !y) return 0;
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), NB);
    nchw_to_nhwc_kernel<<<grid, 256, 0, stream>>>(x, (__nv_bfloat16 *)y, C, HW, Cpad);
    return scda_launch_status();
}

SCDA_API int scda_nhwc_bf16_to_nchw_f32(int NB, int C, int H, int W, const void *x, float *y, cudaStream_t stream)
{
    if (NB <= 0 || C <= 0 || H <= 0 || W <= 0 || !x || !y) return 0;
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), NB);
    nhwc_to_nchw_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)x, y, C, HW);
    return scda_launch_status();
}

SCDA_API int scda_reduce_slabs_f32(const float *slabs, long long slab_stride, int n_slabs, float *dst,
                                   long long n, int accumulate, cudaStream_t stream)
{
    if (!slabs || !dst || n <= 0 || n_slabs < 1) return 0;
    if (n % 4 || slab_stride % 4 || ((uintptr_t)slabs | (uintptr_t)dst) % 16) return 0;
    // enough threads for ~2 blocks per SM: split the slab walk over 2 / 4 / 8 lanes when the output alone is too small
    const long long want = (long long)kNumSMs * 2 * 256;
    int lanes = 1;
    while (lanes < 8 && (n / 4) * lanes < want && lanes * 2 <= n_slabs) lanes *= 2;
    if (lanes == 8)
        reduce_slabs_lanes_kernel<8><<<grid_for(n / 4, 32), 256, 0, stream>>>(slabs, slab_stride, n_slabs, dst, n, accumulate);
    else if (lanes == 4)
        reduce_slabs_lanes_kernel<4><<<grid_for(n / 4, 64), 256, 0, stream>>>(slabs, slab_stride, n_slabs, dst, n, accumulate);
    else if (lanes == 2)
        reduce_slabs_lanes_kernel<2><<<grid_for(n / 4, 128), 256, 0, stream>>>(slabs, slab_stride, n_slabs, dst, n, accumulate);
    else
        reduce_slabs_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(slabs, slab_stride, n_slabs, dst, n, accumulate);
    return scda_launch_status();
}

SCDA_API int scda_colsum_bf16(long long M, int N, const void *x, long long ld, float *out, cudaStream_t stream)
{
    // out[N] += column sums of x[M, ld] (bf16); N and ld even
    if (M <= 0 || N <= 0 || !x || !out || (N & 1) || (ld & 1) || ((uintptr_t)x % 4)) return 0;
    if (N % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x % 16) == 0 && (N <= 2048 || N % 2048 == 0)) {
        const int cols = N < 2048 ? N : 2048, groups = cols / 8, lanes_r = 256 / groups;
        long long blocks = (M + 8ll * lanes_r - 1) / (8ll * lanes_r);          // >= 8 rows per thread
        const long long cap = (long long)kNumSMs * 4 / ceil_div(N, 2048);
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        colsum8_kernel<<<dim3((unsigned)blocks, ceil_div(N, 2048)), 256, 0, stream>>>((const __nv_bfloat16 *)x, ld, M, N,
                                                                                    out);
        return scda_launch_status();
    }
    const int gx = ceil_div(N, 64);
    long long gy = (M + 255) / 256;
    const long long cap = (long long)kNumSMs * 8 / gx;
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    colsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>((const __nv_bfloat16 *)x, ld, M, N, out);
    return scda_launch_status();
}

SCDA_API int scda_colsum_f32(long long M, int N, const float *x, float *out, cudaStream_t stream)
{
    // out[N] += column sums of the contiguous fp32 matrix x[M, N]
    if (M <= 0 || N <= 0 || !x || !out || (N & 3) || ((uintptr_t)x % 16)) return 0;
    const int groups = N >> 2;
    if (groups > 256 || (groups & (groups - 1))) return 0;
    const int lanes_r = 256 / groups;
    long long blocks = (M + 4LL * lanes_r - 1) / (4LL * lanes_r);        // >= 4 rows per thread
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    colsum_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, M, N, out);
    return scda_launch_status();
}
