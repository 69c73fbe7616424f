// RoIAlign (one bilinear sample per grid point), forward and backward, sm_100a.
//
// Replaces ROIAlignForward / ROIAlignBackward of the reference
// (extensions/_roi_align/src/roi_align_kernel.cu:15-70, 94-143).  This is the
// old single-sample variant: roi_extent = (end - start) * scale + 1,
// bin = extent / (aligned - 1), sample at start + p * bin; zero outside the
// map.  The reference's mixed float/double expression is reproduced operation
// by operation (the order nvcc 12.9 emits for the unmodified source), so the
// forward agrees to the last bit; backward differs only in the order the
// atomic additions land.
//
// One CTA per (RoI, 64-channel chunk): the AH*AW sample taps (offset, the two
// fractions, inside flag) depend only on the RoI, so they are computed once
// into shared memory and reused by every channel of the chunk.  Each thread
// produces 4 consecutive outputs and issues one 128-bit streaming store.
#include "common.cuh"

namespace {

constexpr int kAlignThreads = 256;
constexpr int kAlignChunkC = 64;

struct Tap {
    int ul;      // offset of the up-left tap in a channel plane, -1 = outside
    float hr, wr;
};

// roi_align_kernel.cu:33-53
__device__ __forceinline__ Tap make_tap(const float *r, float scale, int H, int W, int AH, int AW,
                                        int ph, int pw)
{
    const float x0 = __fmul_rn(r[1], scale), y0 = __fmul_rn(r[2], scale);
    const float x1 = __fmul_rn(r[3], scale), y1 = __fmul_rn(r[4], scale);
    const float rw = fmaxf(__fadd_rn(__fsub_rn(x1, x0), 1.f), 0.f);
    const float rh = fmaxf(__fadd_rn(__fsub_rn(y1, y0), 1.f), 0.f);
    const float bin_h = (float)__ddiv_rn((double)rh, (double)AH - 1.0);
    const float bin_w = (float)__ddiv_rn((double)rw, (double)AW - 1.0);
    const float h = __fmaf_rn((float)ph, bin_h, y0);
    const float w = __fmaf_rn((float)pw, bin_w, x0);
    Tap t;
    const int hs = (int)fminf(floorf(h), (float)(H - 2));
    const int ws = (int)fminf(floorf(w), (float)(W - 2));
    t.hr = __fsub_rn(h, (float)hs);
    t.wr = __fsub_rn(w, (float)ws);
    t.ul = (h < 0 || h >= H || w < 0 || w >= W) ? -1 : hs * W + ws;
    return t;
}

__device__ __forceinline__ int image_offset(const float *r, int C, int H, int W)
{
    // roi_batch_ind is a float in the reference; the product is formed in float
    return (int)__fmul_rn(__fmul_rn(__fmul_rn(r[0], (float)C), (float)H), (float)W);
}

template <bool kVec>
__global__ void __launch_bounds__(kAlignThreads)
roi_align_fwd_kernel(const float *__restrict__ feat, float scale, int H, int W, int C, int AH,
                     int AW, const float *__restrict__ rois, float *__restrict__ out)
{
    extern __shared__ Tap s_tap[];
    const int n = blockIdx.x, c0 = blockIdx.y * kAlignChunkC;
    const float *r = rois + 5 * n;
    const int bins = AH * AW;
    for (int b = threadIdx.x; b < bins; b += kAlignThreads)
        s_tap[b] = make_tap(r, scale, H, W, AH, AW, b / AW, b % AW);
    __syncthreads();

    const int cn = min(kAlignChunkC, C - c0);
    const int total = cn * bins;
    const long long obase = ((long long)n * C + c0) * bins;
    const float *__restrict__ img = feat + image_offset(r, C, H, W) + (long long)c0 * H * W;
    constexpr int kPer = kVec ? 4 : 1;

    for (int e0 = threadIdx.x * kPer; e0 < total; e0 += kAlignThreads * kPer) {
        float val[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int e = e0 + k;
            const int c = e / bins, b = e - c * bins;
            const Tap t = s_tap[b];
            float v = 0.f;
            if (t.ul >= 0) {
                const float *p = img + c * H * W + t.ul;
                const float ul = __ldg(p), ur = __ldg(p + 1), dl = __ldg(p + W), dr = __ldg(p + W + 1);
                const double omh = __dsub_rn(1.0, (double)t.hr), omw = __dsub_rn(1.0, (double)t.wr);
                double acc = __fma_rn(omw, __dmul_rn(omh, (double)ul),
                                      __dmul_rn(__dmul_rn(omh, (double)ur), (double)t.wr));
                acc = __fma_rn(omw, (double)__fmul_rn(t.hr, dl), acc);
                acc = __dadd_rn(acc, (double)__fmul_rn(t.wr, __fmul_rn(t.hr, dr)));
                v = (float)acc;
            }
            val[k] = v;
        }
        if (kVec)
            st_stream_f4(out + obase + e0, make_float4(val[0], val[1 % kPer], val[2 % kPer], val[3 % kPer]));
        else
            out[obase + e0] = val[0];
    }
}

template <bool kVec>
__global__ void __launch_bounds__(kAlignThreads)
roi_align_bwd_kernel(const float *__restrict__ top_diff, float scale, int H, int W, int C, int AH,
                     int AW, const float *__restrict__ rois, float *__restrict__ bottom_diff)
{
    extern __shared__ Tap s_tap[];
    const int n = blockIdx.x, c0 = blockIdx.y * kAlignChunkC;
    const float *r = rois + 5 * n;
    const int bins = AH * AW;
    for (int b = threadIdx.x; b < bins; b += kAlignThreads)
        s_tap[b] = make_tap(r, scale, H, W, AH, AW, b / AW, b % AW);
    __syncthreads();

    const int cn = min(kAlignChunkC, C - c0);
    const int total = cn * bins;
    const long long obase = ((long long)n * C + c0) * bins;
    float *__restrict__ img = bottom_diff + image_offset(r, C, H, W) + (long long)c0 * H * W;
    constexpr int kPer = kVec ? 4 : 1;

    for (int e0 = threadIdx.x * kPer; e0 < total; e0 += kAlignThreads * kPer) {
        float g[kPer];
        if (kVec) {
            const float4 v = ld_stream_f4(top_diff + obase + e0);
            g[0] = v.x; g[1 % kPer] = v.y; g[2 % kPer] = v.z; g[3 % kPer] = v.w;
        } else {
            g[0] = top_diff[obase + e0];
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int e = e0 + k;
            const int c = e / bins, b = e - c * bins;
            const Tap t = s_tap[b];
            if (t.ul < 0) continue;
            float *p = img + c * H * W + t.ul;
            // roi_align_kernel.cu:136-139: the first two products are formed in
            // double ((1. - h_ratio)), the last two in float
            const double gh = __dmul_rn((double)g[k], __dsub_rn(1.0, (double)t.hr));
            const float omw = __fsub_rn(1.f, t.wr);
            const float gl = __fmul_rn(g[k], t.hr);
            red_add_f32(p, (float)__dmul_rn(gh, (double)omw));
            red_add_f32(p + 1, (float)__dmul_rn(gh, (double)t.wr));
            red_add_f32(p + W, __fmul_rn(gl, omw));
            red_add_f32(p + W + 1, __fmul_rn(gl, t.wr));
        }
    }
}

bool align_args_ok(int num_rois, int H, int W, int C, int AH, int AW, const void *a, const void *b,
                   const void *c)
{
    return num_rois >= 0 && H >= 2 && W >= 2 && C > 0 && AH >= 2 && AW >= 2 && a && b && c;
}

}  // namespace

SCDA_API int ROIAlignForwardLaucher(const float *bottom_data, const float spatial_scale,
                                    const int num_rois, const int height, const int width,
                                    const int channels, const int aligned_height,
                                    const int aligned_width, const float *bottom_rois,
                                    float *top_data, cudaStream_t stream)
{
    if (!align_args_ok(num_rois, height, width, channels, aligned_height, aligned_width,
                       bottom_data, bottom_rois, top_data))
        return 0;
    if (num_rois == 0) return 1;
    dim3 grid(num_rois, ceil_div(channels, kAlignChunkC));
    const size_t smem = sizeof(Tap) * aligned_height * aligned_width;
    if (smem > 48 * 1024) return 0;
    const bool vec = channels % 4 == 0 && (uintptr_t)top_data % 16 == 0;
    if (vec)
        roi_align_fwd_kernel<true><<<grid, kAlignThreads, smem, stream>>>(
            bottom_data, spatial_scale, height, width, channels, aligned_height, aligned_width,
            bottom_rois, top_data);
    else
        roi_align_fwd_kernel<false><<<grid, kAlignThreads, smem, stream>>>(
            bottom_data, spatial_scale, height, width, channels, aligned_height, aligned_width,
            bottom_rois, top_data);
    return scda_launch_status();
}

SCDA_API int ROIAlignBackwardLaucher(const float *top_diff, const float spatial_scale,
                                     const int batch_size, const int num_rois, const int height,
                                     const int width, const int channels,
                                     const int aligned_height, const int aligned_width,
                                     const float *bottom_rois, float *bottom_diff,
                                     cudaStream_t stream)
{
    (void)batch_size;
    if (!align_args_ok(num_rois, height, width, channels, aligned_height, aligned_width, top_diff,
                       bottom_rois, bottom_diff))
        return 0;
    if (num_rois == 0) return 1;
    dim3 grid(num_rois, ceil_div(channels, kAlignChunkC));
    const size_t smem = sizeof(Tap) * aligned_height * aligned_width;
    if (smem > 48 * 1024) return 0;
    const bool vec = channels % 4 == 0 && (uintptr_t)top_diff % 16 == 0;
    if (vec)
        roi_align_bwd_kernel<true><<<grid, kAlignThreads, smem, stream>>>(
            top_diff, spatial_scale, height, width, channels, aligned_height, aligned_width,
            bottom_rois, bottom_diff);
    else
        roi_align_bwd_kernel<false><<<grid, kAlignThreads, smem, stream>>>(
            top_diff, spatial_scale, height, width, channels, aligned_height, aligned_width,
            bottom_rois, bottom_diff);
    return scda_launch_status();
}
