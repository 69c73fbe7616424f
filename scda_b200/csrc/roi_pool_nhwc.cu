// RoI max pooling on the detector's NHWC bf16 feature map, forward and backward.
//
// Same operator as ROIPoolForward / ROIPoolBackward of the reference
// (extensions/_roi_pooling/src/roi_pooling_kernel.cu:24-93, 128-203: round() of the scaled
// RoI, bin = roi / pooled, [floor(ph*bin), ceil((ph+1)*bin)) windows clipped to the map, max
// with the first maximum in row-major order, empty bin -> 0) in the layout the tensor-core
// stages around it use (scda_b200/tc_detector.py):
//   features [NB, H, W, C] bf16  ->  out [R, C, PH, PW] bf16 (the reference's channel-major
//   order, i.e. exactly the [R, C*PH*PW] A operand of fc6) + argmax [R, C, PH, PW] uint16 =
//   h * W + w of the maximum inside the RoI's image (0xFFFF for an empty bin).
// max() of bf16 values is exact, so the result equals the fp32 operator on the same values.
//
// Forward: one CTA per RoI.  A warp owns a bin; its 32 lanes read 32 x 16 B = 512 B of
// consecutive channels per pixel (fully coalesced, the 2 MB map sits in L2) and keep 8
// running maxima each.  Results are transposed through shared memory so that the CTA writes
// its RoI's 2 * C * PH * PW bytes of output (and as many of argmax) as one contiguous stream
// of 16-byte stores.  Compulsory HBM traffic at the model shape (1x32x64x512, 512 RoIs, 7x7):
// 2.1 MB + 25.7 MB + 25.7 MB = 53.5 MB (the fp32 NCHW form moves 107 MB).
// Backward: each (RoI, bin, channel) gradient goes to exactly one feature element: the dout
// and argmax streams are read once (16-byte loads) and scattered with red.global.add.f32
// into the fp32 gradient map (4 MB, L2 resident).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kThreadsRP = 256;
constexpr int kWarpsRP = kThreadsRP / 32;

struct PoolRoiN {
    int batch, x0, y0;
    float bin_h, bin_w;
};

// roi_pooling_kernel.cu:45-56
__device__ __forceinline__ PoolRoiN load_roi(const float *r, float scale, int PH, int PW)
{
    PoolRoiN q;
    q.batch = (int)r[0];
    q.x0 = (int)roundf(r[1] * scale);
    q.y0 = (int)roundf(r[2] * scale);
    const int x1 = (int)roundf(r[3] * scale);
    const int y1 = (int)roundf(r[4] * scale);
    const int rw = (int)fmaxf((float)(x1 - q.x0 + 1), 1.f);
    const int rh = (int)fmaxf((float)(y1 - q.y0 + 1), 1.f);
    q.bin_h = __fdiv_rn((float)rh, (float)PH);
    q.bin_w = __fdiv_rn((float)rw, (float)PW);
    return q;
}

__device__ __forceinline__ int clamp_e(int v, int hi) { return (int)fminf(fmaxf((float)v, 0.f), (float)hi); }

// 8 consecutive channels of one pixel as floats (T = bf16: one 16-byte load; T = float: two)
__device__ __forceinline__ void load8(const __nv_bfloat16 *p, float (&f)[8])
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat16 *e = reinterpret_cast<const __nv_bfloat16 *>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(e[j]);
}
__device__ __forceinline__ void load8(const float *p, float (&f)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store1(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store1(float *p, float v) { *p = v; }
__device__ __forceinline__ float as_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float as_float(float v) { return v; }

// T = __nv_bfloat16: the throughput mode's feature map; T = float: the fp32-parity mode (tc.set_precision)
template <typename T>
__global__ void __launch_bounds__(kThreadsRP)
roi_pool_nhwc_fwd_kernel(const T *__restrict__ feat, float scale, int NB, int H, int W, int C, int PH,
                         int PW, const float *__restrict__ rois, T *__restrict__ out,
                         unsigned short *__restrict__ argmax)
{
    extern __shared__ __align__(16) unsigned char rp_smem[];
    const int bins = PH * PW;
    // [C][bins] values, [C][bins] argmax, then the bin edges
    T *s_val = reinterpret_cast<T *>(rp_smem);
    unsigned short *s_arg = reinterpret_cast<unsigned short *>(s_val + (size_t)C * bins);
    int *hs = reinterpret_cast<int *>(s_arg + (size_t)C * bins + ((C * bins) & 1));
    int *he = hs + PH, *ws = he + PH, *we = ws + PW;
    const int n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PoolRoiN q = load_roi(rois + 5 * n, scale, PH, PW);
    for (int i = threadIdx.x; i < PH + PW; i += kThreadsRP) {
        if (i < PH) {
            hs[i] = clamp_e((int)floorf(__fmul_rn((float)i, q.bin_h)) + q.y0, H);
            he[i] = clamp_e((int)ceilf(__fmul_rn((float)(i + 1), q.bin_h)) + q.y0, H);
        } else {
            const int j = i - PH;
            ws[j] = clamp_e((int)floorf(__fmul_rn((float)j, q.bin_w)) + q.x0, W);
            we[j] = clamp_e((int)ceilf(__fmul_rn((float)(j + 1), q.bin_w)) + q.x0, W);
        }
    }
    __syncthreads();
    const int groups = C >> 3;                       // 8-channel (16-byte) groups
    const int b_ok = q.batch >= 0 && q.batch < NB;
    const T *img = feat + (long long)(b_ok ? q.batch : 0) * H * W * C;
    for (int b = warp; b < bins; b += kWarpsRP) {
        const int ph = b / PW, pw = b - ph * PW;
        const int h0 = hs[ph], h1 = he[ph], w0 = ws[pw], w1 = we[pw];
        const bool empty = h1 <= h0 || w1 <= w0 || !b_ok;
        for (int g = lane; g < groups; g += 32) {
            float best[8];
            int where[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = empty ? 0.f : -3.4e38f; where[j] = 0xFFFF; }
            if (!empty) {
                for (int h = h0; h < h1; ++h) {
                    const T *row = img + ((long long)h * W) * C + g * 8;
                    for (int w = w0; w < w1; ++w) {
                        float f[8];
                        load8(row + (long long)w * C, f);
                        const int pos = h * W + w;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (f[j] > best[j]) { best[j] = f[j]; where[j] = pos; }
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = g * 8 + j;
                store1(s_val + c * bins + b, best[j]);
                s_arg[c * bins + b] = (unsigned short)where[j];
            }
        }
    }
    __syncthreads();
    // contiguous write-out of this RoI's [C * bins] values and argmax
    const long long obase = (long long)n * C * bins;
    const int total = C * bins;
    if ((total & 7) == 0) {
        const uint4 *sv = reinterpret_cast<const uint4 *>(s_val);
        const uint4 *sa = reinterpret_cast<const uint4 *>(s_arg);
        uint4 *ov = reinterpret_cast<uint4 *>(out + obase);
        uint4 *oa = reinterpret_cast<uint4 *>(argmax + obase);
        for (int i = threadIdx.x; i < total / 8; i += kThreadsRP) oa[i] = sa[i];
        for (int i = threadIdx.x; i < total * (int)sizeof(T) / 16; i += kThreadsRP) ov[i] = sv[i];
    } else {
        for (int i = threadIdx.x; i < total; i += kThreadsRP) {
            out[obase + i] = s_val[i];
            argmax[obase + i] = s_arg[i];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreadsRP)
roi_pool_nhwc_bwd_kernel(const T *__restrict__ dout, const unsigned short *__restrict__ argmax,
                         const float *__restrict__ rois, int NB, int HW, int C, int bins,
                         float *__restrict__ dfeat)
{
    // one CTA per RoI: element i of the RoI's [C * bins] block belongs to channel i / bins
    const int n = blockIdx.x;
    const int batch = (int)rois[5 * n];
    if (batch < 0 || batch >= NB) return;
    float *img = dfeat + (long long)batch * HW * C;
    const long long base = (long long)n * C * bins;
    const int total = C * bins;
    if ((total & 7) == 0) {
        const uint4 *av = reinterpret_cast<const uint4 *>(argmax + base);
        for (int i = threadIdx.x; i < total / 8; i += kThreadsRP) {
            const uint4 a = __ldg(av + i);
            float ge[8];
            load8(dout + base + (long long)i * 8, ge);
            const unsigned short *ae = reinterpret_cast<const unsigned short *>(&a);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = i * 8 + j;
                const float gj = ge[j];
                if (ae[j] != 0xFFFF && gj != 0.f) red_add_f32(img + (long long)ae[j] * C + e / bins, gj);
            }
        }
    } else {
        for (int e = threadIdx.x; e < total; e += kThreadsRP) {
            const unsigned short a = argmax[base + e];
            const float gj = as_float(dout[base + e]);
            if (a != 0xFFFF && gj != 0.f) red_add_f32(img + (long long)a * C + e / bins, gj);
        }
    }
}

size_t fwd_smem(int C, int PH, int PW, size_t elem)
{
    const size_t bins = (size_t)PH * PW;
    return (elem + 2) * (size_t)C * bins + 4 + sizeof(int) * 2 * (size_t)(PH + PW) + 16;
}

template <typename T>
int roi_pool_nhwc_fwd(const void *feat, float spatial_scale, int num_rois, int batch, int H, int W, int C, int PH,
                      int PW, const float *rois, void *out, unsigned short *argmax, cudaStream_t stream)
{
    if (num_rois < 0 || batch <= 0 || H <= 0 || W <= 0 || C <= 0 || PH <= 0 || PW <= 0 || !feat || !rois || !out ||
        !argmax)
        return 0;
    if (C % 8 || (long long)H * W >= 0xFFFF) return 0;
    if (((uintptr_t)feat | (uintptr_t)out | (uintptr_t)argmax) % 16) return 0;
    if (num_rois == 0) return 1;
    const size_t smem = fwd_smem(C, PH, PW, sizeof(T));
    if (smem > 220 * 1024) return 0;
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(roi_pool_nhwc_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr_smem = smem;
    }
    roi_pool_nhwc_fwd_kernel<T><<<num_rois, kThreadsRP, smem, stream>>>((const T *)feat, spatial_scale, batch, H, W, C,
                                                                         PH, PW, rois, (T *)out, argmax);
    return scda_launch_status();
}

template <typename T>
int roi_pool_nhwc_bwd(const void *dout, const unsigned short *argmax, const float *rois, int num_rois, int batch,
                      int H, int W, int C, int PH, int PW, float *dfeat, cudaStream_t stream)
{
    if (num_rois < 0 || batch <= 0 || H <= 0 || W <= 0 || C <= 0 || PH <= 0 || PW <= 0 || !dout || !argmax || !rois ||
        !dfeat)
        return 0;
    if (((uintptr_t)dout | (uintptr_t)argmax) % 16) return 0;
    cudaError_t e = cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)batch * H * W * C, stream);
    if (e != cudaSuccess) return -(int)e;
    if (num_rois == 0) return 1;
    roi_pool_nhwc_bwd_kernel<T><<<num_rois, kThreadsRP, 0, stream>>>((const T *)dout, argmax, rois, batch, H * W, C,
                                                                      PH * PW, dfeat);
    return scda_launch_status();
}

}  // namespace

SCDA_API int scda_roi_pool_nhwc_bf16_fwd(const void *feat, float spatial_scale, int num_rois, int batch, int H,
                                         int W, int C, int PH, int PW, const float *rois, void *out,
                                         unsigned short *argmax, cudaStream_t stream)
{
    return roi_pool_nhwc_fwd<__nv_bfloat16>(feat, spatial_scale, num_rois, batch, H, W, C, PH, PW, rois, out, argmax,
                                            stream);
}

SCDA_API int scda_roi_pool_nhwc_bf16_bwd(const void *dout, const unsigned short *argmax, const float *rois,
                                         int num_rois, int batch, int H, int W, int C, int PH, int PW,
                                         float *dfeat, cudaStream_t stream)
{
    return roi_pool_nhwc_bwd<__nv_bfloat16>(dout, argmax, rois, num_rois, batch, H, W, C, PH, PW, dfeat, stream);
}

SCDA_API int scda_roi_pool_nhwc_f32_fwd(const float *feat, float spatial_scale, int num_rois, int batch, int H,
                                        int W, int C, int PH, int PW, const float *rois, float *out,
                                        unsigned short *argmax, cudaStream_t stream)
{
    return roi_pool_nhwc_fwd<float>(feat, spatial_scale, num_rois, batch, H, W, C, PH, PW, rois, out, argmax, stream);
}

SCDA_API int scda_roi_pool_nhwc_f32_bwd(const float *dout, const unsigned short *argmax, const float *rois,
                                        int num_rois, int batch, int H, int W, int C, int PH, int PW,
                                        float *dfeat, cudaStream_t stream)
{
    return roi_pool_nhwc_bwd<float>(dout, argmax, rois, num_rois, batch, H, W, C, PH, PW, dfeat, stream);
}
