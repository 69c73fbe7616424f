// InstanceNorm2d (+ ReLU / LeakyReLU) forward and backward on channels-last fp32 tensors.
//
// Replaces nn.InstanceNorm2d(affine=False, eps=1e-5) and the activation that follows it in
// the reconstruction networks (models/faster_rcnn/common_net.py:59-80 INSResBlock,
// :279-293 LeakyReLUConvTranspose2d_2 of the reference).  torch evaluates instance norm as a
// batch norm over a [1, N*C, H, W] NCHW view, which forces an NCHW copy in front of it and an
// NHWC copy behind it for every cuDNN tensor-core convolution around it (2.5 ms of layout
// transposes + 1.6 ms of batch-norm kernels per iteration in
// profiles/r1_launches_c_step_tc_summary.txt).  Here the data stays [N, H*W, C] (C innermost):
//   forward : per-(n, c) shifted sums in two deterministic stages -> mean, rstd;
//             y = act((x - mean) * rstd)                       (x read twice, y written once)
//   backward: g = dy * act'(xhat); per-(n, c) sums of g and g * xhat (two stages);
//             dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// All kernels are HBM bound (float4 accesses, one pass per tensor per stage).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kNT = 256;

// 4 consecutive elements (index in units of 4) of an fp32 or bf16 array as a float4
template <typename T>
__device__ __forceinline__ float4 ld4(const T *p, long long i4);
template <>
__device__ __forceinline__ float4 ld4<float>(const float *p, long long i4) { return ld_stream_f4(p + i4 * 4); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p, long long i4)
{
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p) + i4);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&q.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&q.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <typename T>
__device__ __forceinline__ void st4(T *p, long long i4, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float *p, long long i4, float4 v)
{
    *reinterpret_cast<float4 *>(p + i4 * 4) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16 *p, long long i4, float4 v)
{
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 q;
    q.x = *reinterpret_cast<const uint32_t *>(&a);
    q.y = *reinterpret_cast<const uint32_t *>(&b);
    *(reinterpret_cast<uint2 *>(p) + i4) = q;
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope)
{
    if (act == 1) return v > 0.f ? v : 0.f;
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}
__device__ __forceinline__ float act_grad(float xhat, int act, float slope)
{
    if (act == 1) return xhat > 0.f ? 1.f : 0.f;
    if (act == 2) return xhat > 0.f ? 1.f : slope;
    return 1.f;
}

// stage 1 of a per-(n, c) reduction over HW: block (chunk, n) accumulates two quantities per
// channel over its rows and writes them to part[n][chunk][2][C]
//   mode 0: (x - shift), (x - shift)^2           shift = x[n, 0, c]
//   mode 1: g, g * xhat                           g = dy * act'(xhat)
template <int kMode, typename TDy>
__global__ void __launch_bounds__(kNT)
in_partial_kernel(const float *__restrict__ x, const TDy *__restrict__ dy, const float *__restrict__ mean,
                  const float *__restrict__ rstd, float *__restrict__ part, int HW, int C, int chunks, int act,
                  float slope)
{
    __shared__ float4 sh[2][kNT];
    const int vec = C >> 2;                 // float4 lanes per row
    const int rows_per_pass = kNT / vec;
    const int lane = threadIdx.x % vec, rlane = threadIdx.x / vec;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int rows_per_chunk = (HW + chunks - 1) / chunks;
    const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    const float *xb = x + (long long)n * HW * C;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    float4 p0, p1;
    if (kMode == 0) {
        p0 = *reinterpret_cast<const float4 *>(xb + lane * 4);     // shift
    } else {
        p0 = *reinterpret_cast<const float4 *>(mean + (long long)n * C + lane * 4);
        p1 = *reinterpret_cast<const float4 *>(rstd + (long long)n * C + lane * 4);
    }
    if (rlane < rows_per_pass) {
        for (int r = r0 + rlane; r < r1; r += rows_per_pass) {
            const float4 v = ld_stream_f4(xb + (long long)r * C + lane * 4);
            if (kMode == 0) {
                const float d0 = v.x - p0.x, d1 = v.y - p0.y, d2 = v.z - p0.z, d3 = v.w - p0.w;
                a.x += d0; a.y += d1; a.z += d2; a.w += d3;
                b.x += d0 * d0; b.y += d1 * d1; b.z += d2 * d2; b.w += d3 * d3;
            } else {
                const float4 g4 = ld4<TDy>(dy, ((long long)n * HW + r) * vec + lane);
                const float h0 = (v.x - p0.x) * p1.x, h1 = (v.y - p0.y) * p1.y, h2 = (v.z - p0.z) * p1.z,
                            h3 = (v.w - p0.w) * p1.w;
                const float g0 = g4.x * act_grad(h0, act, slope), g1 = g4.y * act_grad(h1, act, slope),
                            g2 = g4.z * act_grad(h2, act, slope), g3 = g4.w * act_grad(h3, act, slope);
                a.x += g0; a.y += g1; a.z += g2; a.w += g3;
                b.x += g0 * h0; b.y += g1 * h1; b.z += g2 * h2; b.w += g3 * h3;
            }
        }
    }
    sh[0][threadIdx.x] = a;
    sh[1][threadIdx.x] = b;
    __syncthreads();
    if (rlane == 0) {
        for (int k = 1; k < rows_per_pass; ++k) {
            const float4 u = sh[0][k * vec + lane], w = sh[1][k * vec + lane];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
        }
        float *o = part + (((long long)n * chunks + chunk) * 2) * C + lane * 4;
        *reinterpret_cast<float4 *>(o) = a;
        *reinterpret_cast<float4 *>(o + C) = b;
    }
}

// stage 2: sum over the chunks in a fixed order.  Block = 32 channels x 8 chunk lanes: lane k
// adds chunks k, k+8, ... (independent loads in flight), the 8 partial sums are then added in
// lane order.  mode 0 -> mean, rstd; mode 1 -> mean(g), mean(g xhat)
template <int kMode>
__global__ void __launch_bounds__(256)
in_final_kernel(const float *__restrict__ part, const float *__restrict__ x, float *__restrict__ o0,
                float *__restrict__ o1, int HW, int C, int chunks, float eps)
{
    __shared__ float sa[8][32], sb[8][32];
    const int n = blockIdx.y, cl = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float a = 0.f, b = 0.f;
    if (c < C) {
        for (int k = kl; k < chunks; k += 8) {
            const float *p = part + (((long long)n * chunks + k) * 2) * C + c;
            a += p[0];
            b += p[C];
        }
    }
    sa[kl][cl] = a;
    sb[kl][cl] = b;
    __syncthreads();
    if (kl != 0 || c >= C) return;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        a += sa[k][cl];
        b += sb[k][cl];
    }
    const float inv = 1.f / (float)HW;
    if (kMode == 0) {
        const float shift = x[(long long)n * HW * C + c];
        const float m = a * inv;
        const float var = fmaxf(b * inv - m * m, 0.f);
        o0[(long long)n * C + c] = shift + m;
        o1[(long long)n * C + c] = rsqrtf(var + eps);
    } else {
        o0[(long long)n * C + c] = a * inv;
        o1[(long long)n * C + c] = b * inv;
    }
}

template <typename TY>
__global__ void __launch_bounds__(kNT)
in_apply_kernel(const float *__restrict__ x, const float *__restrict__ mean, const float *__restrict__ rstd,
                TY *__restrict__ y, long long total4, int HW, int C, int act, float slope)
{
    const int vec = C >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i % vec);
        const long long row = i / vec;
        const int n = (int)(row / HW);
        const float4 v = ld_stream_f4(x + i * 4);
        const float4 m = *reinterpret_cast<const float4 *>(mean + (long long)n * C + lane * 4);
        const float4 s = *reinterpret_cast<const float4 *>(rstd + (long long)n * C + lane * 4);
        float4 o;
        o.x = act_fwd((v.x - m.x) * s.x, act, slope);
        o.y = act_fwd((v.y - m.y) * s.y, act, slope);
        o.z = act_fwd((v.z - m.z) * s.z, act, slope);
        o.w = act_fwd((v.w - m.w) * s.w, act, slope);
        st4<TY>(y, i, o);
    }
}

template <typename TDy, typename TDx>
__global__ void __launch_bounds__(kNT)
in_bwd_apply_kernel(const float *__restrict__ x, const TDy *__restrict__ dy, const float *__restrict__ mean,
                    const float *__restrict__ rstd, const float *__restrict__ mg, const float *__restrict__ mgx,
                    TDx *__restrict__ dx, long long total4, int HW, int C, int act, float slope)
{
    const int vec = C >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i % vec);
        const long long row = i / vec;
        const int n = (int)(row / HW);
        const long long o = (long long)n * C + lane * 4;
        const float4 v = ld_stream_f4(x + i * 4), g4 = ld4<TDy>(dy, i);
        const float4 m = *reinterpret_cast<const float4 *>(mean + o), s = *reinterpret_cast<const float4 *>(rstd + o);
        const float4 a = *reinterpret_cast<const float4 *>(mg + o), b = *reinterpret_cast<const float4 *>(mgx + o);
        const float h0 = (v.x - m.x) * s.x, h1 = (v.y - m.y) * s.y, h2 = (v.z - m.z) * s.z, h3 = (v.w - m.w) * s.w;
        float4 r;
        r.x = s.x * (g4.x * act_grad(h0, act, slope) - a.x - h0 * b.x);
        r.y = s.y * (g4.y * act_grad(h1, act, slope) - a.y - h1 * b.y);
        r.z = s.z * (g4.z * act_grad(h2, act, slope) - a.z - h2 * b.z);
        r.w = s.w * (g4.w * act_grad(h3, act, slope) - a.w - h3 * b.w);
        st4<TDx>(dx, i, r);
    }
}

int pick_chunks(int N, int HW, int C)
{
    const int rows_per_pass = kNT / (C >> 2);
    int chunks = (kNumSMs * 4 + N - 1) / N;
    const int max_chunks = (HW + rows_per_pass - 1) / rows_per_pass;
    if (chunks > max_chunks) chunks = max_chunks;
    return chunks < 1 ? 1 : chunks;
}

bool shape_ok(int N, int HW, int C) { return N > 0 && HW > 0 && C >= 4 && C % 4 == 0 && (C >> 2) <= kNT && kNT % (C >> 2) == 0; }

int apply_grid(long long total4)
{
    long long want = (total4 + kNT - 1) / kNT;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(want > cap ? cap : (want < 1 ? 1 : want));
}

}  // namespace

SCDA_API size_t scda_instnorm_workspace_bytes(int N, int HW, int C)
{
    if (!shape_ok(N, HW, C)) return 0;
    return sizeof(float) * (size_t)N * pick_chunks(N, HW, C) * 2 * C;
}

SCDA_API int scda_instnorm_act_fwd_nhwc(int N, int HW, int C, const float *x, void *y, int y_dtype, float *mean,
                                        float *rstd, float eps, int act, float slope, void *workspace,
                                        size_t workspace_bytes, cudaStream_t stream)
{
    if (!shape_ok(N, HW, C) || !x || !y || !mean || !rstd || !workspace || act < 0 || act > 2) return 0;
    if (y_dtype != 0 && y_dtype != 1) return 0;
    if (workspace_bytes < scda_instnorm_workspace_bytes(N, HW, C)) return 0;
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)workspace) % 16) return 0;
    const int chunks = pick_chunks(N, HW, C);
    float *part = (float *)workspace;
    in_partial_kernel<0, float><<<dim3(chunks, N), kNT, 0, stream>>>(x, nullptr, nullptr, nullptr, part, HW, C,
                                                                     chunks, act, slope);
    in_final_kernel<0><<<dim3((C + 31) / 32, N), 256, 0, stream>>>(part, x, mean, rstd, HW, C, chunks, eps);
    const long long total4 = (long long)N * HW * (C >> 2);
    if (y_dtype == 0)
        in_apply_kernel<float><<<apply_grid(total4), kNT, 0, stream>>>(x, mean, rstd, (float *)y, total4, HW, C, act,
                                                                       slope);
    else
        in_apply_kernel<__nv_bfloat16><<<apply_grid(total4), kNT, 0, stream>>>(x, mean, rstd, (__nv_bfloat16 *)y,
                                                                               total4, HW, C, act, slope);
    return scda_launch_status();
}

SCDA_API int scda_instnorm_act_fwd_nhwc_f32(int N, int HW, int C, const float *x, float *y, float *mean, float *rstd,
                                            float eps, int act, float slope, void *workspace,
                                            size_t workspace_bytes, cudaStream_t stream)
{
    return scda_instnorm_act_fwd_nhwc(N, HW, C, x, y, 0, mean, rstd, eps, act, slope, workspace, workspace_bytes,
                                      stream);
}

namespace {
template <typename TDy>
int instnorm_bwd(int N, int HW, int C, const float *x, const TDy *dy, const float *mean, const float *rstd, void *dx,
                 int dx_dtype, int act, float slope, float *part, float *mg, float *mgx, int chunks,
                 cudaStream_t stream)
{
    in_partial_kernel<1, TDy><<<dim3(chunks, N), kNT, 0, stream>>>(x, dy, mean, rstd, part, HW, C, chunks, act, slope);
    in_final_kernel<1><<<dim3((C + 31) / 32, N), 256, 0, stream>>>(part, x, mg, mgx, HW, C, chunks, 0.f);
    const long long total4 = (long long)N * HW * (C >> 2);
    if (dx_dtype == 0)
        in_bwd_apply_kernel<TDy, float><<<apply_grid(total4), kNT, 0, stream>>>(x, dy, mean, rstd, mg, mgx, (float *)dx,
                                                                                total4, HW, C, act, slope);
    else
        in_bwd_apply_kernel<TDy, __nv_bfloat16><<<apply_grid(total4), kNT, 0, stream>>>(
            x, dy, mean, rstd, mg, mgx, (__nv_bfloat16 *)dx, total4, HW, C, act, slope);
    return scda_launch_status();
}
}  // namespace

SCDA_API int scda_instnorm_act_bwd_nhwc(int N, int HW, int C, const float *x, const void *dy, int dy_dtype,
                                        const float *mean, const float *rstd, void *dx, int dx_dtype, int act,
                                        float slope, void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    if (!shape_ok(N, HW, C) || !x || !dy || !dx || !mean || !rstd || !workspace || act < 0 || act > 2) return 0;
    if ((dy_dtype != 0 && dy_dtype != 1) || (dx_dtype != 0 && dx_dtype != 1)) return 0;
    const size_t need = scda_instnorm_workspace_bytes(N, HW, C) + sizeof(float) * 2 * (size_t)N * C;
    if (workspace_bytes < need) return 0;
    if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)workspace) % 16)
        return 0;
    const int chunks = pick_chunks(N, HW, C);
    float *part = (float *)workspace;
    float *mg = part + (size_t)N * chunks * 2 * C, *mgx = mg + (size_t)N * C;
    if (dy_dtype == 0)
        return instnorm_bwd<float>(N, HW, C, x, (const float *)dy, mean, rstd, dx, dx_dtype, act, slope, part, mg, mgx,
                                   chunks, stream);
    return instnorm_bwd<__nv_bfloat16>(N, HW, C, x, (const __nv_bfloat16 *)dy, mean, rstd, dx, dx_dtype, act, slope,
                                       part, mg, mgx, chunks, stream);
}

SCDA_API int scda_instnorm_act_bwd_nhwc_f32(int N, int HW, int C, const float *x, const float *dy, const float *mean,
                                            const float *rstd, float *dx, int act, float slope, void *workspace,
                                            size_t workspace_bytes, cudaStream_t stream)
{
    return scda_instnorm_act_bwd_nhwc(N, HW, C, x, dy, 0, mean, rstd, dx, 0, act, slope, workspace, workspace_bytes,
                                      stream);
}

// ------------------------------------------------------------------------------------------
// Bilinear x2 up-sampling, align_corners=True, channels-last fp32 (the `Interpolate` in front
// of the two decoder convolutions, models/faster_rcnn/common_net.py:160-169, 279-293).
// Index arithmetic = PyTorch's upsample_bilinear2d: r = (in-1)/(out-1) in fp32, src = r * dst,
// i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1.
// Forward: thread = (output pixel, float4 of channels).  Backward is the transpose written as
// a GATHER (thread = input pixel x float4; it enumerates the few output rows / columns whose
// taps land on it with the same fp32 index arithmetic) — no atomics, deterministic.
namespace {

__device__ __forceinline__ void bil_coord(float r, int dst, int in, int &i0, int &i1, float &l0, float &l1)
{
    const float src = r * (float)dst;
    i0 = (int)src;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

template <typename TY>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float *__restrict__ x, TY *__restrict__ y, int N, int H, int W, int C, float rh,
                      float rw)
{
    const int Ho = H * 2, Wo = W * 2, vec = C >> 2;
    const long long total = (long long)N * Ho * Wo * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i % vec);
        long long t = i / vec;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        int h0, h1, w0, w1;
        float a0, a1, b0, b1;
        bil_coord(rh, ho, H, h0, h1, a0, a1);
        bil_coord(rw, wo, W, w0, w1, b0, b1);
        const float *base = x + (long long)n * H * W * C + lane * 4;
        const float4 p00 = __ldg(reinterpret_cast<const float4 *>(base + ((long long)h0 * W + w0) * C));
        const float4 p01 = __ldg(reinterpret_cast<const float4 *>(base + ((long long)h0 * W + w1) * C));
        const float4 p10 = __ldg(reinterpret_cast<const float4 *>(base + ((long long)h1 * W + w0) * C));
        const float4 p11 = __ldg(reinterpret_cast<const float4 *>(base + ((long long)h1 * W + w1) * C));
        float4 o;
        o.x = a0 * (b0 * p00.x + b1 * p01.x) + a1 * (b0 * p10.x + b1 * p11.x);
        o.y = a0 * (b0 * p00.y + b1 * p01.y) + a1 * (b0 * p10.y + b1 * p11.y);
        o.z = a0 * (b0 * p00.z + b1 * p01.z) + a1 * (b0 * p10.z + b1 * p11.z);
        o.w = a0 * (b0 * p00.w + b1 * p01.w) + a1 * (b0 * p10.w + b1 * p11.w);
        st4<TY>(y, i, o);
    }
}

// weight with which output index `dst` taps input index `i` along one dimension
__device__ __forceinline__ float bil_weight(float r, int dst, int in, int i)
{
    int i0, i1;
    float l0, l1;
    bil_coord(r, dst, in, i0, i1, l0, l1);
    float w = 0.f;
    if (i0 == i) w += l0;
    if (i1 == i) w += l1;
    return w;
}

template <typename TDy>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const TDy *__restrict__ dy, float *__restrict__ dx, int N, int H, int W, int C, float rh,
                      float rw)
{
    const int Ho = H * 2, Wo = W * 2, vec = C >> 2;
    const long long total = (long long)N * H * W * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i % vec);
        long long t = i / vec;
        const int w = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        // outputs whose source index lies in (h-1, h+1): dst in ((h-1)/r, (h+1)/r)
        const int hlo = max(0, (int)floorf((float)(h - 1) / rh) - 1), hhi = min(Ho - 1, (int)ceilf((float)(h + 1) / rh) + 1);
        const int wlo = max(0, (int)floorf((float)(w - 1) / rw) - 1), whi = min(Wo - 1, (int)ceilf((float)(w + 1) / rw) + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long base4 = (long long)n * Ho * Wo * vec + lane;
        for (int ho = hlo; ho <= hhi; ++ho) {
            const float wh = bil_weight(rh, ho, H, h);
            if (wh == 0.f) continue;
            for (int wo = wlo; wo <= whi; ++wo) {
                const float ww = bil_weight(rw, wo, W, w);
                if (ww == 0.f) continue;
                const float4 g = ld4<TDy>(dy, base4 + ((long long)ho * Wo + wo) * vec);
                const float k = wh * ww;
                acc.x += k * g.x; acc.y += k * g.y; acc.z += k * g.z; acc.w += k * g.w;
            }
        }
        *reinterpret_cast<float4 *>(dx + i * 4) = acc;
    }
}

}  // namespace

SCDA_API int scda_upsample_bilinear2x_nhwc(int N, int H, int W, int C, const float *x, void *y, int y_dtype,
                                           cudaStream_t stream)
{
    if (N <= 0 || H < 2 || W < 2 || C < 4 || C % 4 || !x || !y || (y_dtype != 0 && y_dtype != 1)) return 0;
    if (((uintptr_t)x | (uintptr_t)y) % 16) return 0;
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    const long long total = (long long)N * H * 2 * W * 2 * (C >> 2);
    if (y_dtype == 0)
        upsample2x_fwd_kernel<float><<<apply_grid(total), 256, 0, stream>>>(x, (float *)y, N, H, W, C, rh, rw);
    else
        upsample2x_fwd_kernel<__nv_bfloat16><<<apply_grid(total), 256, 0, stream>>>(x, (__nv_bfloat16 *)y, N, H, W, C,
                                                                                    rh, rw);
    return scda_launch_status();
}

SCDA_API int scda_upsample_bilinear2x_nhwc_f32(int N, int H, int W, int C, const float *x, float *y,
                                               cudaStream_t stream)
{
    return scda_upsample_bilinear2x_nhwc(N, H, W, C, x, y, 0, stream);
}

SCDA_API int scda_upsample_bilinear2x_bwd_nhwc(int N, int H, int W, int C, const void *dy, int dy_dtype, float *dx,
                                               cudaStream_t stream)
{
    if (N <= 0 || H < 2 || W < 2 || C < 4 || C % 4 || !dy || !dx || (dy_dtype != 0 && dy_dtype != 1)) return 0;
    if (((uintptr_t)dy | (uintptr_t)dx) % 16) return 0;
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    const long long total = (long long)N * H * W * (C >> 2);
    if (dy_dtype == 0)
        upsample2x_bwd_kernel<float><<<apply_grid(total), 256, 0, stream>>>((const float *)dy, dx, N, H, W, C, rh, rw);
    else
        upsample2x_bwd_kernel<__nv_bfloat16><<<apply_grid(total), 256, 0, stream>>>((const __nv_bfloat16 *)dy, dx, N, H,
                                                                                    W, C, rh, rw);
    return scda_launch_status();
}

SCDA_API int scda_upsample_bilinear2x_bwd_nhwc_f32(int N, int H, int W, int C, const float *dy, float *dx,
                                                   cudaStream_t stream)
{
    return scda_upsample_bilinear2x_bwd_nhwc(N, H, W, C, dy, 0, dx, stream);
}
