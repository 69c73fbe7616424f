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

// ---------------------------------------------------------------------------
// Plane-resident forms (the main ones: any map whose channel plane fits in shared memory).
//
// The per-(RoI, chunk) kernels above gather four scalars per output straight from L2: 4-byte reads that each
// pull a 32-byte sector (26 MB requested, ~200 MB of sectors moved at config 1), and the backward lands
// 4 scattered L2 atomics per output (6.4 M at config 1 — the L2 atomic rate is the bound, ours and the
// reference's alike).  Turned inside out: a CTA owns `cc` channel planes of one image in shared memory (row
// pitch odd: rows fall in different banks) and walks over its share of the RoIs, 16 at a time, whose sample
// taps are derived cooperatively into shared memory.
//   forward : planes staged once, coalesced; every output = 4 shared-memory reads + the reference's arithmetic
//             (same operation order: same bits); streaming stores, consecutive lanes = consecutive addresses.
//   backward: the planes are ACCUMULATORS (shared-memory atomics; the gradient stream is read once,
//             coalesced); at the end each CTA adds its non-zero cells to the caller's gradient with
//             16-byte vector reductions (red.global.add.v4.f32) — global atomic traffic falls from
//             4 per output to one vector per 4 touched cells per RoI split.
constexpr int kAPThreads = 512;
constexpr int kAPRoiBatch = 16;

__device__ __forceinline__ int div_small(int x, float inv)      // floor(x / d), 0 <= x < 2^21, inv = 1.f / d
{
    return __float2int_rz(__fmul_rn(__int2float_rn(x) + 0.5f, inv));
}

struct PlaneTap {
    int ul;      // offset of the up-left tap in a shared-memory plane (h * pitch + w), -1 = outside
    float hr, wr;
};

// the taps of RoIs [rb, rb + nb) and their image indices
__device__ __forceinline__ void stage_taps(const float *__restrict__ rois, int rb, int nb, float scale, int H, int W,
                                           int AH, int AW, int pitch, PlaneTap *s_tap, int *s_img)
{
    const int bins = AH * AW;
    const float inv_bins = 1.f / (float)bins, inv_aw = 1.f / (float)AW;
    for (int t = threadIdx.x; t < nb * bins; t += kAPThreads) {
        const int j = div_small(t, inv_bins), b = t - j * bins;
        const int ph = div_small(b, inv_aw), pw = b - ph * AW;
        const float *r = rois + 5 * (rb + j);
        const Tap tp = make_tap(r, scale, H, W, AH, AW, ph, pw);
        PlaneTap q;
        q.hr = tp.hr;
        q.wr = tp.wr;
        if (tp.ul >= 0) {
            const int hs = tp.ul / W;
            q.ul = hs * pitch + (tp.ul - hs * W);
        } else {
            q.ul = -1;
        }
        s_tap[t] = q;
        if (b == 0) s_img[j] = (int)r[0];
    }
}

// smallest image index > cur among RoIs [r0, r1), 0x7fffffff if none (all threads get the value)
__device__ __forceinline__ int next_image(const float *__restrict__ rois, int r0, int r1, int cur, int *s_next)
{
    if (threadIdx.x == 0) *s_next = 0x7fffffff;
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += kAPThreads) {
        const int b = (int)__ldg(rois + 5 * r);
        if (b > cur) atomicMin(s_next, b);
    }
    __syncthreads();
    return *s_next;
}

__global__ void __launch_bounds__(kAPThreads, 2)
roi_align_fwd_plane_kernel(const float *__restrict__ feat, float scale, int H, int W, int C, int AH, int AW,
                           const float *__restrict__ rois, int R, float *__restrict__ out, int cc, int pitch,
                           int rois_per_cta, int vec_in)
{
    extern __shared__ __align__(16) unsigned char s_bytes[];
    __shared__ int s_next;
    __shared__ int s_img[kAPRoiBatch];
    float *s_plane = reinterpret_cast<float *>(s_bytes);
    PlaneTap *s_tap = reinterpret_cast<PlaneTap *>(s_plane + (size_t)cc * H * pitch);
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * cc, cn = min(cc, C - c0);
    const int r0 = blockIdx.y * rois_per_cta, r1 = min(R, r0 + rois_per_cta);
    const int bins = AH * AW, HW = H * W, run = cn * bins, psz = H * pitch;
    const float inv_bins = 1.f / (float)bins, inv_run = 1.f / (float)run;

    int cur = -1;
    for (;;) {
        const int img = next_image(rois, r0, r1, cur, &s_next);
        if (img == 0x7fffffff) break;
        {
            // the cn planes are one contiguous stretch of the NCHW map: independent 16-byte loads, four in
            // flight per thread
            const float *__restrict__ src = feat + ((long long)img * C + c0) * HW;
            const float inv_w = 1.f / (float)W;
            if (vec_in) {
                const int n4 = (cn * HW) >> 2;
                for (int i0 = tid; i0 < n4; i0 += 4 * kAPThreads) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kAPThreads;
                        if (i < n4) v[u] = __ldg(reinterpret_cast<const float4 *>(src) + i);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kAPThreads;
                        if (i >= n4) continue;
                        const int row = div_small(i << 2, inv_w), w = (i << 2) - row * W;   // row = c * H + h
                        float *dst = s_plane + (size_t)row * pitch + w;
                        dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w;
                    }
                }
            } else {
                const int n1 = cn * HW;
                for (int i0 = tid; i0 < n1; i0 += 4 * kAPThreads) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kAPThreads;
                        if (i < n1) v[u] = __ldg(src + i);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kAPThreads;
                        if (i >= n1) continue;
                        const int row = div_small(i, inv_w), w = i - row * W;
                        s_plane[(size_t)row * pitch + w] = v[u];
                    }
                }
            }
        }
        for (int rb = r0; rb < r1; rb += kAPRoiBatch) {
            const int nb = min(kAPRoiBatch, r1 - rb);
            __syncthreads();
            stage_taps(rois, rb, nb, scale, H, W, AH, AW, pitch, s_tap, s_img);
            __syncthreads();
            const int total = nb * run;
            for (int o = tid; o < total; o += kAPThreads) {
                const int j = div_small(o, inv_run), rem = o - j * run;
                if (s_img[j] != img) continue;
                const int c = div_small(rem, inv_bins), b = rem - c * bins;
                const PlaneTap t = s_tap[j * bins + b];
                float v = 0.f;
                if (t.ul >= 0) {
                    const float *p = s_plane + (size_t)c * psz + t.ul;
                    const float ul = p[0], ur = p[1], dl = p[pitch], dr = p[pitch + 1];
                    const double omh = __dsub_rn(1.0, (double)t.hr), omw = __dsub_rn(1.0, (double)t.wr);
                    double acc = __fma_rn(omw, __dmul_rn(omh, (double)ul),
                                          __dmul_rn(__dmul_rn(omh, (double)ur), (double)t.wr));
                    acc = __fma_rn(omw, (double)__fmul_rn(t.hr, dl), acc);
                    acc = __dadd_rn(acc, (double)__fmul_rn(t.wr, __fmul_rn(t.hr, dr)));
                    v = (float)acc;
                }
                asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(out + ((long long)(rb + j) * C + c0) * bins + rem),
                             "f"(v) : "memory");
            }
        }
        cur = img;
    }
}

__global__ void __launch_bounds__(kAPThreads, 2)
roi_align_bwd_plane_kernel(const float *__restrict__ top_diff, float scale, int H, int W, int C, int AH, int AW,
                           const float *__restrict__ rois, int R, float *__restrict__ bottom_diff, int cc,
                           int pitch, int rois_per_cta, int vec4)
{
    extern __shared__ __align__(16) unsigned char s_bytes[];
    __shared__ int s_next;
    __shared__ int s_img[kAPRoiBatch];
    float *s_plane = reinterpret_cast<float *>(s_bytes);
    PlaneTap *s_tap = reinterpret_cast<PlaneTap *>(s_plane + (size_t)cc * H * pitch);
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * cc, cn = min(cc, C - c0);
    const int r0 = blockIdx.y * rois_per_cta, r1 = min(R, r0 + rois_per_cta);
    const int bins = AH * AW, HW = H * W, run = cn * bins, psz = H * pitch;
    const float inv_bins = 1.f / (float)bins, inv_run = 1.f / (float)run, inv_h = 1.f / (float)H;

    int cur = -1;
    for (;;) {
        const int img = next_image(rois, r0, r1, cur, &s_next);
        if (img == 0x7fffffff) break;
        for (int i = tid; i < cn * psz; i += kAPThreads) s_plane[i] = 0.f;
        for (int rb = r0; rb < r1; rb += kAPRoiBatch) {
            const int nb = min(kAPRoiBatch, r1 - rb);
            __syncthreads();
            stage_taps(rois, rb, nb, scale, H, W, AH, AW, pitch, s_tap, s_img);
            __syncthreads();
            const int total = nb * run;
            for (int o0 = tid; o0 < total; o0 += 4 * kAPThreads) {
                // the gradient stream first (independent loads), then the shared-memory atomics
                float gv[4];
                int jj[4], rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int o = o0 + u * kAPThreads;
                    jj[u] = -1;
                    if (o < total) {
                        const int j = div_small(o, inv_run), rem = o - j * run;
                        if (s_img[j] == img) {
                            jj[u] = j;
                            rr[u] = rem;
                            gv[u] = __ldg(top_diff + ((long long)(rb + j) * C + c0) * bins + rem);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (jj[u] < 0) continue;
                    const int c = div_small(rr[u], inv_bins), b = rr[u] - c * bins;
                    const PlaneTap t = s_tap[jj[u] * bins + b];
                    if (t.ul < 0) continue;
                    const float g = gv[u];
                    float *p = s_plane + (size_t)c * psz + t.ul;
                    // roi_align_kernel.cu:136-139: the first two products are formed in double, the last two in float
                    const double gh = __dmul_rn((double)g, __dsub_rn(1.0, (double)t.hr));
                    const float omw = __fsub_rn(1.f, t.wr);
                    const float gl = __fmul_rn(g, t.hr);
                    atomicAdd(p, (float)__dmul_rn(gh, (double)omw));
                    atomicAdd(p + 1, (float)__dmul_rn(gh, (double)t.wr));
                    atomicAdd(p + pitch, __fmul_rn(gl, omw));
                    atomicAdd(p + pitch + 1, __fmul_rn(gl, t.wr));
                }
            }
        }
        __syncthreads();
        // add the touched cells to the caller's gradient
        const int wq = (W + 3) >> 2;                       // column quads per row
        const float inv_wq = 1.f / (float)wq;
        for (int i = tid; i < cn * H * wq; i += kAPThreads) {
            const int row = div_small(i, inv_wq), w = (i - row * wq) << 2;
            const int c = div_small(row, inv_h), h = row - c * H;
            const float *p = s_plane + (size_t)c * psz + h * pitch + w;
            float *gp = bottom_diff + ((long long)img * C + c0 + c) * HW + (long long)h * W + w;
            const int nw = min(4, W - w);
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = k < nw ? p[k] : 0.f;
            if (v[0] == 0.f && v[1] == 0.f && v[2] == 0.f && v[3] == 0.f) continue;
            if (vec4 && nw == 4) {
                asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(gp), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                             "f"(v[3]) : "memory");
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nw && v[k] != 0.f) red_add_f32(gp + k, v[k]);
            }
        }
        __syncthreads();
        cur = img;
    }
}

struct PlanePlan {
    int ok, cc, pitch, per_cta;
    size_t smem;
    dim3 grid;
};

PlanePlan plan_align_plane(int R, int H, int W, int C, int AH, int AW, int split_div)
{
    PlanePlan p;
    p.ok = 0;
    p.pitch = W | 1;
    const size_t plane = sizeof(float) * (size_t)H * p.pitch;
    const size_t taps = sizeof(PlaneTap) * (size_t)kAPRoiBatch * AH * AW;
    const size_t budget = 100 * 1024;                      // two CTAs of 512 threads x 64 registers per SM
    if (plane + taps > 200 * 1024) return p;
    int cc = taps < budget ? (int)((budget - taps) / plane) : 0;
    if (cc < 1) cc = 1;
    if (cc > C) cc = C;
    for (int d = cc; 2 * d > cc; --d)
        if (C % d == 0) { cc = d; break; }
    if ((long long)kAPRoiBatch * cc * AH * AW >= (1 << 21) || (long long)cc * H * ((W + 3) / 4) >= (1 << 21)) return p;
    p.cc = cc;
    p.smem = plane * cc + taps;
    const int chunks = ceil_div(C, cc);
    const int per_sm = p.smem > 110 * 1024 ? 1 : 2;
    int rsplit = (kNumSMs * per_sm) / chunks / split_div;
    const int max_split = ceil_div(R, kAPRoiBatch);
    if (rsplit > max_split) rsplit = max_split;
    if (rsplit < 1) rsplit = 1;
    p.per_cta = ceil_div(ceil_div(R, rsplit), kAPRoiBatch) * kAPRoiBatch;
    p.grid = dim3(chunks, ceil_div(R, p.per_cta));
    p.ok = 1;
    return p;
}

template <typename K>
int plane_smem_attr(K kernel, size_t smem, size_t *attr)
{
    if (smem > 48 * 1024 && smem > *attr) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        *attr = smem;
    }
    return 1;
}

bool align_args_ok(int num_rois, int H, int W, int C, int AH, int AW, const void *a, const void *b,
                   const void *c)
{
    return num_rois >= 0 && H >= 2 && W >= 2 && C > 0 && AH >= 2 && AW >= 2 && a && b && c;
}

int g_align_bwd_split_div = 2;

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
    const PlanePlan pl = plan_align_plane(num_rois, height, width, channels, aligned_height, aligned_width, 1);
    if (pl.ok) {
        static size_t attr = 0;
        const int st = plane_smem_attr(roi_align_fwd_plane_kernel, pl.smem, &attr);
        if (st != 1) return st;
        roi_align_fwd_plane_kernel<<<pl.grid, kAPThreads, pl.smem, stream>>>(
            bottom_data, spatial_scale, height, width, channels, aligned_height, aligned_width, bottom_rois,
            num_rois, top_data, pl.cc, pl.pitch, pl.per_cta,
            width % 4 == 0 && (height * width) % 4 == 0 && (uintptr_t)bottom_data % 16 == 0);
        return scda_launch_status();
    }
    // maps whose channel plane does not fit in shared memory
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
    const PlanePlan pl = plan_align_plane(num_rois, height, width, channels, aligned_height, aligned_width,
                                          g_align_bwd_split_div);
    if (pl.ok) {
        static size_t attr = 0;
        const int st = plane_smem_attr(roi_align_bwd_plane_kernel, pl.smem, &attr);
        if (st != 1) return st;
        const int vec4 = width % 4 == 0 && (uintptr_t)bottom_diff % 16 == 0;
        roi_align_bwd_plane_kernel<<<pl.grid, kAPThreads, pl.smem, stream>>>(
            top_diff, spatial_scale, height, width, channels, aligned_height, aligned_width, bottom_rois, num_rois,
            bottom_diff, pl.cc, pl.pitch, pl.per_cta, vec4);
        return scda_launch_status();
    }
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
