// RoI max pooling, forward and backward, for sm_100a.
//
// Replaces ROIPoolForward / ROIPoolBackward of the reference
// (extensions/_roi_pooling/src/roi_pooling_kernel.cu:24-93, 128-203); results
// are bit-identical on output and argmax (max and integer bin arithmetic only).
//
// Shape of the work at the model's operating point (features 1x512x32x64,
// 512 RoIs, 7x7): 4.2 MB read (L2 resident after first touch), 51.4 MB output
// + 51.4 MB argmax written -> HBM-write bound.  Design:
//   forward : one CTA per (RoI, 64-channel chunk); the RoI's integer bin edges
//             are derived once per CTA into shared memory instead of once per
//             output element.  Main form (pooled_height <= 8): separable max,
//             one warp per (RoI, channel[s]) — see roi_pool_fwd_warp_kernel.
//             Generic form: each thread produces 4 consecutive outputs and
//             issues one 128-bit streaming store for values and one for argmax.
//   backward: the reference gathers — every input element loops over all RoIs
//             (O(B*C*H*W*R)).  Here the 51 MB gradient and argmax streams are
//             read exactly once with 128-bit loads and scattered with
//             fire-and-forget red.global.add into the 4 MB input gradient,
//             which lives in L2.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kPoolThreads = 256;
constexpr int kPoolChunkC = 64;

struct PoolRoi {
    int batch, x0, y0;
    float bin_h, bin_w;
};

// roi_pooling_kernel.cu:45-56
__device__ __forceinline__ PoolRoi load_pool_roi(const float *r, float scale, int PH, int PW)
{
    PoolRoi q;
    q.batch = (int)r[0];
    q.x0 = (int)roundf(r[1] * scale);
    q.y0 = (int)roundf(r[2] * scale);
    int x1 = (int)roundf(r[3] * scale);
    int y1 = (int)roundf(r[4] * scale);
    int rw = (int)fmaxf((float)(x1 - q.x0 + 1), 1.f);
    int rh = (int)fmaxf((float)(y1 - q.y0 + 1), 1.f);
    q.bin_h = __fdiv_rn((float)rh, (float)PH);
    q.bin_w = __fdiv_rn((float)rw, (float)PW);
    return q;
}

__device__ __forceinline__ int clamp_edge(int v, int hi)
{
    return (int)fminf(fmaxf((float)v, 0.f), (float)hi);
}

template <bool kVec>
__global__ void __launch_bounds__(kPoolThreads)
roi_pool_fwd_kernel(const float *__restrict__ feat, float scale, int H, int W, int C, int PH,
                    int PW, const float *__restrict__ rois, float *__restrict__ out,
                    int *__restrict__ argmax)
{
    extern __shared__ int s_edge[];  // hs[PH] he[PH] ws[PW] we[PW]
    int *hs = s_edge, *he = hs + PH, *ws = he + PH, *we = ws + PW;
    const int n = blockIdx.x;
    const int c0 = blockIdx.y * kPoolChunkC;
    const PoolRoi q = load_pool_roi(rois + 5 * n, scale, PH, PW);

    for (int i = threadIdx.x; i < PH + PW; i += kPoolThreads) {
        if (i < PH) {
            hs[i] = clamp_edge((int)floorf(__fmul_rn((float)i, q.bin_h)) + q.y0, H);
            he[i] = clamp_edge((int)ceilf(__fmul_rn((float)(i + 1), q.bin_h)) + q.y0, H);
        } else {
            int j = i - PH;
            ws[j] = clamp_edge((int)floorf(__fmul_rn((float)j, q.bin_w)) + q.x0, W);
            we[j] = clamp_edge((int)ceilf(__fmul_rn((float)(j + 1), q.bin_w)) + q.x0, W);
        }
    }
    __syncthreads();

    const int bins = PH * PW;
    const int cn = min(kPoolChunkC, C - c0);
    const int total = cn * bins;
    const long long obase = ((long long)n * C + c0) * bins;
    const int plane0 = (q.batch * C + c0) * H * W;
    constexpr int kPer = kVec ? 4 : 1;

    for (int e0 = threadIdx.x * kPer; e0 < total; e0 += kPoolThreads * kPer) {
        float val[kPer];
        int idx[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int e = e0 + k;  // kVec => total % 4 == 0, always in range
            const int c = e / bins, b = e - c * bins;
            const int ph = b / PW, pw = b - ph * PW;
            const int h0 = hs[ph], h1 = he[ph], w0 = ws[pw], w1 = we[pw];
            const int plane = plane0 + c * H * W;
            const float *__restrict__ p = feat + plane;
            float best = (h1 <= h0 || w1 <= w0) ? 0.f : -FLT_MAX;
            int where = -1;
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) {
                    float v = __ldg(p + h * W + w);
                    if (v > best) { best = v; where = plane + h * W + w; }
                }
            val[k] = best;
            idx[k] = where;
        }
        if (kVec) {
            st_stream_f4(out + obase + e0, make_float4(val[0], val[1 % kPer], val[2 % kPer], val[3 % kPer]));
            if (argmax)
                st_stream_i4(argmax + obase + e0, make_int4(idx[0], idx[1 % kPer], idx[2 % kPer], idx[3 % kPer]));
        } else {
            out[obase + e0] = val[0];
            if (argmax) argmax[obase + e0] = idx[0];
        }
    }
}

// ---------------------------------------------------------------------------
// Forward, warp-cooperative form (used when pooled_height <= 8).
//
// Max over a bin window is separable.  A warp owns (RoI, channel[s]):
//   phase A: lanes lie along the RoI's clipped column range; each lane reduces
//            its column over the h-range of every row-bin -> T[ph][w] (value and
//            the first row that attains it).  Loads are coalesced row segments
//            and every lane runs the same trip counts (edges come from shared
//            memory), so there is no divergence.  Narrow RoIs pack 32/Lc
//            channels into one warp so lanes stay busy.
//   phase B: lanes take (channel, ph, pw) outputs and reduce T[ph][ws..we) from
//            shared memory, keeping the lexicographically first (h, w) among
//            equal maxima — the element the reference's row-major strict-'>'
//            scan selects (post-ReLU maps are full of exact ties at 0).
//   Stores are consecutive addresses across the warp.
constexpr int kPoolWarps = kPoolThreads / 32;

template <int kPH>
__global__ void __launch_bounds__(kPoolThreads)
roi_pool_fwd_warp_kernel(const float *__restrict__ feat, float scale, int H, int W, int C, int PW,
                         const float *__restrict__ rois, float *__restrict__ out,
                         int *__restrict__ argmax, int rowlen, int edge_words)
{
    extern __shared__ int s_raw[];
    int *hs = s_raw, *he = hs + kPH, *ws = he + kPH, *we = ws + PW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *Tv = reinterpret_cast<float *>(s_raw + edge_words) + warp * (2 * kPH * rowlen);
    int *Th = reinterpret_cast<int *>(Tv + kPH * rowlen);

    const int n = blockIdx.x;
    const int c0 = blockIdx.y * kPoolChunkC;
    const PoolRoi q = load_pool_roi(rois + 5 * n, scale, kPH, PW);
    for (int i = threadIdx.x; i < kPH + PW; i += kPoolThreads) {
        if (i < kPH) {
            hs[i] = clamp_edge((int)floorf(__fmul_rn((float)i, q.bin_h)) + q.y0, H);
            he[i] = clamp_edge((int)ceilf(__fmul_rn((float)(i + 1), q.bin_h)) + q.y0, H);
        } else {
            int j = i - kPH;
            ws[j] = clamp_edge((int)floorf(__fmul_rn((float)j, q.bin_w)) + q.x0, W);
            we[j] = clamp_edge((int)ceilf(__fmul_rn((float)(j + 1), q.bin_w)) + q.x0, W);
        }
    }
    __syncthreads();

    const int w_lo = ws[0];
    const int rw = we[PW - 1] - w_lo;  // clipped column extent (<= W); <= 0: every bin empty
    int Lc = 32;
    if (rw <= 16) Lc = 16;
    if (rw <= 8) Lc = 8;
    if (rw <= 4) Lc = 4;
    if (rw <= 2) Lc = 2;
    if (rw <= 1) Lc = 1;
    const int cpw = 32 / Lc;
    const int sub = lane / Lc, wl = lane - sub * Lc;
    const int nchunk = Lc == 32 ? (rw + 31) / 32 : (rw > 0 ? 1 : 0);
    const int bins = kPH * PW;
    const int cn = min(kPoolChunkC, C - c0);
    const int HW = H * W;
    const long long obase = ((long long)n * C + c0) * bins;
    const int plane0 = (q.batch * C + c0) * HW;

    for (int cb = warp * cpw; cb < cn; cb += kPoolWarps * cpw) {
        // ---- phase A
        const bool cvalid = cb + sub < cn;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int wcol = ch * 32 + wl;
            const bool valid = cvalid && wcol < rw;
            const float *__restrict__ p = feat + plane0 + (cb + sub) * HW + w_lo + wcol;
#pragma unroll
            for (int ph = 0; ph < kPH; ++ph) {
                const int h0 = hs[ph], h1 = he[ph];
                float best = -FLT_MAX;
                int bh = -1;
                if (valid) {
                    for (int h = h0; h < h1; ++h) {
                        const float v = __ldg(p + h * W);
                        if (v > best) { best = v; bh = h; }
                    }
                }
                if (Lc < 32 || wcol < rw) {  // the last 32-column chunk may overhang the row
                    const int ti = ph * rowlen + sub * Lc + wcol;
                    Tv[ti] = best;
                    Th[ti] = bh;
                }
            }
        }
        __syncwarp();
        // ---- phase B
        const int ntask = min(cpw, cn - cb) * bins;
        for (int t = lane; t < ntask; t += 32) {
            const int s = t / bins, b = t - s * bins;
            const int ph = b / PW, pw = b - ph * PW;
            const int w0 = ws[pw], w1 = we[pw];
            const bool empty = he[ph] <= hs[ph] || w1 <= w0;
            float best = empty ? 0.f : -FLT_MAX;
            int bh = 0x7fffffff, bw = -1;
            if (!empty) {
                const int row = ph * rowlen + s * Lc - w_lo;
                for (int w = w0; w < w1; ++w) {
                    const float v = Tv[row + w];
                    const int h = Th[row + w];
                    if (h >= 0 && (v > best || (v == best && h < bh))) { best = v; bh = h; bw = w; }
                }
            }
            const long long o = obase + (long long)(cb + s) * bins + b;
            out[o] = best;
            if (argmax) argmax[o] = bw >= 0 ? plane0 + (cb + s) * HW + bh * W + bw : -1;
        }
        __syncwarp();
    }
}

template <int kPH>
int launch_pool_fwd_warp(const float *feat, float scale, int R, int H, int W, int C, int PW,
                         const float *rois, float *out, int *argmax, cudaStream_t stream)
{
    const int rowlen = W > 32 ? W : 32;
    const int edge_words = (2 * kPH + 2 * PW + 3) & ~3;
    const size_t smem = sizeof(int) * ((size_t)edge_words + (size_t)kPoolWarps * 2 * kPH * rowlen);
    if (smem > 200 * 1024) return 2;  // caller falls back to the per-output kernel
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_warp_kernel<kPH>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
    }
    dim3 grid(R, ceil_div(C, kPoolChunkC));
    roi_pool_fwd_warp_kernel<kPH><<<grid, kPoolThreads, smem, stream>>>(
        feat, scale, H, W, C, PW, rois, out, argmax, rowlen, edge_words);
    return scda_launch_status();
}

// ---------------------------------------------------------------------------
// Forward, plane-resident form (the main one: any map whose channel plane fits in shared memory).
//
// The warp form above reads every RoI's window from L2 with one dependent load per row: at the model's
// operating point (512 RoIs x 512 channels, windows of ~200 cells) that is ~210 MB of scattered L2 reads behind
// chains of ~20 load latencies per (RoI, channel) — 235 us, 7 % of the HBM rate the 103 MB of output could be
// written at.  Here the loop nest is turned inside out: a CTA owns FOUR channel planes of one image, stages
// them once in shared memory interleaved per pixel ([h][w][4 channels]: one 16-byte shared-memory load feeds
// four channels) and walks over its share of the RoIs, 32 at a time, whose integer bin edges are derived
// cooperatively into shared memory.  A thread owns a (RoI, bin) and runs the reference's own row-major
// strict-'>' scan (roi_pooling_kernel.cu:75-87) for the four channels at once — bit-identical values and
// argmax, no tie rule to re-derive; the scan's loop and address arithmetic (the bulk of the instructions: the
// first, one-channel-per-thread version of this kernel was issue bound at 160 us) is paid once per four
// outputs.  Consecutive lanes own consecutive bins of one RoI, so each of a thread's four value stores (and
// four argmax stores) lands in a stretch of up to 128 contiguous bytes per warp: streaming 4-byte stores.
// Global memory sees the feature map read once per RoI split plus the output stream.
constexpr int kPlaneThreads = 512;
constexpr int kPlaneRoiBatch = 32;
constexpr int kPlaneC = 4;

// floor(x / d) for 0 <= x < 2^21 with inv = 1.f / d: (x + 0.5) / d is at least 0.5 / d away from an integer,
// more than the rounding error of the two float operations
__device__ __forceinline__ int div_small(int x, float inv)
{
    return __float2int_rz(__fmul_rn(__int2float_rn(x) + 0.5f, inv));
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kPlaneThreads, kMinBlocks)
roi_pool_fwd_plane_kernel(const float *__restrict__ feat, float scale, int H, int W, int C, int PH, int PW,
                          const float *__restrict__ rois, int R, float *__restrict__ out,
                          int *__restrict__ argmax, int pitch, int rois_per_cta, int vec_in, int roi_batch)
{
    extern __shared__ __align__(16) unsigned char s_bytes[];
    __shared__ int s_next;
    __shared__ int s_img[kPlaneRoiBatch];
    const int ne = PH + PW, ew = 2 * ne;
    const int bins = PH * PW, HW = H * W;
    float4 *s_plane = reinterpret_cast<float4 *>(s_bytes);                                   // [H][pitch] x 4 ch
    unsigned short *s_edge = reinterpret_cast<unsigned short *>(s_plane + (size_t)H * pitch);   // [batch][hs he ws we]
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kPlaneC, cn = min(kPlaneC, C - c0);
    const int r0 = blockIdx.y * rois_per_cta, r1 = min(R, r0 + rois_per_cta);
    const float inv_bins = 1.f / (float)bins, inv_pw = 1.f / (float)PW;
    const float inv_h = 1.f / (float)H, inv_ne = 1.f / (float)ne;

    int cur = -1;   // image whose planes are resident
    for (;;) {
        // next image index (ascending) referenced by one of this CTA's RoIs
        if (tid == 0) s_next = 0x7fffffff;
        __syncthreads();
        for (int r = r0 + tid; r < r1; r += kPlaneThreads) {
            const int b = (int)__ldg(rois + 5 * r);
            if (b > cur) atomicMin(&s_next, b);
        }
        __syncthreads();
        const int img = s_next;
        if (img == 0x7fffffff) break;
        {
            // the cn planes are one contiguous stretch of the NCHW map: independent 16-byte loads, four in
            // flight per thread (one dependent load per row made the staging alone ~13 us of memory latency)
            float *sp = reinterpret_cast<float *>(s_plane);
            const float *__restrict__ src = feat + ((long long)img * C + c0) * HW;
            const float inv_w = 1.f / (float)W;
            if (vec_in) {
                const int n4 = (cn * HW) >> 2;
                for (int i0 = tid; i0 < n4; i0 += 4 * kPlaneThreads) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kPlaneThreads;
                        if (i < n4) v[u] = __ldg(reinterpret_cast<const float4 *>(src) + i);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kPlaneThreads;
                        if (i >= n4) continue;
                        const int row = div_small(i << 2, inv_w), w = (i << 2) - row * W;   // row = c * H + h
                        const int c = div_small(row, inv_h), h = row - c * H;
                        float *dst = sp + ((size_t)h * pitch + w) * kPlaneC + c;
                        dst[0] = v[u].x; dst[kPlaneC] = v[u].y; dst[2 * kPlaneC] = v[u].z; dst[3 * kPlaneC] = v[u].w;
                    }
                }
            } else {
                const int n1 = cn * HW;
                for (int i0 = tid; i0 < n1; i0 += 4 * kPlaneThreads) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kPlaneThreads;
                        if (i < n1) v[u] = __ldg(src + i);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kPlaneThreads;
                        if (i >= n1) continue;
                        const int row = div_small(i, inv_w), w = i - row * W;
                        const int c = div_small(row, inv_h), h = row - c * H;
                        sp[((size_t)h * pitch + w) * kPlaneC + c] = v[u];
                    }
                }
            }
            if (cn < kPlaneC)
                for (int i = tid; i < H * W; i += kPlaneThreads) {
                    const int h = div_small(i, inv_w), w = i - h * W;
                    for (int c = cn; c < kPlaneC; ++c) sp[((size_t)h * pitch + w) * kPlaneC + c] = 0.f;
                }
        }
        for (int rb = r0; rb < r1; rb += roi_batch) {
            const int nb = min(roi_batch, r1 - rb);
            __syncthreads();   // planes staged / previous batch's edges and staged outputs no longer read
            for (int t = tid; t < nb * ne; t += kPlaneThreads) {
                const int j = div_small(t, inv_ne), i = t - j * ne;
                const float *r = rois + 5 * (rb + j);
                const PoolRoi q = load_pool_roi(r, scale, PH, PW);
                unsigned short *e = s_edge + j * ew;
                if (i < PH) {
                    e[i] = clamp_edge((int)floorf(__fmul_rn((float)i, q.bin_h)) + q.y0, H);
                    e[PH + i] = clamp_edge((int)ceilf(__fmul_rn((float)(i + 1), q.bin_h)) + q.y0, H);
                } else {
                    const int k = i - PH;
                    e[2 * PH + k] = clamp_edge((int)floorf(__fmul_rn((float)k, q.bin_w)) + q.x0, W);
                    e[2 * PH + PW + k] = clamp_edge((int)ceilf(__fmul_rn((float)(k + 1), q.bin_w)) + q.x0, W);
                }
                if (i == 0) s_img[j] = q.batch;
            }
            __syncthreads();
            for (int o = tid; o < nb * bins; o += kPlaneThreads) {
                const int j = div_small(o, inv_bins), b = o - j * bins;
                const unsigned short *e = s_edge + j * ew;
                if (s_img[j] != img) continue;
                const int ph = div_small(b, inv_pw), pw = b - ph * PW;
                const int h0 = e[ph], h1 = e[PH + ph], w0 = e[2 * PH + pw], w1 = e[2 * PH + PW + pw];
                const float init = (h1 <= h0 || w1 <= w0) ? 0.f : -FLT_MAX;
                float b0 = init, b1 = init, b2 = init, b3 = init;
                int p0 = -1, p1 = -1, p2 = -1, p3 = -1;
                for (int h = h0; h < h1; ++h) {
                    const float4 *prow = s_plane + h * pitch;
                    int pos = h * W + w0;
                    for (int w = w0; w < w1; ++w, ++pos) {
                        const float4 v = prow[w];
                        if (v.x > b0) { b0 = v.x; p0 = pos; }
                        if (v.y > b1) { b1 = v.y; p1 = pos; }
                        if (v.z > b2) { b2 = v.z; p2 = pos; }
                        if (v.w > b3) { b3 = v.w; p3 = pos; }
                    }
                }
                // consecutive lanes = consecutive bins of one RoI: each of the 2 x cn stores below writes a
                // contiguous stretch of up to 128 bytes
                const long long oi = ((long long)(rb + j) * C + c0) * bins + b;
                const int abase = (img * C + c0) * HW;
                st_stream_f32(out + oi, b0);
                if (cn > 1) st_stream_f32(out + oi + bins, b1);
                if (cn > 2) st_stream_f32(out + oi + 2 * bins, b2);
                if (cn > 3) st_stream_f32(out + oi + 3 * bins, b3);
                if (argmax) {
                    st_stream_s32(argmax + oi, p0 < 0 ? -1 : abase + p0);
                    if (cn > 1) st_stream_s32(argmax + oi + bins, p1 < 0 ? -1 : abase + HW + p1);
                    if (cn > 2) st_stream_s32(argmax + oi + 2 * bins, p2 < 0 ? -1 : abase + 2 * HW + p2);
                    if (cn > 3) st_stream_s32(argmax + oi + 3 * bins, p3 < 0 ? -1 : abase + 3 * HW + p3);
                }
            }
        }
        cur = img;
    }
}

// 1: launched, 2: not applicable (caller takes another form), <0: CUDA error
int launch_pool_fwd_plane(const float *feat, float scale, int R, int H, int W, int C, int PH, int PW,
                          const float *rois, float *out, int *argmax, cudaStream_t stream)
{
    const int pitch = W | 1, bins = PH * PW;
    if (H > 0xFFFF || W > 0xFFFF || PH + PW > 256) return 2;            // edges are kept as 16 bits
    const size_t plane = sizeof(float4) * (size_t)H * pitch;
    const size_t edges = sizeof(short) * (size_t)kPlaneRoiBatch * 2 * (PH + PW) + 16;
    const size_t smem = plane + edges;
    if (smem > 200 * 1024 || (long long)kPlaneRoiBatch * kPlaneC * bins >= (1 << 21)) return 2;
    static int occ = 0;
    if (!occ) {
        const char *e = getenv("SCDA_POOL_OCC");
        occ = (e && *e == '3') ? 3 : 2;
    }
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_plane_kernel<2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(roi_pool_fwd_plane_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr = smem;
    }
    const int chunks = ceil_div(C, kPlaneC);
    int per_sm = (int)((226 * 1024) / (smem + 1024));
    if (per_sm > occ) per_sm = occ;                        // 512 threads x 64 (40) registers
    if (per_sm < 1) per_sm = 1;
    // RoIs per pass: the (RoI, bin) tasks of a pass should fill whole rounds of the CTA's threads
    // (32 RoIs x 49 bins = 3 rounds of 512 + 32 stragglers: 31 RoIs fit in 3)
    int roi_batch = kPlaneRoiBatch;
    double best_fill = 0.0;
    for (int n = kPlaneRoiBatch; n >= kPlaneRoiBatch / 2; --n) {
        const int tasks = n * bins, rounds = ceil_div(tasks, kPlaneThreads);
        const double fill = (double)tasks / ((double)rounds * kPlaneThreads);
        if (fill > best_fill + 1e-9) { best_fill = fill; roi_batch = n; }
    }
    const int nbatch = ceil_div(R, roi_batch);
    int rsplit = (kNumSMs * per_sm) / chunks;              // one wave, every CTA resident
    if (rsplit > nbatch) rsplit = nbatch;
    if (rsplit < 1) rsplit = 1;
    const int per_cta = ceil_div(nbatch, rsplit) * roi_batch;
    dim3 grid(chunks, ceil_div(R, per_cta));
    const int vec_in = W % 4 == 0 && (uintptr_t)feat % 16 == 0;
    if ((long long)kPlaneC * H * W >= (1 << 21)) return 2;
    if (occ == 3)
        roi_pool_fwd_plane_kernel<3><<<grid, kPlaneThreads, smem, stream>>>(feat, scale, H, W, C, PH, PW, rois, R,
                                                                           out, argmax, pitch, per_cta, vec_in,
                                                                           roi_batch);
    else
        roi_pool_fwd_plane_kernel<2><<<grid, kPlaneThreads, smem, stream>>>(feat, scale, H, W, C, PH, PW, rois, R,
                                                                           out, argmax, pitch, per_cta, vec_in,
                                                                           roi_batch);
    return scda_launch_status();
}

template <bool kVec>
__global__ void __launch_bounds__(256)
roi_pool_bwd_scatter_kernel(const float *__restrict__ top_diff, const int *__restrict__ argmax,
                            long long total, float *__restrict__ bottom_diff)
{
    constexpr int kPer = kVec ? 4 : 1;
    const long long stride = (long long)gridDim.x * blockDim.x * kPer;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kPer; i < total;
         i += stride) {
        if (kVec) {
            const float4 g = ld_stream_f4(top_diff + i);
            const int4 a = ld_stream_i4(argmax + i);
            if (a.x >= 0) red_add_f32(bottom_diff + a.x, g.x);
            if (a.y >= 0) red_add_f32(bottom_diff + a.y, g.y);
            if (a.z >= 0) red_add_f32(bottom_diff + a.z, g.z);
            if (a.w >= 0) red_add_f32(bottom_diff + a.w, g.w);
        } else {
            const int a = argmax[i];
            if (a >= 0) red_add_f32(bottom_diff + a, top_diff[i]);
        }
    }
}

}  // namespace

SCDA_API int ROIPoolForwardLaucher(const float *bottom_data, const float spatial_scale,
                                   const int num_rois, const int height, const int width,
                                   const int channels, const int pooled_height,
                                   const int pooled_width, const float *bottom_rois,
                                   float *top_data, int *argmax_data, cudaStream_t stream)
{
    if (num_rois < 0 || height <= 0 || width <= 0 || channels <= 0 || pooled_height <= 0 ||
        pooled_width <= 0 || !bottom_data || !bottom_rois || !top_data)
        return 0;
    if (num_rois == 0) return 1;
    int st = launch_pool_fwd_plane(bottom_data, spatial_scale, num_rois, height, width, channels,
                                   pooled_height, pooled_width, bottom_rois, top_data, argmax_data, stream);
    if (st != 2) return st;
    // maps whose channel plane does not fit in shared memory
    switch (pooled_height) {
#define SCDA_POOL_CASE(P)                                                                       \
    case P:                                                                                     \
        st = launch_pool_fwd_warp<P>(bottom_data, spatial_scale, num_rois, height, width,      \
                                     channels, pooled_width, bottom_rois, top_data,            \
                                     argmax_data, stream);                                     \
        break;
        SCDA_POOL_CASE(1) SCDA_POOL_CASE(2) SCDA_POOL_CASE(3) SCDA_POOL_CASE(4)
        SCDA_POOL_CASE(5) SCDA_POOL_CASE(6) SCDA_POOL_CASE(7) SCDA_POOL_CASE(8)
#undef SCDA_POOL_CASE
    default: break;
    }
    if (st != 2) return st;
    // generic form: pooled_height > 8 or a map too wide for the shared-memory rows
    dim3 grid(num_rois, ceil_div(channels, kPoolChunkC));
    const size_t smem = sizeof(int) * 2 * (pooled_height + pooled_width);
    const bool vec = channels % 4 == 0 && ((uintptr_t)top_data % 16 == 0) &&
                     (argmax_data == nullptr || (uintptr_t)argmax_data % 16 == 0);
    if (vec)
        roi_pool_fwd_kernel<true><<<grid, kPoolThreads, smem, stream>>>(
            bottom_data, spatial_scale, height, width, channels, pooled_height, pooled_width,
            bottom_rois, top_data, argmax_data);
    else
        roi_pool_fwd_kernel<false><<<grid, kPoolThreads, smem, stream>>>(
            bottom_data, spatial_scale, height, width, channels, pooled_height, pooled_width,
            bottom_rois, top_data, argmax_data);
    return scda_launch_status();
}

SCDA_API int ROIPoolBackwardLaucher(const float *top_diff, const float spatial_scale,
                                    const int batch_size, const int num_rois, const int height,
                                    const int width, const int channels, const int pooled_height,
                                    const int pooled_width, const float *bottom_rois,
                                    float *bottom_diff, const int *argmax_data,
                                    cudaStream_t stream)
{
    (void)spatial_scale;
    (void)bottom_rois;  // the argmax already encodes which input element each bin took
    if (batch_size <= 0 || num_rois < 0 || height <= 0 || width <= 0 || channels <= 0 ||
        pooled_height <= 0 || pooled_width <= 0 || !top_diff || !bottom_diff || !argmax_data)
        return 0;
    cudaError_t e = cudaMemsetAsync(bottom_diff, 0,
                                    sizeof(float) * (size_t)batch_size * channels * height * width,
                                    stream);
    if (e != cudaSuccess) return -(int)e;
    const long long total = (long long)num_rois * channels * pooled_height * pooled_width;
    if (total == 0) return 1;
    const bool vec = total % 4 == 0 && ((uintptr_t)top_diff % 16 == 0) &&
                     ((uintptr_t)argmax_data % 16 == 0);
    const int per = vec ? 4 : 1;
    long long want = (total / per + 255) / 256;
    const int grid = (int)(want < (long long)kNumSMs * 16 ? want : (long long)kNumSMs * 16);
    if (vec)
        roi_pool_bwd_scatter_kernel<true><<<grid, 256, 0, stream>>>(top_diff, argmax_data, total,
                                                                    bottom_diff);
    else
        roi_pool_bwd_scatter_kernel<false><<<grid, 256, 0, stream>>>(top_diff, argmax_data, total,
                                                                     bottom_diff);
    return scda_launch_status();
}
