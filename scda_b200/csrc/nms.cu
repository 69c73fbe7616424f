// Greedy NMS on pre-sorted boxes, wholly on device, for sm_100a.
//
// Replaces the reference's gpu_nms (extensions/_nms/src/nms_cuda.c:17-67):
//   nms_kernel (src/cuda/nms_kernel.cu:26-70, all ceil(N/64)^2 tiles on 64-thread
//   CTAs) -> cudaMemcpy of the N x ceil(N/64) x 8 B mask to the host (18 MB at
//   N = 12 000) -> sequential host OR-scan (nms_cuda.c:47-58).
// Here:
//   1. nms_mask_kernel: upper-triangle tiles only; 256-thread CTAs cover a
//      64-row block x 4 column blocks; column boxes are staged in shared
//      memory with their +1 widths/heights precomputed; each thread builds its
//      64-bit word in registers.
//      For the scan the words are stored column-block-major (mask[b][i]) so that
//      step 2 reads them with coalesced loads.
//   2. nms_scan_kernel: one CTA walks the 64-box blocks in order: a coalesced
//      pull of the words of already-kept boxes gives the block's removed set, a
//      warp-parallel fixpoint settles the 64 boxes (see the kernel), and the kept
//      bits are expanded to ascending indices at the end.  Indices and the count
//      are written to device memory; nothing crosses PCIe.
//
// Bit-exactness: survivors depend on `IoU > thresh` comparisons, so the IoU
// must round exactly as the reference kernel's does.  nvcc contracts the
// reference's `Sa + Sb - interS` (Sb being a product) into
// fma(wb, hb, Sa) - interS and keeps an IEEE division (PTX of the unmodified
// file, nvcc 12.9); the intrinsics below pin that sequence.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kColsPerCta = 4;
constexpr int kMaskThreads = kTile * kColsPerCta;

struct ColBox {
    float x1, y1, x2, y2, w, h, pad0, pad1;  // 32 B: two LDS.128 broadcasts
};

__device__ __forceinline__ bool suppresses(float ax1, float ay1, float ax2, float ay2, float Sa,
                                           const ColBox &b, float thresh)
{
    const float left = fmaxf(ax1, b.x1), right = fminf(ax2, b.x2);
    const float top = fmaxf(ay1, b.y1), bottom = fminf(ay2, b.y2);
    const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
    const float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
    const float inter = __fmul_rn(w, h);
    const float den = __fsub_rn(__fmaf_rn(b.w, b.h, Sa), inter);
    // The verdict must be that of the IEEE division.  Most pairs do not need it: disjoint boxes
    // (inter == 0: 0 / den > thresh is false for any den when thresh >= 0), and pairs whose
    // approximate quotient (2 ulp) is further than 1e-6 relative from the threshold.
    if (thresh >= 0.f && den > 0.f) {
        if (!(inter > 0.f)) return false;
        const float q = __fdividef(inter, den);
        if (q > thresh * 1.000001f) return true;
        if (q < thresh * 0.999999f) return false;
    }
    return __fdiv_rn(inter, den) > thresh;
}

// grid = (ceil(cb / kColsPerCta), cb): blockIdx.y = row block, blockIdx.x = group
// of column blocks.
//   kRowMajor = true : mask[i * cb + colb], every word written (zero below the
//                      diagonal) — the reference's N x cb layout, for _nms callers.
//   kRowMajor = false: mask[colb * n + i], diagonal and above only — the layout the
//                      device scan pulls from with coalesced loads.
template <bool kRowMajor>
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(int n_cap, const int *__restrict__ n_dev, float thresh,
                const float *__restrict__ boxes, unsigned long long *__restrict__ mask, int cb_cap)
{
    __shared__ ColBox s_col[kColsPerCta][kTile];
    // live box count: the capacity, or a device-resident count below it
    const int n = n_dev ? max(0, min(n_cap, __ldg(n_dev))) : n_cap;
    const int cb = (n + kTile - 1) / kTile;
    const int rb = blockIdx.y;
    if (rb >= cb || blockIdx.x * kColsPerCta >= cb) return;
    const int sub = threadIdx.x / kTile, lane = threadIdx.x % kTile;
    const int colb = blockIdx.x * kColsPerCta + sub;
    // whole CTA below the diagonal: nothing to compute
    if (!kRowMajor && (blockIdx.x + 1) * kColsPerCta - 1 < rb) return;

    const bool active = colb < cb && colb >= rb;
    if (active) {
        const int j = colb * kTile + lane;
        ColBox c = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (j < n) {
            const float *p = boxes + 5ll * j;
            c.x1 = p[0]; c.y1 = p[1]; c.x2 = p[2]; c.y2 = p[3];
            c.w = __fadd_rn(__fsub_rn(c.x2, c.x1), 1.f);
            c.h = __fadd_rn(__fsub_rn(c.y2, c.y1), 1.f);
        }
        s_col[sub][lane] = c;
    }
    __syncthreads();

    const int i = rb * kTile + lane;
    if (i >= n || colb >= cb) return;
    unsigned long long word = 0;
    if (active) {
        const float *a = boxes + 5ll * i;
        const float ax1 = a[0], ay1 = a[1], ax2 = a[2], ay2 = a[3];
        const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(ax2, ax1), 1.f),
                                   __fadd_rn(__fsub_rn(ay2, ay1), 1.f));
        const int cols = min(n - colb * kTile, kTile);
        const int start = (rb == colb) ? lane + 1 : 0;
#pragma unroll 4
        for (int j = start; j < cols; ++j)
            if (suppresses(ax1, ay1, ax2, ay2, Sa, s_col[sub][j], thresh)) word |= 1ull << j;
    } else if (!kRowMajor) {
        return;
    }
    if (kRowMajor)
        mask[(long long)i * cb_cap + colb] = word;
    else
        mask[(long long)colb * n_cap + i] = word;
}

constexpr int kScanThreads = 1024;

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v)
{
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

// Single CTA.  For block b (boxes 64b .. 64b+63):
//   pull   : removed = OR over kept boxes i < 64b of mask[b][i]   (all threads,
//            coalesced predicated loads, warp redux + one shared round)
//   resolve: warp 0 settles the 64 boxes of the block with a parallel fixpoint on
//            the diagonal words — a box with no earlier undecided overlapper is
//            kept, what the newly kept suppress is removed — 4 redux per round,
//            rounds = longest suppression chain inside the block.
//   The kept bits live in shared memory; indices are expanded at the end
//   (ordered compaction by popcount prefix), truncated to max_keep.
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long *__restrict__ mask, int n_cap,
                const int *__restrict__ n_dev, int max_keep, long long *__restrict__ keep_out,
                long long *__restrict__ num_out)
{
    extern __shared__ unsigned long long s_kept[];  // ceil(n_cap / 64) words
    const int n = n_dev ? max(0, min(n_cap, __ldg(n_dev))) : n_cap;
    const int cb = (n + kTile - 1) / kTile;
    __shared__ unsigned long long s_red[kScanThreads / 32];
    __shared__ int s_prefix[kScanThreads / 32];
    __shared__ int s_total, s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) { s_total = 0; s_last = cb; }
    __syncthreads();

    for (int b = 0; b < cb; ++b) {
        const unsigned long long *__restrict__ col = mask + (long long)b * n_cap;
        const int before = b * kTile;
        const int rows = min(n - before, kTile);
        // diagonal words of this block: independent of the pull, issue them first
        unsigned long long d0 = 0, d1 = 0;
        if (warp == 0) {
            if (lane < rows) d0 = __ldg(col + before + lane);
            if (lane + 32 < rows) d1 = __ldg(col + before + lane + 32);
        }
        unsigned long long acc = 0;
        for (int i0 = tid; i0 < before; i0 += kScanThreads * 4) {
            unsigned long long v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kScanThreads;
                v[u] = 0;
                if (i < before && ((s_kept[i >> 6] >> (i & 63)) & 1ull)) v[u] = __ldg(col + i);
            }
            acc |= v[0] | v[1] | v[2] | v[3];
        }
        acc = warp_or64(acc);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long removed = warp_or64(s_red[lane]);
            const unsigned long long valid = rows == 64 ? ~0ull : ((1ull << rows) - 1);
            unsigned long long und = ~removed & valid, kept = 0;
            while (und) {
                const unsigned long long mine =
                    (((und >> lane) & 1ull) ? d0 : 0ull) | (((und >> (lane + 32)) & 1ull) ? d1 : 0ull);
                const unsigned long long contested = warp_or64(mine);
                const unsigned long long now = und & ~contested;   // never empty: lowest bit of und
                kept |= now;
                const unsigned long long hit =
                    (((now >> lane) & 1ull) ? d0 : 0ull) | (((now >> (lane + 32)) & 1ull) ? d1 : 0ull);
                und &= ~now & ~warp_or64(hit);
            }
            if (lane == 0) {
                s_kept[b] = kept;
                const int total = s_total + __popcll(kept);
                s_total = total;
                if (max_keep > 0 && total >= max_keep) s_last = b + 1;
            }
        }
        __syncthreads();
        if (s_last <= b + 1) break;
    }

    // ordered expansion of the kept bits into indices
    const int nb = min(s_last, cb);
    __syncthreads();
    int base = 0;   // running count of survivors before the current chunk of words
    for (int w0 = 0; w0 < nb; w0 += kScanThreads) {
        const int w = w0 + tid;
        const unsigned long long word = w < nb ? s_kept[w] : 0ull;
        const int cnt = __popcll(word);
        int incl = cnt;   // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_prefix[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = s_prefix[lane];
            int inc2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += t;
            }
            s_prefix[lane] = inc2 - v;                 // exclusive warp offsets
            if (lane == 31) s_red[0] = (unsigned long long)inc2;  // chunk total
        }
        __syncthreads();
        int pos = base + s_prefix[warp] + incl - cnt;
        unsigned long long bits = word;
        while (bits) {
            const int bit = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (max_keep <= 0 || pos < max_keep) keep_out[pos] = (long long)w * kTile + bit;
            ++pos;
        }
        base += (int)s_red[0];
        __syncthreads();
    }
    if (tid == 0) *num_out = (max_keep > 0 && base > max_keep) ? max_keep : base;
}

// Scan, wide form (n_cap <= kWideCap): the chain above costs one L2 round trip + two barriers per 64 boxes
// (188 steps at 12 000 boxes: 160-210 us, as long as the mask kernel).  Here a step settles kWideW * 64 = 256
// boxes: the pull walks a LIST of the boxes kept so far (ascending indices in shared memory: at most a couple of
// thousand entries however many boxes were scanned, one round of independent loads) for the step's four column
// blocks at once, and warp 0 runs the same fixpoint on the 256 x 256 diagonal (staged in shared memory; a lane owns 8 rows).
// The list doubles as the output: it is the ascending survivor indices.
constexpr int kWideW = 4;
constexpr int kWideCap = 32768;

__global__ void __launch_bounds__(kScanThreads)
nms_scan_wide_kernel(const unsigned long long *__restrict__ mask, int n_cap, const int *__restrict__ n_dev,
                     int max_keep, long long *__restrict__ keep_out, long long *__restrict__ num_out)
{
    extern __shared__ int s_list[];                  // [n_cap] kept indices, ascending
    __shared__ unsigned long long s_red[kScanThreads / 32][kWideW];
    __shared__ unsigned long long s_d[kWideW][kWideW * kTile];
    __shared__ int s_total, s_done;
    static_assert(kScanThreads == kWideW * kWideW * kTile, "one diagonal word per thread");
    const int n = n_dev ? max(0, min(n_cap, __ldg(n_dev))) : n_cap;
    const int cb = (n + kTile - 1) / kTile;
    const int steps = (cb + kWideW - 1) / kWideW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kRows = kWideW * kTile;            // 256
    constexpr int kPerLane = kRows / 32;             // 8 rows per lane: lane + 32 q
    if (tid == 0) { s_total = 0; s_done = 0; }
    __syncthreads();

    for (int s = 0; s < steps; ++s) {
        const int before = s * kRows;
        const int cb0 = s * kWideW;                  // first column block of the step
        // diagonal words of the step's rows: s_d[w][r] = mask[cb0 + w][before + r], zero below the diagonal;
        // one word per thread, consumed by warp 0 after the barrier below
        {
            const int w = tid >> 8, r = tid & (kRows - 1);
            const int i = before + r;
            unsigned long long v = 0ull;
            if (i < n && w >= (r >> 6) && cb0 + w < cb) v = __ldg(mask + (long long)(cb0 + w) * n_cap + i);
            s_d[w][r] = v;
        }
        // pull: OR of the kept boxes' words for the step's column blocks
        unsigned long long acc[kWideW];
#pragma unroll
        for (int w = 0; w < kWideW; ++w) acc[w] = 0ull;
        const int nk = s_total;
        for (int k0 = tid; k0 < nk; k0 += 2 * kScanThreads) {
            unsigned long long v[2][kWideW];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = k0 + u * kScanThreads;
                const int i = k < nk ? s_list[k] : -1;
#pragma unroll
                for (int w = 0; w < kWideW; ++w)
                    v[u][w] = (i >= 0 && cb0 + w < cb) ? __ldg(mask + (long long)(cb0 + w) * n_cap + i) : 0ull;
            }
#pragma unroll
            for (int w = 0; w < kWideW; ++w) acc[w] |= v[0][w] | v[1][w];
        }
#pragma unroll
        for (int w = 0; w < kWideW; ++w) {
            acc[w] = warp_or64(acc[w]);
            if (lane == 0) s_red[warp][w] = acc[w];
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long und[kWideW], kept[kWideW];
#pragma unroll
            for (int w = 0; w < kWideW; ++w) {
                const unsigned long long removed = warp_or64(s_red[lane][w]);
                const int rows = min(max(n - (before + w * kTile), 0), kTile);
                const unsigned long long valid = rows == 64 ? ~0ull : ((1ull << rows) - 1);
                und[w] = ~removed & valid;
                kept[w] = 0ull;
            }
            for (;;) {
                unsigned long long any = 0ull;
#pragma unroll
                for (int w = 0; w < kWideW; ++w) any |= und[w];
                if (!any) break;
                // words of the undecided rows: what they would suppress if kept
                unsigned long long con[kWideW];
#pragma unroll
                for (int w = 0; w < kWideW; ++w) con[w] = 0ull;
#pragma unroll
                for (int q = 0; q < kPerLane; ++q) {
                    const bool u = (und[q >> 1] >> (lane + 32 * (q & 1))) & 1ull;
#pragma unroll
                    for (int w = 0; w < kWideW; ++w) con[w] |= u ? s_d[w][lane + 32 * q] : 0ull;
                }
                unsigned long long now[kWideW];
#pragma unroll
                for (int w = 0; w < kWideW; ++w) {
                    now[w] = und[w] & ~warp_or64(con[w]);      // no earlier undecided overlapper: kept
                    kept[w] |= now[w];
                }
                unsigned long long hit[kWideW];
#pragma unroll
                for (int w = 0; w < kWideW; ++w) hit[w] = 0ull;
#pragma unroll
                for (int q = 0; q < kPerLane; ++q) {
                    const bool u = (now[q >> 1] >> (lane + 32 * (q & 1))) & 1ull;
#pragma unroll
                    for (int w = 0; w < kWideW; ++w) hit[w] |= u ? s_d[w][lane + 32 * q] : 0ull;
                }
#pragma unroll
                for (int w = 0; w < kWideW; ++w) und[w] &= ~now[w] & ~warp_or64(hit[w]);
            }
            // ordered append of the step's survivors
            int base = s_total;
#pragma unroll
            for (int w = 0; w < kWideW; ++w) {
#pragma unroll
                for (int hbit = 0; hbit < 2; ++hbit) {
                    const int bit = lane + 32 * hbit;
                    if ((kept[w] >> bit) & 1ull)
                        s_list[base + __popcll(kept[w] & ((1ull << bit) - 1ull))] = before + w * kTile + bit;
                }
                base += __popcll(kept[w]);
            }
            if (lane == 0) {
                s_total = base;
                if (max_keep > 0 && base >= max_keep) s_done = 1;
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    const int total = (max_keep > 0 && s_total > max_keep) ? max_keep : s_total;
    for (int k = tid; k < total; k += kScanThreads) keep_out[k] = (long long)s_list[k];
    if (tid == 0) *num_out = total;
}

int g_nms_wide_scan = -1;

int launch_mask(int n, const int *n_dev, const float *boxes, unsigned long long *mask, float thresh,
                bool row_major, cudaStream_t stream)
{
    const int cb = ceil_div(n, kTile);
    dim3 grid(ceil_div(cb, kColsPerCta), cb);
    if (row_major)
        nms_mask_kernel<true><<<grid, kMaskThreads, 0, stream>>>(n, n_dev, thresh, boxes, mask, cb);
    else
        nms_mask_kernel<false><<<grid, kMaskThreads, 0, stream>>>(n, n_dev, thresh, boxes, mask, cb);
    return scda_launch_status();
}

int nms_impl(int n, const int *n_dev, const float *boxes, float thresh, int max_keep,
             int64_t *keep_out, int64_t *num_out, void *workspace, size_t workspace_bytes,
             cudaStream_t stream)
{
    if (n < 0 || !num_out) return 0;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int64_t), stream);
        return e == cudaSuccess ? 1 : -(int)e;
    }
    if (!boxes || !keep_out || !workspace || workspace_bytes < scda_nms_workspace_bytes(n))
        return 0;
    const int cb = ceil_div(n, kTile);
    const size_t smem = sizeof(unsigned long long) * (size_t)cb;
    if (smem > 200 * 1024) return 0;  // n <= ~1.6 M boxes
    unsigned long long *mask = (unsigned long long *)workspace;
    int st = launch_mask(n, n_dev, boxes, mask, thresh, false, stream);
    if (st != 1) return st;
    if (g_nms_wide_scan < 0) {
        const char *e = getenv("SCDA_NMS_WIDE");
        g_nms_wide_scan = (e && *e == '0') ? 0 : 1;
    }
    if (n <= kWideCap && g_nms_wide_scan) {
        const size_t smem_w = sizeof(int) * (size_t)n;
        static size_t attr = 0;
        if (smem_w > 32 * 1024 && smem_w > attr) {     // + ~10 KB of static shared memory
            cudaError_t e = cudaFuncSetAttribute(nms_scan_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_w);
            if (e != cudaSuccess) return -(int)e;
            attr = smem_w;
        }
        nms_scan_wide_kernel<<<1, kScanThreads, smem_w, stream>>>(mask, n, n_dev, max_keep, (long long *)keep_out,
                                                                  (long long *)num_out);
        return scda_launch_status();
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
    }
    nms_scan_kernel<<<1, kScanThreads, smem, stream>>>(mask, n, n_dev, max_keep,
                                                       (long long *)keep_out,
                                                       (long long *)num_out);
    return scda_launch_status();
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Many small independent NMS problems in ONE launch: group g = n_dev[g] (<= n_cap <= 1024) boxes sorted by
// descending score, boxes[g][n_cap][5].  The reference's compute_predicted_bboxes
// (functions/predict_bbox.py:36-52) runs one gpu_nms round trip (mask kernel, 18-line host scan, two PCIe
// copies) per class per image: 8 per image.  One CTA per group: the boxes and the whole suppression bitmask
// live in shared memory (n_cap 1024: 32 KB + 128 KB), the greedy scan walks it with one warp.
// Same IoU arithmetic and `> thresh` verdict as nms_mask_kernel, hence the same survivors as scda_nms per group.
constexpr int kGroupThreads = 256;

__global__ void __launch_bounds__(kGroupThreads)
nms_groups_kernel(int n_cap, const int *__restrict__ n_dev, const float *__restrict__ boxes, float thresh,
                  long long *__restrict__ keep, long long *__restrict__ num_out)
{
    extern __shared__ __align__(16) unsigned char g_smem[];
    const int g = blockIdx.x;
    const int n = max(0, min(n_cap, n_dev[g]));
    const int words = (n_cap + 63) / 64;
    ColBox *s_box = reinterpret_cast<ColBox *>(g_smem);
    unsigned long long *s_mask = reinterpret_cast<unsigned long long *>(s_box + n_cap);
    unsigned long long *s_removed = s_mask + (size_t)n_cap * words;
    const float *gb = boxes + (long long)g * n_cap * 5;
    for (int j = threadIdx.x; j < n; j += kGroupThreads) {
        ColBox c;
        c.x1 = gb[5 * j]; c.y1 = gb[5 * j + 1]; c.x2 = gb[5 * j + 2]; c.y2 = gb[5 * j + 3];
        c.w = __fadd_rn(__fsub_rn(c.x2, c.x1), 1.f);
        c.h = __fadd_rn(__fsub_rn(c.y2, c.y1), 1.f);
        c.pad0 = c.pad1 = 0.f;
        s_box[j] = c;
    }
    for (int w = threadIdx.x; w < words; w += kGroupThreads) s_removed[w] = 0ull;
    __syncthreads();
    // (row i, word w): bits of the boxes j > i that box i suppresses
    for (int t = threadIdx.x; t < n * words; t += kGroupThreads) {
        const int i = t / words, w = t - i * words;
        unsigned long long word = 0ull;
        if (w * 64 + 63 > i) {
            const ColBox a = s_box[i];
            const float Sa = __fmul_rn(a.w, a.h);
            const int j0 = max(w * 64, i + 1), j1 = min(n, w * 64 + 64);
            for (int j = j0; j < j1; ++j)
                if (suppresses(a.x1, a.y1, a.x2, a.y2, Sa, s_box[j], thresh)) word |= 1ull << (j - w * 64);
        }
        s_mask[(size_t)i * words + w] = word;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int kept = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned long long rem = s_removed[i >> 6];          // same value in every lane
            if ((rem >> (i & 63)) & 1ull) continue;
            if (lane == 0) keep[(long long)g * n_cap + kept] = i;
            ++kept;
            for (int w = lane; w < words; w += 32) s_removed[w] |= s_mask[(size_t)i * words + w];
            __syncwarp();
        }
        if (lane == 0) num_out[g] = kept;
    }
}

static size_t scda_nms_groups_smem_bytes(int n_cap)
{
    const size_t words = (size_t)(n_cap + 63) / 64;
    return (size_t)n_cap * sizeof(ColBox) + ((size_t)n_cap * words + words) * sizeof(unsigned long long);
}

SCDA_API int scda_nms_groups(int groups, int n_cap, const int *n_dev, const float *boxes, float thresh,
                             int64_t *keep_out, int64_t *num_out, cudaStream_t stream)
{
    if (groups < 0 || n_cap <= 0 || n_cap > 1024 || !n_dev || !boxes || !keep_out || !num_out) return 0;
    if (groups == 0) return 1;
    const size_t smem = scda_nms_groups_smem_bytes(n_cap);
    if (smem > 220 * 1024) return 0;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(nms_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr = smem;
    }
    nms_groups_kernel<<<groups, kGroupThreads, smem, stream>>>(n_cap, n_dev, boxes, thresh, (long long *)keep_out,
                                                               (long long *)num_out);
    return scda_launch_status();
}

SCDA_API void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev,
                   float nms_overlap_thresh)
{
    if (boxes_num <= 0 || !boxes_dev || !mask_dev) return;
    launch_mask(boxes_num, nullptr, boxes_dev, mask_dev, nms_overlap_thresh, true, (cudaStream_t)0);
}

SCDA_API int scda_nms_mask(int n, const float *boxes, unsigned long long *mask, float thresh,
                           cudaStream_t stream)
{
    if (n < 0 || (n > 0 && (!boxes || !mask))) return 0;
    if (n == 0) return 1;
    return launch_mask(n, nullptr, boxes, mask, thresh, true, stream);
}

SCDA_API size_t scda_nms_workspace_bytes(int n)
{
    if (n <= 0) return 0;
    const size_t cb = (size_t)ceil_div(n, kTile);
    return (size_t)n * cb * sizeof(unsigned long long);
}

SCDA_API int scda_nms(int n, const float *boxes, float thresh, int max_keep, int64_t *keep_out,
                      int64_t *num_out, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream)
{
    return nms_impl(n, nullptr, boxes, thresh, max_keep, keep_out, num_out, workspace,
                    workspace_bytes, stream);
}

SCDA_API int scda_nms_dyn(int n_cap, const int *n_dev, const float *boxes, float thresh,
                          int max_keep, int64_t *keep_out, int64_t *num_out, void *workspace,
                          size_t workspace_bytes, cudaStream_t stream)
{
    if (!n_dev) return 0;
    return nms_impl(n_cap, n_dev, boxes, thresh, max_keep, keep_out, num_out, workspace,
                    workspace_bytes, stream);
}
