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
//   2. nms_scan_kernel: one CTA walks the 64-box blocks in order.  Within a
//      block one thread resolves the survivors from the diagonal words with
//      ffs jumps (work ~ survivors, not 64); then all threads OR the survivors'
//      mask rows into the shared `remv` words of the later blocks, each thread
//      owning its columns.  Indices and the count are written to device
//      memory; nothing crosses PCIe.
//
// Bit-exactness: survivors depend on `IoU > thresh` comparisons, so the IoU
// must round exactly as the reference kernel's does.  nvcc contracts the
// reference's `Sa + Sb - interS` (Sb being a product) into
// fma(wb, hb, Sa) - interS and keeps an IEEE division (PTX of the unmodified
// file, nvcc 12.9); the intrinsics below pin that sequence.
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
    return __fdiv_rn(inter, den) > thresh;
}

// grid = (ceil(cb / kColsPerCta), cb): blockIdx.y = row block, blockIdx.x = group
// of column blocks.  kFull: also write (zero) words below the diagonal so the
// output is a complete N x cb matrix as the reference's callers expect.
template <bool kFull>
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(int n, float thresh, const float *__restrict__ boxes,
                unsigned long long *__restrict__ mask, int cb)
{
    __shared__ ColBox s_col[kColsPerCta][kTile];
    const int rb = blockIdx.y;
    const int sub = threadIdx.x / kTile, lane = threadIdx.x % kTile;
    const int colb = blockIdx.x * kColsPerCta + sub;
    // whole CTA below the diagonal: nothing to compute
    if (!kFull && (blockIdx.x + 1) * kColsPerCta - 1 < rb) return;

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
    } else if (!kFull) {
        return;
    }
    mask[(long long)i * cb + colb] = word;
}

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long *__restrict__ mask, int n, int cb, int max_keep,
                long long *__restrict__ keep_out, long long *__restrict__ num_out)
{
    extern __shared__ unsigned long long s_remv[];  // cb words
    __shared__ unsigned long long s_diag[kTile];
    __shared__ int s_kept[kTile];
    __shared__ int s_nkept, s_total, s_done;

    for (int j = threadIdx.x; j < cb; j += kScanThreads) s_remv[j] = 0;
    if (threadIdx.x == 0) { s_total = 0; s_done = 0; }
    __syncthreads();

    for (int blk = 0; blk < cb; ++blk) {
        if (threadIdx.x < kTile) {
            const int i = blk * kTile + threadIdx.x;
            s_diag[threadIdx.x] = i < n ? __ldg(mask + (long long)i * cb + blk) : 0ull;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int rows = min(n - blk * kTile, kTile);
            const unsigned long long valid = rows == 64 ? ~0ull : ((1ull << rows) - 1);
            unsigned long long removed = s_remv[blk];
            unsigned long long avail = ~removed & valid;
            int k = 0, total = s_total;
            while (avail) {
                const int b = __ffsll((long long)avail) - 1;
                s_kept[k++] = b;
                keep_out[total++] = (long long)blk * kTile + b;
                if (max_keep > 0 && total >= max_keep) { s_done = 1; break; }
                removed |= s_diag[b] | (1ull << b);
                avail &= ~removed;
            }
            s_nkept = k;
            s_total = total;
        }
        __syncthreads();
        if (s_done) break;
        const int k = s_nkept;
        if (k > 0) {
            for (int j = blk + 1 + threadIdx.x; j < cb; j += kScanThreads) {
                unsigned long long acc = 0;
                int t = 0;
                for (; t + 4 <= k; t += 4) {
                    const unsigned long long a0 = __ldg(mask + (long long)(blk * kTile + s_kept[t]) * cb + j);
                    const unsigned long long a1 = __ldg(mask + (long long)(blk * kTile + s_kept[t + 1]) * cb + j);
                    const unsigned long long a2 = __ldg(mask + (long long)(blk * kTile + s_kept[t + 2]) * cb + j);
                    const unsigned long long a3 = __ldg(mask + (long long)(blk * kTile + s_kept[t + 3]) * cb + j);
                    acc |= a0 | a1 | a2 | a3;
                }
                for (; t < k; ++t) acc |= __ldg(mask + (long long)(blk * kTile + s_kept[t]) * cb + j);
                s_remv[j] |= acc;
            }
        }
        // the next iteration's first __syncthreads orders these writes before
        // thread 0 reads s_remv[blk + 1]
    }
    __syncthreads();
    if (threadIdx.x == 0) *num_out = s_total;
}

int launch_mask(int n, const float *boxes, unsigned long long *mask, float thresh, bool full,
                cudaStream_t stream)
{
    const int cb = ceil_div(n, kTile);
    dim3 grid(ceil_div(cb, kColsPerCta), cb);
    if (full)
        nms_mask_kernel<true><<<grid, kMaskThreads, 0, stream>>>(n, thresh, boxes, mask, cb);
    else
        nms_mask_kernel<false><<<grid, kMaskThreads, 0, stream>>>(n, thresh, boxes, mask, cb);
    return scda_launch_status();
}

}  // namespace

SCDA_API void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev,
                   float nms_overlap_thresh)
{
    if (boxes_num <= 0 || !boxes_dev || !mask_dev) return;
    launch_mask(boxes_num, boxes_dev, mask_dev, nms_overlap_thresh, true, (cudaStream_t)0);
}

SCDA_API int scda_nms_mask(int n, const float *boxes, unsigned long long *mask, float thresh,
                           cudaStream_t stream)
{
    if (n < 0 || (n > 0 && (!boxes || !mask))) return 0;
    if (n == 0) return 1;
    return launch_mask(n, boxes, mask, thresh, true, stream);
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
    int st = launch_mask(n, boxes, mask, thresh, false, stream);
    if (st != 1) return st;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
    }
    nms_scan_kernel<<<1, kScanThreads, smem, stream>>>(mask, n, cb, max_keep,
                                                       (long long *)keep_out,
                                                       (long long *)num_out);
    return scda_launch_status();
}
