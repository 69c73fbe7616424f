// Pairwise IoU matrices in the reference's two "no +1" conventions, sm_100a.
//
//   IOUOverlap         replaces IOUOverlapKernel
//                      (extensions/_bbox_helper/src/cuda/iou_overlap_kernel.cu:33-65):
//                      union clamped to >= 1; nvcc contracts area1 + area2 into
//                      fma(w1, h1, area2) in the unmodified source.
//   scda_bbox_overlaps replaces the host-side cython_bbox.bbox_overlaps
//                      (extensions/_cython_bbox/cython_bbox.pyx:32-73): zero unless
//                      both overlaps are strictly positive, no clamp, plain
//                      (uncontracted) float arithmetic as gcc emits it on x86-64.
//
// Both are output-write bound (n1 x n2 x 4 B).  The query boxes are staged in
// shared memory once per CTA; consecutive threads write consecutive columns.
#include "common.cuh"

namespace {

constexpr int kIouThreads = 256;
constexpr int kQueryTile = 256;  // query boxes staged per pass

template <bool kCython>
__global__ void __launch_bounds__(kIouThreads)
iou_kernel(const float *__restrict__ b1, const float *__restrict__ b2, int stride, int n1, int n2,
           float *__restrict__ out, int rows_per_cta)
{
    __shared__ float4 s_q[kQueryTile];
    __shared__ float s_qarea[kQueryTile];
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(row0 + rows_per_cta, n1);
    for (int q0 = 0; q0 < n2; q0 += kQueryTile) {
        const int qn = min(kQueryTile, n2 - q0);
        __syncthreads();
        for (int j = threadIdx.x; j < qn; j += kIouThreads) {
            const float *p = b2 + (long long)(q0 + j) * stride;
            const float4 q = make_float4(p[0], p[1], p[2], p[3]);
            s_q[j] = q;
            s_qarea[j] = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
        }
        __syncthreads();
        const int work = (row1 - row0) * qn;
        for (int t = threadIdx.x; t < work; t += kIouThreads) {
            const int i = row0 + t / qn, j = t % qn;
            const float *a = b1 + (long long)i * stride;
            const float ax1 = __ldg(a), ay1 = __ldg(a + 1), ax2 = __ldg(a + 2), ay2 = __ldg(a + 3);
            const float4 q = s_q[j];
            float v;
            if (kCython) {
                v = 0.f;
                const float iw = __fsub_rn(fminf(ax2, q.z), fmaxf(ax1, q.x));
                if (iw > 0) {
                    const float ih = __fsub_rn(fminf(ay2, q.w), fmaxf(ay1, q.y));
                    if (ih > 0) {
                        const float area = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
                        const float inter = __fmul_rn(iw, ih);
                        const float ua = __fsub_rn(__fadd_rn(area, s_qarea[j]), inter);
                        v = __fdiv_rn(inter, ua);
                    }
                }
            } else {
                const float w = fmaxf(__fsub_rn(fminf(ax2, q.z), fmaxf(ax1, q.x)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(ay2, q.w), fmaxf(ay1, q.y)), 0.f);
                const float inter = __fmul_rn(w, h);
                const float uni = fmaxf(
                    __fsub_rn(__fmaf_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1), s_qarea[j]), inter),
                    1.f);
                v = __fdiv_rn(inter, uni);
            }
            out[(long long)i * n2 + q0 + j] = v;
        }
    }
}

// One IoU of box (ax1, ay1, ax2, ay2) with query q in either convention.
template <bool kCython>
__device__ __forceinline__ float iou_one(float ax1, float ay1, float ax2, float ay2, float4 q)
{
    const float qarea = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
    if (kCython) {
        float v = 0.f;
        const float iw = __fsub_rn(fminf(ax2, q.z), fmaxf(ax1, q.x));
        if (iw > 0) {
            const float ih = __fsub_rn(fminf(ay2, q.w), fmaxf(ay1, q.y));
            if (ih > 0) {
                const float area = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
                const float inter = __fmul_rn(iw, ih);
                const float ua = __fsub_rn(__fadd_rn(area, qarea), inter);
                v = __fdiv_rn(inter, ua);
            }
        }
        return v;
    }
    const float w = fmaxf(__fsub_rn(fminf(ax2, q.z), fmaxf(ax1, q.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(ay2, q.w), fmaxf(ay1, q.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float uni = fmaxf(__fsub_rn(__fmaf_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1), qarea), inter), 1.f);
    return __fdiv_rn(inter, uni);
}

// Main form: a thread owns four consecutive columns of one row.  Its own box and its four query boxes are
// five INDEPENDENT loads (the queries are the same few addresses for every thread: L1 broadcasts), so a
// launch costs one memory latency instead of the three dependent ones of the staged form above (query boxes
// -> barrier -> own box -> store), which is what an 8-column anchors x GT matrix is made of; the four results
// leave as one 16-byte store, consecutive lanes = consecutive addresses.
template <bool kCython, bool kVecIn, bool kVecOut>
__global__ void __launch_bounds__(kIouThreads)
iou_quad_kernel(const float *__restrict__ b1, const float *__restrict__ b2, int stride, int n1, int n2,
                float *__restrict__ out, int nq4, int tasks)
{
    for (int t = blockIdx.x * kIouThreads + threadIdx.x; t < tasks; t += gridDim.x * kIouThreads) {
        const int i = t / nq4, j0 = (t - i * nq4) << 2;
        float4 a;
        if (kVecIn) {
            a = __ldg(reinterpret_cast<const float4 *>(b1) + i);
        } else {
            const float *p = b1 + (long long)i * stride;
            a = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
        }
        float4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = min(j0 + k, n2 - 1);
            if (kVecIn) {
                q[k] = __ldg(reinterpret_cast<const float4 *>(b2) + j);
            } else {
                const float *p = b2 + (long long)j * stride;
                q[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
            }
        }
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = iou_one<kCython>(a.x, a.y, a.z, a.w, q[k]);
        float *o = out + (long long)i * n2 + j0;
        if (kVecOut) {
            *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (j0 + k < n2) o[k] = v[k];
        }
    }
}

template <bool kCython>
int launch_iou(const float *b1, const float *b2, int stride, int n1, int n2, float *out,
               cudaStream_t stream)
{
    if (n1 < 0 || n2 < 0 || stride < 4) return 0;
    if (n1 == 0 || n2 == 0) return 1;
    if (!b1 || !b2 || !out) return 0;
    const int nq4 = (n2 + 3) / 4;
    if ((long long)n1 * nq4 < (1ll << 30)) {
        const int tasks = n1 * nq4;
        int grid = ceil_div(tasks, kIouThreads);
        if (grid > kNumSMs * 8) grid = kNumSMs * 8;
        const bool vin = stride == 4 && (uintptr_t)b1 % 16 == 0 && (uintptr_t)b2 % 16 == 0;
        const bool vout = n2 % 4 == 0 && (uintptr_t)out % 16 == 0;
#define SCDA_IOU_LAUNCH(VI, VO) \
        iou_quad_kernel<kCython, VI, VO><<<grid, kIouThreads, 0, stream>>>(b1, b2, stride, n1, n2, out, nq4, tasks)
        if (vin && vout) SCDA_IOU_LAUNCH(true, true);
        else if (vin) SCDA_IOU_LAUNCH(true, false);
        else if (vout) SCDA_IOU_LAUNCH(false, true);
        else SCDA_IOU_LAUNCH(false, false);
#undef SCDA_IOU_LAUNCH
        return scda_launch_status();
    }
    // enough CTAs for ~4 per SM, at least 8 rows each
    int rows = ceil_div(n1, kNumSMs * 4);
    if (rows < 8) rows = 8;
    iou_kernel<kCython><<<ceil_div(n1, rows), kIouThreads, 0, stream>>>(b1, b2, stride, n1, n2, out,
                                                                       rows);
    return scda_launch_status();
}

}  // namespace

SCDA_API int IOUOverlap(const float *bboxes1_data, const float *bboxes2_data, const int size_bbox,
                        const int num_bbox1, const int num_bbox2, float *top_data,
                        cudaStream_t stream)
{
    return launch_iou<false>(bboxes1_data, bboxes2_data, size_bbox, num_bbox1, num_bbox2, top_data,
                             stream);
}

SCDA_API int scda_bbox_overlaps(int n, const float *boxes, int k, const float *query, float *out,
                                cudaStream_t stream)
{
    return launch_iou<true>(boxes, query, 4, n, k, out, stream);
}
