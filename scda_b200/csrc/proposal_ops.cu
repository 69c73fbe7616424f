// RPN proposal decode: the stretch of compute_rpn_proposals between the top-k and the NMS
// (functions/rpn_proposal.py:53-64 + utils/bbox_helper.py:88-111 of the reference) as ONE kernel:
//   gather anchors / deltas by the top-k order -> decode (centre / size form, widths without +1,
//   float32 exp, float64 products: the dtypes numpy gives the reference) -> clip to
//   [0, w-1] x [0, h-1] -> drop boxes with w+1 or h+1 below roi_min_size -> stable compaction of
//   the survivors -> float32 rows (x1, y1, x2, y2, score) for the NMS, zero rows behind them, and
//   the survivor count on the device.
// The tensor-op form of the same arithmetic takes ~35 launches per image on the critical path of
// the forward; this is one single-CTA launch (12 000 rows: latency, not bandwidth).
// Each float64 product / sum is a separate rounding (__dmul_rn / __dadd_rn), as in the reference's
// sequence of numpy operations: no FMA contraction, so the float32 rows the NMS sees are the same.
#include "common.cuh"

namespace {

constexpr int kPT = 1024;

__global__ void __launch_bounds__(kPT)
rpn_decode_pack_kernel(int pre, const double *__restrict__ anchors, const float *__restrict__ deltas,
                       const long long *__restrict__ order, const float *__restrict__ top, double img_h,
                       double img_w, double min_size, float *__restrict__ packed, int *__restrict__ count)
{
    __shared__ int s_warp[kPT / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (pre + kPT - 1) / kPT;          // consecutive rows per thread: order is preserved
    const int i0 = tid * per, i1 = min(pre, i0 + per);
    // pass 1: count survivors of this thread's rows
    int mine = 0;
    for (int i = i0; i < i1; ++i) {
        const long long a = order[i];
        const double ax1 = anchors[4 * a], ay1 = anchors[4 * a + 1], ax2 = anchors[4 * a + 2], ay2 = anchors[4 * a + 3];
        const double w = __dsub_rn(ax2, ax1), h = __dsub_rn(ay2, ay1);
        const double cx = __dadd_rn(ax1, ax2) / 2.0, cy = __dadd_rn(ay1, ay2) / 2.0;
        const float d0 = deltas[4 * a], d1 = deltas[4 * a + 1], d2 = deltas[4 * a + 2], d3 = deltas[4 * a + 3];
        const double ncx = __dadd_rn(__dmul_rn((double)d0, w), cx), ncy = __dadd_rn(__dmul_rn((double)d1, h), cy);
        const double nw = __dmul_rn((double)expf(d2), w), nh = __dmul_rn((double)expf(d3), h);
        const double x1 = fmin(fmax(__dsub_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y1 = fmin(fmax(__dsub_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const double x2 = fmin(fmax(__dadd_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y2 = fmin(fmax(__dadd_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const bool ok = __dadd_rn(__dsub_rn(x2, x1), 1.0) >= min_size && __dadd_rn(__dsub_rn(y2, y1), 1.0) >= min_size;
        mine += ok ? 1 : 0;
    }
    // exclusive scan of the per-thread counts
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane];
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += u;
        }
        s_warp[lane] = inc2 - v;                     // exclusive warp offsets
        if (lane == 31) s_total = inc2;
    }
    __syncthreads();
    int pos = s_warp[warp] + incl - mine;
    const int total = s_total;
    // pass 2: recompute (cheaper than parking 12 rows x 5 floats per thread) and scatter
    for (int i = i0; i < i1; ++i) {
        const long long a = order[i];
        const double ax1 = anchors[4 * a], ay1 = anchors[4 * a + 1], ax2 = anchors[4 * a + 2], ay2 = anchors[4 * a + 3];
        const double w = __dsub_rn(ax2, ax1), h = __dsub_rn(ay2, ay1);
        const double cx = __dadd_rn(ax1, ax2) / 2.0, cy = __dadd_rn(ay1, ay2) / 2.0;
        const float d0 = deltas[4 * a], d1 = deltas[4 * a + 1], d2 = deltas[4 * a + 2], d3 = deltas[4 * a + 3];
        const double ncx = __dadd_rn(__dmul_rn((double)d0, w), cx), ncy = __dadd_rn(__dmul_rn((double)d1, h), cy);
        const double nw = __dmul_rn((double)expf(d2), w), nh = __dmul_rn((double)expf(d3), h);
        const double x1 = fmin(fmax(__dsub_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y1 = fmin(fmax(__dsub_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const double x2 = fmin(fmax(__dadd_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y2 = fmin(fmax(__dadd_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const bool ok = __dadd_rn(__dsub_rn(x2, x1), 1.0) >= min_size && __dadd_rn(__dsub_rn(y2, y1), 1.0) >= min_size;
        if (ok) {
            float *r = packed + 5ll * pos;
            r[0] = (float)x1; r[1] = (float)y1; r[2] = (float)x2; r[3] = (float)y2; r[4] = top[i];
            ++pos;
        }
    }
    // zero rows behind the survivors
    for (long long k = (long long)total * 5 + tid; k < (long long)pre * 5; k += kPT) packed[k] = 0.f;
    if (tid == 0) count[0] = total;
}

}  // namespace

SCDA_API int scda_rpn_decode_pack(int pre, const double *anchors, const float *deltas, const long long *order,
                                  const float *top_scores, double img_h, double img_w, double min_size,
                                  float *packed, int *count, cudaStream_t stream)
{
    if (pre <= 0 || !anchors || !deltas || !order || !top_scores || !packed || !count) return 0;
    rpn_decode_pack_kernel<<<1, kPT, 0, stream>>>(pre, anchors, deltas, order, top_scores, img_h, img_w, min_size,
                                                  packed, count);
    return scda_launch_status();
}

// ---------------------------------------------------------------------------------------------------
// Crops around the cluster centres: `get_corner_from_center` + the slicing loop of the reference driver
// (tools/faster_rcnn_train_val.py:411-438, 528-557).  The branchy corner rule is a clamp of
// int(c) - R/2 to [0, size - R]; centres stay on the device.  out[k, c, y, x] = image[c, y1_k + y, x1_k + x].
namespace {

__global__ void __launch_bounds__(256)
crop_regions_kernel(const float *__restrict__ image, const float *__restrict__ centers, float *__restrict__ out,
                    int K, int C, int H, int W, int R)
{
    const long long total = (long long)K * C * R * R;
    const int half = R / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % R);
        long long t = i / R;
        const int y = (int)(t % R);
        t /= R;
        const int c = (int)(t % C), k = (int)(t / C);
        long long x1 = (long long)centers[2 * k] - half, y1 = (long long)centers[2 * k + 1] - half;
        x1 = x1 < 0 ? 0 : (x1 > W - R ? W - R : x1);
        y1 = y1 < 0 ? 0 : (y1 > H - R ? H - R : y1);
        out[i] = __ldg(image + ((long long)c * H + (y1 + y)) * W + (x1 + x));
    }
}

}  // namespace

SCDA_API int scda_crop_regions(int K, int C, int H, int W, int R, const float *image, const float *centers,
                               float *out, cudaStream_t stream)
{
    if (K <= 0 || C <= 0 || R <= 0 || R > H || R > W || (R & 1) || !image || !centers || !out) return 0;
    const long long total = (long long)K * C * R * R;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    crop_regions_kernel<<<(unsigned)blocks, 256, 0, stream>>>(image, centers, out, K, C, H, W, R);
    return scda_launch_status();
}
