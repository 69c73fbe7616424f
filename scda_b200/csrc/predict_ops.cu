// Final detections of one image at test time: compute_predicted_bboxes of the reference
// (functions/predict_bbox.py:13-66) around ONE batched NMS launch.
//
// Reference, per class and image: numpy decode of the class's deltas (utils/bbox_helper.py:88-103 in float64,
// de-normalised by the precomputed stds / means, :29-31), clip to the image (:105-111), score threshold, argsort,
// H2D, GPU NMS mask, D2H, host scan; then a numpy top-n over all classes.  The tensor-op form of the same steps
// (scda_b200/functions/predict_bbox.py, kept for batches of several images) is ~60 eager launches around
// scda_nms_groups: 0.9 ms of launch overhead for a few thousand boxes (profiles/r2_final_config2.json).  Here:
//   1. predict_prepare_kernel, one CTA per foreground class: score threshold -> (key, RoI) pairs -> bitonic sort
//      by descending score (ties by RoI index) -> float64 decode + clip of the sorted RoIs -> dets [C-1][n][5]
//      (the layout scda_nms_groups reads) + the live count per class;
//   2. scda_nms_groups (csrc/nms.cu);
//   3. predict_topn_kernel, one CTA: the survivors of all classes ranked by descending score (rank = number of
//      survivors that beat it, ties by class then position: one binary search per class, no sort) -> rows
//      [batch, x1, y1, x2, y2, score, class] of the best top_n + their count.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kPT = 256;
constexpr int kMaxRois = 1024;

__global__ void __launch_bounds__(kPT)
predict_prepare_kernel(int n, int num_classes, const float *__restrict__ rois, int roi_stride,
                       const float *__restrict__ cls, const float *__restrict__ loc, int normalize, double s0,
                       double s1, double s2, double s3, double m0, double m1, double m2, double m3, double img_h,
                       double img_w, float score_thresh, float *__restrict__ dets, int *__restrict__ n_live)
{
    __shared__ float s_key[kMaxRois];
    __shared__ int s_idx[kMaxRois];
    __shared__ int s_cnt;
    const int c = blockIdx.x + 1, tid = threadIdx.x;
    int P = 1;
    while (P < n) P <<= 1;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int live = 0;
    for (int i = tid; i < P; i += kPT) {
        float key = -FLT_MAX;                                  // padding sorts behind everything
        if (i < n) {
            const float s = cls[(long long)i * num_classes + c];
            const bool ok = score_thresh > 0.f ? s > score_thresh : true;
            key = ok ? s : -1.0f;
            live += ok;
        }
        s_key[i] = key;
        s_idx[i] = i;
    }
    if (live) atomicAdd(&s_cnt, live);
    __syncthreads();
    // bitonic sort, descending key, ascending RoI index on equal keys
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += kPT) {
                const int p = i ^ j;
                if (p > i) {
                    const float ka = s_key[i], kb = s_key[p];
                    const int ia = s_idx[i], ib = s_idx[p];
                    const bool a_first = ka > kb || (ka == kb && ia < ib);      // a belongs in front of b
                    const bool up = (i & k) == 0;                               // this run is in final order
                    if (up ? !a_first : a_first) {
                        s_key[i] = kb; s_key[p] = ka;
                        s_idx[i] = ib; s_idx[p] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) n_live[blockIdx.x] = s_cnt;
    float *out = dets + (long long)blockIdx.x * n * 5;
    for (int r = tid; r < n; r += kPT) {
        const int i = s_idx[r];
        const float *rb = rois + (long long)i * roi_stride + 1;
        const float x1 = rb[0], y1 = rb[1], x2 = rb[2], y2 = rb[3];
        const float *d = loc + ((long long)i * num_classes + c) * 4;
        double d0 = (double)d[0], d1 = (double)d[1], d2 = (double)d[2], d3 = (double)d[3];
        if (normalize) {
            d0 = __dadd_rn(__dmul_rn(d0, s0), m0);
            d1 = __dadd_rn(__dmul_rn(d1, s1), m1);
            d2 = __dadd_rn(__dmul_rn(d2, s2), m2);
            d3 = __dadd_rn(__dmul_rn(d3, s3), m3);
        }
        const double bw = (double)__fsub_rn(x2, x1), bh = (double)__fsub_rn(y2, y1);
        const double cx = (double)(__fadd_rn(x1, x2) / 2.f), cy = (double)(__fadd_rn(y1, y2) / 2.f);
        const double ncx = __dadd_rn(__dmul_rn(d0, bw), cx), ncy = __dadd_rn(__dmul_rn(d1, bh), cy);
        const double nw = __dmul_rn(exp(d2), bw), nh = __dmul_rn(exp(d3), bh);
        const double bx1 = fmin(fmax(__dsub_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double by1 = fmin(fmax(__dsub_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const double bx2 = fmin(fmax(__dadd_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double by2 = fmin(fmax(__dadd_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        float *o = out + (long long)r * 5;
        o[0] = (float)bx1; o[1] = (float)by1; o[2] = (float)bx2; o[3] = (float)by2; o[4] = s_key[r];
    }
}

constexpr int kTT = 1024;

__global__ void __launch_bounds__(kTT)
predict_topn_kernel(int groups, int n, const float *__restrict__ dets, const long long *__restrict__ keep,
                    const long long *__restrict__ n_keep, float batch_ix, int top_n, float *__restrict__ rows,
                    int *__restrict__ count)
{
    extern __shared__ unsigned char smem_raw[];
    float *s_score = reinterpret_cast<float *>(smem_raw);                 // [groups * n]
    int *s_code = reinterpret_cast<int *>(s_score + (size_t)groups * n);  // class << 16 | position in its kept list
    __shared__ int s_off[65];
    const int tid = threadIdx.x;
    if (tid == 0) {
        int t = 0;
        for (int g = 0; g < groups; ++g) {
            s_off[g] = t;
            long long k = n_keep[g];
            t += (int)(k < 0 ? 0 : (k > n ? n : k));
        }
        s_off[groups] = t;
    }
    __syncthreads();
    const int total = s_off[groups];
    for (int g = 0; g < groups; ++g) {
        const int cnt = s_off[g + 1] - s_off[g];
        for (int j = tid; j < cnt; j += kTT) {
            const long long r = keep[(long long)g * n + j];
            s_score[s_off[g] + j] = dets[((long long)g * n + r) * 5 + 4];
            s_code[s_off[g] + j] = (g << 16) | j;
        }
    }
    __syncthreads();
    const int lim = top_n < total ? top_n : total;
    if (tid == 0) count[0] = lim;
    for (int t = tid; t < total; t += kTT) {
        const float st = s_score[t];
        const int ct = s_code[t];
        // every class's survivors are in descending score order (NMS keeps the order of its input): the number
        // of them that beat candidate t is a binary search per class — classes in front of t's win ties, classes
        // behind it do not, inside its own class the position decides
        const int gt_ = ct >> 16;
        int rank = ct & 0xffff;
        for (int g = 0; g < groups; ++g) {
            if (g == gt_) continue;
            int lo = s_off[g], hi = s_off[g + 1];
            const int base = lo;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float sm = s_score[mid];
                const bool beats = g < gt_ ? sm >= st : sm > st;
                if (beats) lo = mid + 1; else hi = mid;
            }
            rank += lo - base;
        }
        if (rank < lim) {
            const int g = ct >> 16, j = ct & 0xffff;
            const long long r = keep[(long long)g * n + j];
            const float *d = dets + ((long long)g * n + r) * 5;
            float *o = rows + (long long)rank * 7;
            o[0] = batch_ix; o[1] = d[0]; o[2] = d[1]; o[3] = d[2]; o[4] = d[3]; o[5] = d[4]; o[6] = (float)(g + 1);
        }
    }
}

}  // namespace

SCDA_API int scda_predict_prepare(int n, int num_classes, const float *rois, int roi_stride, const float *cls,
                                  const float *loc, int normalize, const double *stds, const double *means,
                                  double img_h, double img_w, float score_thresh, float *dets, int *n_live,
                                  cudaStream_t stream)
{
    if (n <= 0 || n > kMaxRois || num_classes < 2 || num_classes > 65 || roi_stride < 5) return 0;
    if (!rois || !cls || !loc || !dets || !n_live || (normalize && (!stds || !means))) return 0;
    const double one[4] = {1.0, 1.0, 1.0, 1.0}, zero[4] = {0.0, 0.0, 0.0, 0.0};
    const double *s = normalize ? stds : one, *m = normalize ? means : zero;
    predict_prepare_kernel<<<num_classes - 1, kPT, 0, stream>>>(n, num_classes, rois, roi_stride, cls, loc,
                                                                normalize ? 1 : 0, s[0], s[1], s[2], s[3], m[0], m[1],
                                                                m[2], m[3], img_h, img_w, score_thresh, dets, n_live);
    return scda_launch_status();
}

SCDA_API int scda_predict_topn(int groups, int n, const float *dets, const int64_t *keep, const int64_t *n_keep,
                               float batch_ix, int top_n, float *rows, int *count, cudaStream_t stream)
{
    if (groups <= 0 || groups > 64 || n <= 0 || n > kMaxRois || top_n <= 0) return 0;
    if (!dets || !keep || !n_keep || !rows || !count) return 0;
    const size_t smem = (size_t)groups * n * (sizeof(float) + sizeof(int));
    if (smem > 200 * 1024) return 0;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(predict_topn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr = smem;
    }
    predict_topn_kernel<<<1, kTT, smem, stream>>>(groups, n, dets, (const long long *)keep, (const long long *)n_keep,
                                                  batch_ix, top_n, rows, count);
    return scda_launch_status();
}
