// Training targets of the RCNN head as ONE kernel: compute_proposal_targets of the reference
// (functions/proposal_target.py:17-177), per image.
//
// Reference: numpy on the host — append the ground-truth boxes to the proposals (:42-43), clip, cython IoU
// R x G (:49), positives > 0.5 / negatives in [lo, hi) (:63-76), np.random.choice down to 128 positives and
// the rest negatives (:99-111), class-specific box targets normalised by the precomputed means / stds
// (:135-144), padding to the batch size by resampling (:149-155), four H2D copies.  The tensor-op form of the
// same steps (functions/proposal_target.py here, kept as the fallback for shapes beyond this kernel's shared
// memory) is ~150 launches of 2-3 us + two library radix sorts on the critical path of the detector forward.
//
// One CTA (R <= 4096 boxes: the work is a few hundred thousand operations, all of it latency):
//   IoU with every ground truth (the cython arithmetic of csrc/iou.cu, operation by operation) -> row maximum
//   and first arg-maximum -> positive / negative flags -> ordered ranks by block scan -> key-driven draws ->
//   padding -> encode -> write.
// Random draws follow functions/_sampling.py: candidate j (j-th set entry in ascending index order) owns
// keys[j]; "choose k of n" = the k candidates with the smallest keys, in key order (a bitonic sort of
// (key, j) pairs in shared memory, skipped when nothing has to be dropped).
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kTT = 1024;
constexpr int kMaxR = 4096;
constexpr int kMaxG = 256;

struct TargetParams {
    int cap, ldb, G, append_gts, bs, nc, want_pos, normalize;
    float img_h, img_w, pos_thresh, neg_hi, neg_lo, batch_ix;
    double mean[4], stdv[4];
};

template <int kThreads>
__device__ __forceinline__ int block_scan_excl(int v, int *s_w, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < kThreads / 32 ? s_w[lane] : 0;
        int inc2 = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += t;
        }
        s_w[lane] = inc2 - w;
        if (lane == 31) s_w[32] = inc2;
    }
    __syncthreads();
    *total = s_w[32];
    return s_w[warp] + incl - v;
}

// order-preserving 64-bit key of a double
__device__ __forceinline__ unsigned long long sortable64(double d)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// cython_bbox.pyx:44-72 (as csrc/iou.cu, kCython)
__device__ __forceinline__ float iou_cython(float4 a, float4 q)
{
    float v = 0.f;
    const float iw = __fsub_rn(fminf(a.z, q.z), fmaxf(a.x, q.x));
    if (iw > 0) {
        const float ih = __fsub_rn(fminf(a.w, q.w), fmaxf(a.y, q.y));
        if (ih > 0) {
            const float area = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
            const float qarea = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
            const float inter = __fmul_rn(iw, ih);
            v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area, qarea), inter));
        }
    }
    return v;
}

// `choose` of functions/_sampling.py for the members listed in s_mem[0 .. count): at most `want` of them into
// s_sel[0 .. min(count, want)) — all of them in ascending order when count <= want, else the `want` members
// with the smallest keys in key order.  s_key / s_ix: scratch of >= next_pow2(count) entries.
__device__ void choose_members(const int *s_mem, int count, int want, const double *__restrict__ keys,
                               unsigned long long *s_key, int *s_ix, int *s_sel, int sel_cap)
{
    const int tid = threadIdx.x;
    if (count <= want) {
        for (int j = tid; j < count && j < sel_cap; j += kTT) s_sel[j] = s_mem[j];
        __syncthreads();
        return;
    }
    int N = 2;
    while (N < count) N <<= 1;
    for (int j = tid; j < N; j += kTT) {
        s_key[j] = j < count ? sortable64(keys[j]) : ~0ull;
        s_ix[j] = j;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (N >> 1); t += kTT) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                const int hi = lo | j;
                const bool up = (lo & k) == 0;                          // ascending half
                const unsigned long long a = s_key[lo], b = s_key[hi];
                const int ia = s_ix[lo], ib = s_ix[hi];
                const bool gt = a > b || (a == b && ia > ib);
                if (gt == up) {
                    s_key[lo] = b; s_key[hi] = a;
                    s_ix[lo] = ib; s_ix[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < want && r < sel_cap; r += kTT) s_sel[r] = s_mem[s_ix[r]];
    __syncthreads();
}

__global__ void __launch_bounds__(kTT)
proposal_targets_kernel(TargetParams p, const float *__restrict__ boxes, const long long *__restrict__ n_boxes,
                        const float *__restrict__ gts, const double *__restrict__ keys_pos,
                        const double *__restrict__ keys_neg, const double *__restrict__ keys_pad,
                        float *__restrict__ out_rois, long long *__restrict__ out_labels,
                        float *__restrict__ out_loc_t, float *__restrict__ out_loc_w)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const int R = p.append_gts ? p.cap + p.G : p.cap;
    int Npad = 2;
    while (Npad < R) Npad <<= 1;
    float4 *s_roi = reinterpret_cast<float4 *>(t_smem);                       // [R]
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_roi + R);   // [Npad]
    int *s_ix = reinterpret_cast<int *>(s_key + Npad);                         // [Npad]
    int *s_mem_pos = s_ix + Npad;                                              // [R]
    int *s_mem_neg = s_mem_pos + R;                                            // [R]
    int *s_sel_pos = s_mem_neg + R;                                            // [bs]
    int *s_sel_neg = s_sel_pos + p.bs;                                         // [bs]
    short *s_am = reinterpret_cast<short *>(s_sel_neg + p.bs);                 // [R]
    __shared__ float4 s_gt[kMaxG];
    __shared__ float s_gt_label[kMaxG];
    __shared__ unsigned char s_gt_ok[kMaxG];
    __shared__ int s_w[33];
    const int tid = threadIdx.x;
    const long long nb = max(0ll, min((long long)p.cap, n_boxes[0]));

    for (int g = tid; g < p.G; g += kTT) {
        const float *q = gts + 5 * g;
        const float4 b = make_float4(q[0], q[1], q[2], q[3]);
        s_gt[g] = b;
        s_gt_label[g] = q[4];
        s_gt_ok[g] = (b.z > __fadd_rn(b.x, 1.f)) && (b.w > __fadd_rn(b.y, 1.f));
    }
    __syncthreads();
    const float wmax = __fsub_rn(p.img_w, 1.f), hmax = __fsub_rn(p.img_h, 1.f);
    // per thread: boxes tid, tid + 1024, ... (ascending inside a thread is not needed: ranks come from scans
    // over CONTIGUOUS ownership below, so flags go to shared memory first)
    unsigned char *s_flag = reinterpret_cast<unsigned char *>(s_am + R + (R & 1));   // [R]: 1 pos, 2 neg
    for (int i = tid; i < R; i += kTT) {
        float4 b;
        bool live;
        if (i < p.cap) {
            const float *q = boxes + (long long)i * p.ldb;
            b = make_float4(q[0], q[1], q[2], q[3]);
            live = i < nb;
        } else {
            b = s_gt[i - p.cap];
            live = s_gt_ok[i - p.cap];
        }
        b.x = fminf(fmaxf(b.x, 0.f), wmax);
        b.y = fminf(fmaxf(b.y, 0.f), hmax);
        b.z = fminf(fmaxf(b.z, 0.f), wmax);
        b.w = fminf(fmaxf(b.w, 0.f), hmax);
        s_roi[i] = b;
        float mx = -FLT_MAX;
        int am = 0;
        for (int g = 0; g < p.G; ++g) {
            const float v = s_gt_ok[g] ? iou_cython(b, s_gt[g]) : -1.f;
            if (v > mx) { mx = v; am = g; }
        }
        s_am[i] = (short)am;
        const bool pos = live && mx > p.pos_thresh;
        const bool neg = live && mx < p.neg_hi && mx >= p.neg_lo && !pos;
        s_flag[i] = pos ? 1 : (neg ? 2 : 0);
    }
    __syncthreads();
    // member lists in ascending index order: thread t owns boxes [t * per, (t + 1) * per)
    const int per = (R + kTT - 1) / kTT;
    const int i0 = min(R, tid * per), i1 = min(R, i0 + per);
    int cp = 0, cn = 0;
    for (int i = i0; i < i1; ++i) {
        cp += s_flag[i] == 1;
        cn += s_flag[i] == 2;
    }
    int count_pos, count_neg;
    int op = block_scan_excl<kTT>(cp, s_w, &count_pos);
    int on = block_scan_excl<kTT>(cn, s_w, &count_neg);
    for (int i = i0; i < i1; ++i) {
        if (s_flag[i] == 1) s_mem_pos[op++] = i;
        if (s_flag[i] == 2) s_mem_neg[on++] = i;
    }
    __syncthreads();
    const int n_pos = min(count_pos, p.want_pos);
    choose_members(s_mem_pos, count_pos, p.want_pos, keys_pos, s_key, s_ix, s_sel_pos, p.bs);
    const int want_neg = p.bs - n_pos;
    const int n_neg = min(count_neg, want_neg);
    choose_members(s_mem_neg, count_neg, want_neg, keys_neg, s_key, s_ix, s_sel_neg, p.bs);
    const int total = n_pos + n_neg;

    const int row = p.nc * 4;
    for (int r = tid; r < p.bs; r += kTT) {
        int src = r;
        if (r >= total) {
            // padding by resampling rows [0, total) (:149-155): floor(u * total), as the key-driven draw
            const double u = keys_pad[r - total];
            long long rep = (long long)floor(__dmul_rn(u, (double)total));
            if (rep > p.bs - 1) rep = p.bs - 1;
            if (rep < 0) rep = 0;
            src = (int)rep;
        }
        const bool have = total > 0 && src < total;
        const bool is_pos = have && src < n_pos;
        int ri = 0;
        if (have) ri = is_pos ? s_sel_pos[src] : s_sel_neg[src - n_pos];
        const float4 b = s_roi[ri];
        const int g = s_am[ri];
        const float4 q = s_gt[g];
        float *ro = out_rois + 5 * r;
        ro[0] = p.batch_ix; ro[1] = b.x; ro[2] = b.y; ro[3] = b.z; ro[4] = b.w;
        const long long label = is_pos ? (long long)(int)s_gt_label[g] : 0ll;
        out_labels[r] = label;
        float *lt = out_loc_t + (long long)r * row, *lw = out_loc_w + (long long)r * row;
        for (int k = 0; k < row; ++k) { lt[k] = 0.f; lw[k] = 0.f; }
        if (is_pos) {
            // utils/bbox_helper.py:60-85 in float32, then (t - mean) / std in float64 (:135-144)
            const float bw = __fsub_rn(b.z, b.x), bh = __fsub_rn(b.w, b.y);
            const float bx = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f), by = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
            const float gw = __fsub_rn(q.z, q.x), gh = __fsub_rn(q.w, q.y);
            const float gx = __fmul_rn(__fadd_rn(q.x, q.z), 0.5f), gy = __fmul_rn(__fadd_rn(q.y, q.w), 0.5f);
            float t[4] = {__fdiv_rn(__fsub_rn(gx, bx), bw), __fdiv_rn(__fsub_rn(gy, by), bh),
                          logf(__fdiv_rn(gw, bw)), logf(__fdiv_rn(gh, bh))};
            int cls = (int)label;
            cls = cls < 0 ? 0 : (cls > p.nc - 1 ? p.nc - 1 : cls);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double v = (double)t[k];
                if (p.normalize) v = __ddiv_rn(__dsub_rn(v, p.mean[k]), p.stdv[k]);
                lt[cls * 4 + k] = (float)v;
                lw[cls * 4 + k] = 1.f;
            }
        }
    }
}

size_t targets_smem(int R, int bs)
{
    int Npad = 2;
    while (Npad < R) Npad <<= 1;
    return sizeof(float4) * (size_t)R + 12 * (size_t)Npad + sizeof(int) * (2 * (size_t)R + 2 * (size_t)bs) +
           2 * (size_t)(R + (R & 1)) + (size_t)R + 16;
}

}  // namespace

SCDA_API int scda_proposal_targets(int cap, int ldb, const float *boxes, const long long *n_boxes, int G,
                                   const float *gts, int append_gts, float img_h, float img_w, float pos_thresh,
                                   float neg_hi, float neg_lo, int want_pos, int batch_size, int num_classes,
                                   int normalize, const double *means4, const double *stds4, float batch_ix,
                                   const double *keys_pos, const double *keys_neg, const double *keys_pad,
                                   float *rois, long long *labels, float *loc_targets, float *loc_weights,
                                   cudaStream_t stream)
{
    if (cap <= 0 || ldb < 4 || G <= 0 || G > kMaxG || batch_size <= 0 || num_classes <= 0 || want_pos < 0) return 0;
    if (!boxes || !n_boxes || !gts || !keys_pos || !keys_neg || !keys_pad || !rois || !labels || !loc_targets ||
        !loc_weights || (normalize && (!means4 || !stds4)))
        return 0;
    const int R = append_gts ? cap + G : cap;
    if (R > kMaxR) return 0;
    const size_t smem = targets_smem(R, batch_size);
    if (smem > 200 * 1024) return 0;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(proposal_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr = smem;
    }
    TargetParams p;
    p.cap = cap; p.ldb = ldb; p.G = G; p.append_gts = append_gts ? 1 : 0; p.bs = batch_size; p.nc = num_classes;
    p.want_pos = want_pos; p.normalize = normalize ? 1 : 0;
    p.img_h = img_h; p.img_w = img_w; p.pos_thresh = pos_thresh; p.neg_hi = neg_hi; p.neg_lo = neg_lo;
    p.batch_ix = batch_ix;
    for (int k = 0; k < 4; ++k) {
        p.mean[k] = normalize ? means4[k] : 0.0;
        p.stdv[k] = normalize ? stds4[k] : 1.0;
    }
    proposal_targets_kernel<<<1, kTT, smem, stream>>>(p, boxes, n_boxes, gts, keys_pos, keys_neg, keys_pad, rois,
                                                     labels, loc_targets, loc_weights);
    return scda_launch_status();
}
