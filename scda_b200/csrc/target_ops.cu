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

// ---------------------------------------------------------------------------------------------------
// RPN anchor targets as ONE kernel: compute_anchor_targets of the reference (functions/anchor_target.py:16-116)
// for one image.  Reference: numpy on the host — cython IoU of the K*A anchors with the ground truth (:51),
// per-anchor maximum / arg-maximum, per-ground-truth maximum with ties kept and maxima below 0.1 dropped
// (:59-65), labels -1 / 0 / 1 (:69-79), np.random.choice sub-sampling to 128 positives and 256 in all (:82-94),
// encode without +1 widths in float64 (utils/bbox_helper.py:60-85), three H2D copies.  The tensor-op form
// (functions/anchor_target.py here, kept for batches of several images) is ~150 launches beside the backbone.
//
// One CTA of 1024 threads; thread t owns the anchors [t * per, (t + 1) * per) so that block scans give ordered
// ranks.  The IoU row of an anchor (G <= 256 values) is recomputed in each of the two passes that need it
// instead of being stored.  Draws follow functions/_sampling.py `drop`: candidate j (j-th member in ascending
// index order) owns keys[j]; the (count - keep) members with the SMALLEST keys are removed (ties: lower rank
// first) — an 8-pass radix selection of the cut-off key over keys[0 .. count), no sort.
namespace {

constexpr int kAT = 1024;

struct AnchorParams {
    int KA, A, fh, fw, G, want_pos, batch_total;
    float neg_thresh, pos_thresh, gt_floor;
};

// exact cut of `drop`: the n_remove smallest (key, rank) pairs among keys[0 .. count).  Returns the cut-off key T
// (sortable form), *n_lt = number of keys < T; the first (n_remove - *n_lt) ranks with key == T are removed too.
__device__ unsigned long long select_smallest(const double *__restrict__ keys, int count, int n_remove, int *s_hist,
                                              int *s_misc, int *n_lt)
{
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned long long prefix = 0ull, mask = 0ull;
    int need = n_remove, below = 0;                  // need: rank (from the bottom, 1-based) inside the matching keys
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int d = tid; d < 256; d += kAT) s_hist[d] = 0;
        __syncthreads();
        for (int j = tid; j < count; j += kAT) {
            const unsigned long long k = sortable64(keys[j]);
            if ((k & mask) == prefix) atomicAdd(&s_hist[(int)((k >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            int mine = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) mine += s_hist[8 * lane + q];
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
            const int first = __ffs(hit) - 1;
            if (lane == first) {
                int under = incl - mine;
                int d = 8 * lane;
                for (int q = 0; q < 8; ++q, ++d) {
                    const int h = s_hist[d];
                    if (under + h >= need) break;
                    under += h;
                }
                s_misc[0] = d;
                s_misc[1] = under;
            }
        }
        __syncthreads();
        const int d = s_misc[0], under = s_misc[1];
        prefix |= (unsigned long long)d << shift;
        mask |= 255ull << shift;
        need -= under;
        below += under;
        __syncthreads();
    }
    *n_lt = below;
    return prefix;
}

__global__ void __launch_bounds__(kAT)
anchor_targets_kernel(AnchorParams p, const float *__restrict__ anchors32, const double *__restrict__ anchors64,
                      const float *__restrict__ gts, const double *__restrict__ keys_pos,
                      const double *__restrict__ keys_neg, long long *__restrict__ cls_targets,
                      float *__restrict__ loc_targets, float *__restrict__ loc_masks,
                      long long *__restrict__ normalizer)
{
    extern __shared__ __align__(16) unsigned char a_smem[];
    signed char *s_lab = reinterpret_cast<signed char *>(a_smem);                   // [KA]
    unsigned char *s_arg = reinterpret_cast<unsigned char *>(s_lab + p.KA);         // [KA] matched ground truth
    __shared__ float4 s_gt[kMaxG];
    __shared__ int s_gtmax[kMaxG];                   // float bits of the per-ground-truth maximum (IoU >= 0)
    __shared__ int s_hist[256];
    __shared__ int s_w[33];
    __shared__ int s_misc[2];
    const int tid = threadIdx.x;
    for (int g = tid; g < p.G; g += kAT) {
        const float *q = gts + 5 * g;
        s_gt[g] = make_float4(q[0], q[1], q[2], q[3]);
        s_gtmax[g] = 0;
    }
    __syncthreads();
    const int per = (p.KA + kAT - 1) / kAT;
    const int i0 = min(p.KA, tid * per), i1 = min(p.KA, i0 + per);
    // pass 1: per-ground-truth maxima
    for (int i = i0; i < i1; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(anchors32) + i);
        for (int g = 0; g < p.G; ++g) {
            const float v = iou_cython(a, s_gt[g]);
            if (v > 0.f) atomicMax(&s_gtmax[g], __float_as_int(v));
        }
    }
    __syncthreads();
    // pass 2: labels
    int cp = 0, cn = 0;
    for (int i = i0; i < i1; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(anchors32) + i);
        float mx = -FLT_MAX;
        int am = 0, last_hit = -1;
        for (int g = 0; g < p.G; ++g) {
            const float v = iou_cython(a, s_gt[g]);
            if (v > mx) { mx = v; am = g; }
            float gm = __int_as_float(s_gtmax[g]);
            if (gm < p.gt_floor) gm = -1.f;
            if (v == gm) last_hit = g;               // duplicates resolve to the largest g (:62-65)
        }
        int lab = -1;
        if (mx < p.neg_thresh) lab = 0;
        if (last_hit >= 0) lab = 1;
        if (mx > p.pos_thresh) lab = 1;
        s_lab[i] = (signed char)lab;
        s_arg[i] = (unsigned char)(last_hit >= 0 ? last_hit : am);
        cp += lab == 1;
        cn += lab == 0;
    }
    int count_pos, count_neg;
    int base_pos = block_scan_excl<kAT>(cp, s_w, &count_pos);
    // positives: keep at most want_pos
    int n_pos = count_pos;
    if (count_pos > p.want_pos) {
        int n_lt;
        const int n_remove = count_pos - p.want_pos;
        const unsigned long long T = select_smallest(keys_pos, count_pos, n_remove, s_hist, s_misc, &n_lt);
        // ties on T: the first (n_remove - n_lt) of them in rank order go
        int eq_mine = 0, r = base_pos;
        for (int i = i0; i < i1; ++i)
            if (s_lab[i] == 1) { eq_mine += sortable64(keys_pos[r]) == T; ++r; }
        int eq_total;
        int eq_before = block_scan_excl<kAT>(eq_mine, s_w, &eq_total);
        const int quota = n_remove - n_lt;
        r = base_pos;
        for (int i = i0; i < i1; ++i)
            if (s_lab[i] == 1) {
                const unsigned long long k = sortable64(keys_pos[r]);
                if (k < T || (k == T && eq_before++ < quota)) { s_lab[i] = -1; }
                ++r;
            }
        n_pos = p.want_pos;
    }
    int base_neg = block_scan_excl<kAT>(cn, s_w, &count_neg);
    const int want_neg = p.batch_total - n_pos;
    int n_neg = count_neg;
    if (count_neg > want_neg) {
        int n_lt;
        const int n_remove = count_neg - want_neg;
        const unsigned long long T = select_smallest(keys_neg, count_neg, n_remove, s_hist, s_misc, &n_lt);
        int eq_mine = 0, r = base_neg;
        for (int i = i0; i < i1; ++i)
            if (s_lab[i] == 0) { eq_mine += sortable64(keys_neg[r]) == T; ++r; }
        int eq_total;
        int eq_before = block_scan_excl<kAT>(eq_mine, s_w, &eq_total);
        const int quota = n_remove - n_lt;
        r = base_neg;
        for (int i = i0; i < i1; ++i)
            if (s_lab[i] == 0) {
                const unsigned long long k = sortable64(keys_neg[r]);
                if (k < T || (k == T && eq_before++ < quota)) { s_lab[i] = -1; }
                ++r;
            }
        n_neg = want_neg;
    }
    __syncthreads();
    if (tid == 0) normalizer[0] = max(1, n_pos + n_neg);
    // outputs in their [A, fh, fw] / [4A, fh, fw] order: consecutive threads write consecutive addresses
    const int plane = p.fh * p.fw;
    for (int o = tid; o < p.A * plane; o += kAT) {
        const int a = o / plane, cell = o - a * plane;
        const int i = cell * p.A + a;
        const int lab = s_lab[i];
        cls_targets[o] = (long long)lab;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (lab == 1) {
            // utils/bbox_helper.py:60-85: anchor in float64, ground-truth extents formed in float32 first
            const double ax1 = anchors64[4 * i], ay1 = anchors64[4 * i + 1], ax2 = anchors64[4 * i + 2],
                         ay2 = anchors64[4 * i + 3];
            const float4 q = s_gt[s_arg[i]];
            const double bw = __dsub_rn(ax2, ax1), bh = __dsub_rn(ay2, ay1);
            const double bx = __dadd_rn(ax1, ax2) / 2.0, by = __dadd_rn(ay1, ay2) / 2.0;
            const float gw = __fsub_rn(q.z, q.x), gh = __fsub_rn(q.w, q.y);
            const float gx = __fmul_rn(__fadd_rn(q.x, q.z), 0.5f), gy = __fmul_rn(__fadd_rn(q.y, q.w), 0.5f);
            t[0] = (float)__ddiv_rn(__dsub_rn((double)gx, bx), bw);
            t[1] = (float)__ddiv_rn(__dsub_rn((double)gy, by), bh);
            t[2] = (float)log(__ddiv_rn((double)gw, bw));
            t[3] = (float)log(__ddiv_rn((double)gh, bh));
        }
        const float m = lab == 1 ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            loc_targets[(long long)(a * 4 + k) * plane + cell] = t[k];
            loc_masks[(long long)(a * 4 + k) * plane + cell] = m;
        }
    }
}

}  // namespace

SCDA_API int scda_anchor_targets(int A, int fh, int fw, const float *anchors32, const double *anchors64, int G,
                                 const float *gts, float neg_thresh, float pos_thresh, int want_pos,
                                 int batch_total, const double *keys_pos, const double *keys_neg,
                                 long long *cls_targets, float *loc_targets, float *loc_masks, long long *normalizer,
                                 cudaStream_t stream)
{
    if (A <= 0 || fh <= 0 || fw <= 0 || G <= 0 || G > kMaxG || want_pos < 0 || batch_total < want_pos) return 0;
    if (!anchors32 || !anchors64 || !gts || !keys_pos || !keys_neg || !cls_targets || !loc_targets || !loc_masks ||
        !normalizer || (uintptr_t)anchors32 % 16)
        return 0;
    const long long KA = (long long)A * fh * fw;
    if (KA > 100000) return 0;
    const size_t smem = 2 * (size_t)KA + 16;
    static size_t attr = 0;
    if (smem > 32 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(anchor_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return -(int)e;
        attr = smem;
    }
    AnchorParams p;
    p.KA = (int)KA; p.A = A; p.fh = fh; p.fw = fw; p.G = G; p.want_pos = want_pos; p.batch_total = batch_total;
    p.neg_thresh = neg_thresh; p.pos_thresh = pos_thresh; p.gt_floor = 0.1f;
    anchor_targets_kernel<<<1, kAT, smem, stream>>>(p, anchors32, anchors64, gts, keys_pos, keys_neg, cls_targets,
                                                   loc_targets, loc_masks, normalizer);
    return scda_launch_status();
}
