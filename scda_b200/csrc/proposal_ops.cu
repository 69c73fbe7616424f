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
// The whole front half of compute_rpn_proposals (functions/rpn_proposal.py:49-64) in two launches:
//   scores [KA] -> the `pre` best anchors in descending score order (the reference: numpy argpartition +
//   argsort on the host) -> decode / clip / min-size filter / compaction as above.
// Before: torch.topk (12 launches of a multi-block radix select) + a 5-launch radix sort + the single-CTA
// decode kernel above, ~0.5 ms of mostly launch gaps and dependent-load latency on the critical path of the
// detector forward.  Now:
//   1. rpn_select_kernel (one CTA): the scores become order-preserving 32-bit keys in shared memory; four
//      8-bit radix-select passes find the key of the pre-th best; the candidates (keys above it, and the
//      lowest-indexed ties on it) are written out in ANCHOR order — a stable compaction, so everything
//      downstream is deterministic.
//   2. rpn_rank_decode_kernel (pre / 32 CTAs): every CTA holds all candidate keys in shared memory; eight
//      lanes count, for one candidate, how many keys beat it (>= before it, > behind it: ties rank by anchor
//      index) — its position in the sorted order, no sorting network; one of the eight decodes the
//      anchor and writes the row at that position.  The last CTA to finish (a ticket armed by kernel 1)
//      drops the rows that failed the min-size test with one ordered scan and writes the count.
namespace {

constexpr int kSelThreads = 1024;
constexpr int kRankThreads = 256;
constexpr int kRankLanes = 8;                       // lanes per candidate
constexpr int kRankPerCta = kRankThreads / kRankLanes;

__device__ __forceinline__ unsigned sortable_key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// exclusive prefix of v over the CTA's threads (in thread order); *total = the sum.  s_w: >= 33 ints.
template <int kThreads>
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_w, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                                 // s_w may still be read from a previous call
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

__global__ void __launch_bounds__(kSelThreads)
rpn_select_kernel(int KA, int K, const float *__restrict__ scores, unsigned *__restrict__ cand_key,
                  int *__restrict__ cand_idx, unsigned *__restrict__ ticket)
{
    extern __shared__ unsigned s_keys[];              // [KA]
    __shared__ int s_hist[256];
    __shared__ int s_w[33];
    __shared__ unsigned s_prefix;
    __shared__ int s_need;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < KA; i += kSelThreads) s_keys[i] = sortable_key(__ldg(scores + i));
    if (tid == 0) { s_prefix = 0u; s_need = K; *ticket = 0u; }
    unsigned prefix = 0u, mask = 0u;
    int need = K;                                     // rank (from the top) of the wanted key among the matching ones
    if (K < KA) {
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int d = tid; d < 256; d += kSelThreads) s_hist[d] = 0;
            __syncthreads();
            for (int i = tid; i < KA; i += kSelThreads) {
                const unsigned k = s_keys[i];
                if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
            }
            __syncthreads();
            if (tid < 32) {
                // lane l: the eight digits 255 - 8l ... 248 - 8l, from the top
                int mine = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) mine += s_hist[255 - 8 * lane - q];
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
                const int first = __ffs(hit) - 1;     // need <= number of matching keys: always found
                if (lane == first) {
                    int above = incl - mine;          // matching keys with a larger digit than this lane's
                    int d = 255 - 8 * lane;
                    for (int q = 0; q < 8; ++q, --d) {
                        const int h = s_hist[d];
                        if (above + h >= need) break;
                        above += h;
                    }
                    s_prefix = prefix | ((unsigned)d << shift);
                    s_need = need - above;
                }
            }
            __syncthreads();
            prefix = s_prefix;
            need = s_need;
            mask |= 255u << shift;
        }
    }
    __syncthreads();
    // stable compaction: thread t owns anchors [t * per, (t + 1) * per)
    const unsigned T = prefix;                        // key of the K-th best (0: everything is taken)
    const int per = (KA + kSelThreads - 1) / kSelThreads;
    const int i0 = min(KA, tid * per), i1 = min(KA, i0 + per);
    int gt = 0, eq = 0;
    for (int i = i0; i < i1; ++i) {
        const unsigned k = s_keys[i];
        gt += k > T;
        eq += k == T;
    }
    int total;
    int eq_before = 0;
    if (K < KA) eq_before = block_exclusive_scan<kSelThreads>(eq, s_w, &total);
    const int eq_take = K < KA ? max(0, min(eq, need - eq_before)) : eq;
    int pos = block_exclusive_scan<kSelThreads>(gt + eq_take, s_w, &total);
    int eq_left = eq_take;
    for (int i = i0; i < i1; ++i) {
        const unsigned k = s_keys[i];
        bool take = k > T;
        if (k == T && eq_left > 0) { take = true; --eq_left; }
        if (take) {
            cand_key[pos] = k;
            cand_idx[pos] = i;
            ++pos;
        }
    }
}

__global__ void __launch_bounds__(kRankThreads)
rpn_rank_decode_kernel(int K, int Kpad, const unsigned *__restrict__ cand_key, const int *__restrict__ cand_idx,
                       const double *__restrict__ anchors, const float *__restrict__ deltas, double img_h,
                       double img_w, double min_size, float *__restrict__ rows, unsigned char *__restrict__ ok_flag,
                       unsigned *__restrict__ ticket, float *__restrict__ packed, int *__restrict__ count)
{
    extern __shared__ __align__(16) unsigned s_k[];   // [Kpad], zero beyond K
    __shared__ int s_w[33];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    {
        const int n4 = Kpad >> 2;                    // 16-byte loads, four in flight per thread
        uint4 *s4 = reinterpret_cast<uint4 *>(s_k);
        for (int g0 = tid; g0 < n4; g0 += 4 * kRankThreads) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = g0 + u * kRankThreads;
                if (g < n4) v[u] = __ldg(reinterpret_cast<const uint4 *>(cand_key) + g);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = g0 + u * kRankThreads;
                if (g >= n4) continue;
                uint4 w = v[u];
                if (4 * g + 1 >= K) w.y = 0u;
                if (4 * g + 2 >= K) w.z = 0u;
                if (4 * g + 3 >= K) w.w = 0u;
                s4[g] = w;
            }
        }
    }
    __syncthreads();
    const int i = blockIdx.x * kRankPerCta + tid / kRankLanes, part = tid % kRankLanes;
    const bool live = i < K;
    const unsigned ki = live ? s_k[i] : 0xffffffffu;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(s_k);
    const int gi = (live ? i : 0) >> 2, ng = Kpad >> 2;
    int cnt = 0;
    // candidates before i rank above it on ties, those behind it do not
    for (int g = part; g < gi; g += kRankLanes) {
        const uint4 v = k4[g];
        cnt += (v.x >= ki) + (v.y >= ki) + (v.z >= ki) + (v.w >= ki);
    }
    {
        const int first_after = gi + 1 + ((part - (gi + 1)) % kRankLanes + kRankLanes) % kRankLanes;
        for (int g = first_after; g < ng; g += kRankLanes) {
            const uint4 v = k4[g];
            cnt += (v.x > ki) + (v.y > ki) + (v.z > ki) + (v.w > ki);
        }
    }
    if (part == 0 && live) {
        const uint4 v = k4[gi];
        const unsigned e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = 4 * gi + q;
            cnt += j < i ? (e[q] >= ki) : (j > i ? (e[q] > ki) : 0);
        }
    }
#pragma unroll
    for (int o = 1; o < kRankLanes; o <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (part == 0 && live) {
        const int r = cnt;                           // position in descending order
        const long long a = cand_idx[i];
        const double ax1 = anchors[4 * a], ay1 = anchors[4 * a + 1], ax2 = anchors[4 * a + 2], ay2 = anchors[4 * a + 3];
        const double w = __dsub_rn(ax2, ax1), h = __dsub_rn(ay2, ay1);
        const double cx = __dadd_rn(ax1, ax2) / 2.0, cy = __dadd_rn(ay1, ay2) / 2.0;
        const float4 d = __ldg(reinterpret_cast<const float4 *>(deltas) + a);
        const double ncx = __dadd_rn(__dmul_rn((double)d.x, w), cx), ncy = __dadd_rn(__dmul_rn((double)d.y, h), cy);
        const double nw = __dmul_rn((double)expf(d.z), w), nh = __dmul_rn((double)expf(d.w), h);
        const double x1 = fmin(fmax(__dsub_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y1 = fmin(fmax(__dsub_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const double x2 = fmin(fmax(__dadd_rn(ncx, nw / 2.0), 0.0), img_w - 1.0);
        const double y2 = fmin(fmax(__dadd_rn(ncy, nh / 2.0), 0.0), img_h - 1.0);
        const bool ok = __dadd_rn(__dsub_rn(x2, x1), 1.0) >= min_size && __dadd_rn(__dsub_rn(y2, y1), 1.0) >= min_size;
        float *rr = rows + 5ll * r;
        rr[0] = (float)x1; rr[1] = (float)y1; rr[2] = (float)x2; rr[3] = (float)y2; rr[4] = key_to_float(ki);
        ok_flag[r] = ok ? 1 : 0;
    }
    // the last CTA compacts
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // thread t owns the 16 * per16 consecutive ranks behind 16 * per16 * t: their flags are per16 aligned 16-byte
    // loads, all in flight at once; a block scan gives every survivor its destination; the row copy then runs
    // over DESTINATIONS through a source list in shared memory (the key array is dead by now), so that the
    // gathers of consecutive threads are independent and the stores coalesced
    int *s_src = reinterpret_cast<int *>(s_k);        // [<= K] rank of the row that lands at position d
    const int per16 = (K + 16 * kRankThreads - 1) / (16 * kRankThreads);
    const int r0 = tid * per16 * 16;
    int mine = 0;
    for (int v0 = 0; v0 < per16; v0 += 4) {
        uint4 f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + (v0 + u) * 16;
            f[u] = (v0 + u < per16 && r < K) ? __ldcg(reinterpret_cast<const uint4 *>(ok_flag + r)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + (v0 + u) * 16;
            const unsigned w[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int bte = 0; bte < 4; ++bte)
                    if (r + 4 * q + bte < K && ((w[q] >> (8 * bte)) & 1u)) ++mine;
        }
    }
    int total;
    int pos = block_exclusive_scan<kRankThreads>(mine, s_w, &total);
    for (int v = 0; v < per16; ++v) {
        const int r = r0 + v * 16;
        if (r >= K) break;
        const uint4 f = __ldcg(reinterpret_cast<const uint4 *>(ok_flag + r));
        const unsigned w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int bte = 0; bte < 4; ++bte)
                if (r + 4 * q + bte < K && ((w[q] >> (8 * bte)) & 1u)) s_src[pos++] = r + 4 * q + bte;
    }
    __syncthreads();
    const int n5 = total * 5;
    for (int k0 = tid; k0 < n5; k0 += 8 * kRankThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kRankThreads;
            if (k < n5) {
                const int d = k / 5, c = k - 5 * d;
                v[u] = __ldcg(rows + 5ll * s_src[d] + c);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kRankThreads;
            if (k < n5) packed[k] = v[u];
        }
    }
    for (long long k = (long long)total * 5 + tid; k < (long long)K * 5; k += kRankThreads) packed[k] = 0.f;
    if (tid == 0) { count[0] = total; *ticket = 0u; }
}

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

SCDA_API size_t scda_rpn_proposal_rows_workspace_bytes(int KA, int pre)
{
    const size_t K = (size_t)((pre <= 0 || pre > KA) ? KA : pre);
    return align16(K * 4) + align16(K * 4) + align16(K * 20) + align16(K) + 16 + 16;
}

SCDA_API int scda_rpn_proposal_rows(int KA, int pre, const float *scores, const double *anchors, const float *deltas,
                                    double img_h, double img_w, double min_size, float *packed, int *count,
                                    void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    if (KA <= 0 || !scores || !anchors || !deltas || !packed || !count || !workspace) return 0;
    if ((uintptr_t)deltas % 16 || (uintptr_t)workspace % 16) return 0;
    const int K = (pre <= 0 || pre > KA) ? KA : pre;
    if (workspace_bytes < scda_rpn_proposal_rows_workspace_bytes(KA, pre)) return 0;
    const int Kpad = (K + 3) & ~3;
    const size_t smem1 = sizeof(unsigned) * (size_t)KA, smem2 = sizeof(unsigned) * (size_t)Kpad;
    if (smem1 > 200 * 1024) return 0;               // KA <= 51 200 anchors per image
    static size_t attr1 = 0, attr2 = 0;
    if (smem1 > 48 * 1024 && smem1 > attr1) {
        cudaError_t e = cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return -(int)e;
        attr1 = smem1;
    }
    if (smem2 > 48 * 1024 && smem2 > attr2) {
        cudaError_t e = cudaFuncSetAttribute(rpn_rank_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem2);
        if (e != cudaSuccess) return -(int)e;
        attr2 = smem2;
    }
    unsigned char *ws = (unsigned char *)workspace;
    unsigned *cand_key = (unsigned *)ws;
    ws += align16((size_t)K * 4);
    int *cand_idx = (int *)ws;
    ws += align16((size_t)K * 4);
    float *rows = (float *)ws;
    ws += align16((size_t)K * 20);
    unsigned char *ok_flag = ws;
    ws += align16((size_t)K);
    unsigned *ticket = (unsigned *)ws;
    rpn_select_kernel<<<1, kSelThreads, smem1, stream>>>(KA, K, scores, cand_key, cand_idx, ticket);
    int st = scda_launch_status();
    if (st != 1) return st;
    rpn_rank_decode_kernel<<<ceil_div(K, kRankPerCta), kRankThreads, smem2, stream>>>(
        K, Kpad, cand_key, cand_idx, anchors, deltas, img_h, img_w, min_size, rows, ok_flag, ticket, packed, count);
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
