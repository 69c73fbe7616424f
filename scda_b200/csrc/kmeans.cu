// Region grouping on the device: k-means of the RoI centres + selection of `threshold`
// member rows per cluster, in one single-CTA kernel.
//
// Replaces the host round trip of compute_cluster_targets (functions/mask.py:193-237 of
// the reference): D2H of the RoIs and of the 512 x 4096 fc7 block, scikit-learn
// `KMeans(n_clusters, random_state=0).fit(centres)` on float32 centres, np.where /
// np.random.choice per cluster, H2D of the gathered block — twice per iteration, each a
// full device synchronisation.  Here the labels, the cluster centres and the row indices to
// gather are produced on the device; the feature rows never leave HBM and nothing
// synchronises, so the whole iteration can be captured in a CUDA graph.
//
// Algorithm = scikit-learn's (the reference does not pin a version; restated from 1.9,
// sklearn/cluster/_kmeans.py:_kmeans_plusplus / _kmeans_single_lloyd and
// _k_means_lloyd.pyx:lloyd_iter_chunked_dense), float32 data as the reference passes it:
//   X = ((x1+x2)/2, (y1+y2)/2) in float32;  X -= mean (sequential float32 sums)
//   k-means++ seeding, n_init = 1, 2 + int(ln k) local trials: the first centre index and
//     the uniform draws of RandomState(0) are data independent, the host passes them in;
//     distances as sklearn's float32 path computes them (float64 arithmetic on the float32
//     data + float32 squared norms, rounded to float32, clipped at 0); candidates by
//     searchsorted on the sequential float32 cumulative sum
//   Lloyd: argmin_j (|c_j|^2 - 2 x.c_j) in float32 (first minimum wins), centre sums per
//     256-sample chunk then chunk order, centre *= 1/count, stop on unchanged labels or
//     sum(shift^2) <= 1e-4 * mean(var(X)); final E-step if not strictly converged
//   centres += mean.
// BLAS-internal summation orders (sdot / sgemm) are not specified, so agreement with a
// given scikit-learn build is exact except at float32 rounding knife edges (a candidate or
// a label flipping when two float32 numbers differ in the last bit); tests/ pins the
// agreement rate against the installed scikit-learn.
#include "common.cuh"

namespace {

constexpr int kMaxN = 2048;
constexpr int kMaxK = 16;
constexpr int kMaxTrials = 8;
constexpr int kThreadsKM = 256;
constexpr int kChunk = 256;      // sklearn CHUNK_SIZE

struct KmSmem {
    float x[kMaxN][2];
    float ysq[kMaxN];            // float32 squared norms of the centred points
    float closest[kMaxN];
    float cand[kMaxTrials][kMaxN];
    float cums[kMaxN];
    int label[kMaxN];
    int label_old[kMaxN];
    float centers[kMaxK][2];
    float centers_new[kMaxK][2];
    float csq[kMaxK];
    float weight[kMaxK];
    double red[kThreadsKM];
    float mean[2];
    float tol;
    float pot;
    float cand_pot[kMaxTrials];
    int cand_id[kMaxTrials];
    int flag;
    int members[kMaxK];
};

__device__ double block_sum(double v, double *red)
{
    const int t = threadIdx.x;
    red[t] = v;
    __syncthreads();
    for (int s = kThreadsKM / 2; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// The sequential float32 sums below (numpy's row-by-row reductions, sklearn's chunked centre sums) are dependent
// chains by definition; what made them slow was one shared-memory load latency per element in front of every
// add.  kSeqBatch values are loaded first (independent loads, pipelined), then added in order: same sums, same
// order, ~8x shorter.
constexpr int kSeqBatch = 16;

// sklearn.metrics.pairwise._euclidean_distances_upcast for one (centre, point) pair
__device__ __forceinline__ float upcast_dist(const float cx, const float cy, const float px, const float py,
                                             const float ysq)
{
    const double dot = (double)cx * (double)px + (double)cy * (double)py;
    const double xx = (double)cx * (double)cx + (double)cy * (double)cy;
    const float d = (float)(-2.0 * dot + xx + (double)ysq);
    return d > 0.f ? d : 0.f;
}

__global__ void __launch_bounds__(kThreadsKM, 1)
kmeans_kernel(const float *__restrict__ rois, int roi_stride, int n, int K, int first_id,
              const double *__restrict__ uni, int trials, int max_iter, float tol_rel,
              const float *__restrict__ pick_uniform, int T, int *__restrict__ labels_out,
              float *__restrict__ centers_out, int *__restrict__ counts_out, long long *__restrict__ index_out,
              int *__restrict__ member_ws)
{
    extern __shared__ unsigned char km_raw[];
    KmSmem &S = *reinterpret_cast<KmSmem *>(km_raw);
    const int t = threadIdx.x;

    // centres of the RoIs, float32 as numpy computes (x2 + x1) / 2.0 on a float32 array
    for (int i = t; i < n; i += kThreadsKM) {
        const float *p = rois + (long long)i * roi_stride;
        S.x[i][0] = __fdiv_rn(__fadd_rn(p[3], p[1]), 2.0f);
        S.x[i][1] = __fdiv_rn(__fadd_rn(p[4], p[2]), 2.0f);
    }
    __syncthreads();
    if (t < 2) {                       // X.mean(axis=0): sequential float32 sum down the rows
        float acc = 0.f;
        for (int i0 = 0; i0 < n; i0 += kSeqBatch) {
            float v[kSeqBatch];
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u) v[u] = i0 + u < n ? S.x[i0 + u][t] : 0.f;
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u)
                if (i0 + u < n) acc = __fadd_rn(acc, v[u]);
        }
        S.mean[t] = __fdiv_rn(acc, (float)n);
    }
    __syncthreads();
    for (int i = t; i < n; i += kThreadsKM) {
        S.x[i][0] = __fsub_rn(S.x[i][0], S.mean[0]);
        S.x[i][1] = __fsub_rn(S.x[i][1], S.mean[1]);
        S.ysq[i] = __fadd_rn(__fmul_rn(S.x[i][0], S.x[i][0]), __fmul_rn(S.x[i][1], S.x[i][1]));
        S.label[i] = -1;
        S.label_old[i] = -1;
    }
    __syncthreads();
    if (t < 2) {                       // np.var(X, axis=0) in float32, then tol = mean(var) * tol_rel
        float m = 0.f;
        for (int i0 = 0; i0 < n; i0 += kSeqBatch) {
            float xv[kSeqBatch];
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u) xv[u] = i0 + u < n ? S.x[i0 + u][t] : 0.f;
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u)
                if (i0 + u < n) m = __fadd_rn(m, xv[u]);
        }
        m = __fdiv_rn(m, (float)n);
        float v = 0.f;
        for (int i0 = 0; i0 < n; i0 += kSeqBatch) {
            float dd[kSeqBatch];
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u) {
                const float d = i0 + u < n ? __fsub_rn(S.x[i0 + u][t], m) : 0.f;
                dd[u] = __fmul_rn(d, d);
            }
#pragma unroll
            for (int u = 0; u < kSeqBatch; ++u)
                if (i0 + u < n) v = __fadd_rn(v, dd[u]);
        }
        S.cums[t] = __fdiv_rn(v, (float)n);
    }
    __syncthreads();
    if (t == 0) S.tol = __fmul_rn(__fdiv_rn(__fadd_rn(S.cums[0], S.cums[1]), 2.0f), tol_rel);

    // ---- k-means++ seeding
    if (t == 0) {
        S.centers[0][0] = S.x[first_id][0];
        S.centers[0][1] = S.x[first_id][1];
    }
    __syncthreads();
    double part = 0.0;
    for (int i = t; i < n; i += kThreadsKM) {
        const float d = upcast_dist(S.centers[0][0], S.centers[0][1], S.x[i][0], S.x[i][1], S.ysq[i]);
        S.closest[i] = d;
        part += (double)d;
    }
    double pot = block_sum(part, S.red);
    if (t == 0) S.pot = (float)pot;
    __syncthreads();
    for (int c = 1; c < K; ++c) {
        if (t == 0) {                  // np.cumsum in float32 is sequential
            float acc = 0.f;
            for (int i0 = 0; i0 < n; i0 += kSeqBatch) {
                float v[kSeqBatch];
#pragma unroll
                for (int u = 0; u < kSeqBatch; ++u) v[u] = i0 + u < n ? S.closest[i0 + u] : 0.f;
#pragma unroll
                for (int u = 0; u < kSeqBatch; ++u)
                    if (i0 + u < n) {
                        acc = __fadd_rn(acc, v[u]);
                        S.cums[i0 + u] = acc;
                    }
            }
        }
        __syncthreads();
        if (t < trials) {              // searchsorted(cums, rand, side='left'), clipped
            const double r = uni[(c - 1) * trials + t] * (double)S.pot;
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((double)S.cums[mid] < r) lo = mid + 1; else hi = mid;
            }
            S.cand_id[t] = lo < n - 1 ? lo : n - 1;
        }
        __syncthreads();
        for (int q = 0; q < trials; ++q) {
            const int id = S.cand_id[q];
            const float cx = S.x[id][0], cy = S.x[id][1];
            double p = 0.0;
            for (int i = t; i < n; i += kThreadsKM) {
                float d = upcast_dist(cx, cy, S.x[i][0], S.x[i][1], S.ysq[i]);
                d = fminf(d, S.closest[i]);
                S.cand[q][i] = d;
                p += (double)d;
            }
            const double tot = block_sum(p, S.red);
            if (t == 0) S.cand_pot[q] = (float)tot;
        }
        __syncthreads();
        if (t == 0) {
            int best = 0;
            for (int q = 1; q < trials; ++q)
                if (S.cand_pot[q] < S.cand_pot[best]) best = q;
            S.flag = best;
            S.pot = S.cand_pot[best];
            S.centers[c][0] = S.x[S.cand_id[best]][0];
            S.centers[c][1] = S.x[S.cand_id[best]][1];
        }
        __syncthreads();
        const int best = S.flag;
        for (int i = t; i < n; i += kThreadsKM) S.closest[i] = S.cand[best][i];
        __syncthreads();
    }

    // ---- Lloyd iterations
    const int n_chunks = (n + kChunk - 1) / kChunk;
    bool strict = false;
    for (int it = 0; it < max_iter; ++it) {
        if (t < K)
            S.csq[t] = __fadd_rn(__fmul_rn(S.centers[t][0], S.centers[t][0]),
                                 __fmul_rn(S.centers[t][1], S.centers[t][1]));
        __syncthreads();
        int changed = 0;
        for (int i = t; i < n; i += kThreadsKM) {
            int lab = 0;
            float best = 0.f;
            for (int j = 0; j < K; ++j) {
                const float dot = __fadd_rn(__fmul_rn(S.x[i][0], S.centers[j][0]),
                                            __fmul_rn(S.x[i][1], S.centers[j][1]));
                const float d = __fadd_rn(S.csq[j], __fmul_rn(-2.0f, dot));
                if (j == 0 || d < best) { best = d; lab = j; }
            }
            S.label[i] = lab;
            changed |= (lab != S.label_old[i]);
        }
        if (t == 0) S.flag = 0;
        __syncthreads();
        if (changed) S.flag = 1;
        // M step: thread (j, k) sums its cluster's coordinate chunk by chunk, in sample order
        if (t < K * 2) {
            const int j = t >> 1, k = t & 1;
            float tot = 0.f, wtot = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch) {
                float acc = 0.f, w = 0.f;
                const int e = min(n, (ch + 1) * kChunk);
                for (int i0 = ch * kChunk; i0 < e; i0 += kSeqBatch) {
                    float xv[kSeqBatch];
                    bool in[kSeqBatch];
#pragma unroll
                    for (int u = 0; u < kSeqBatch; ++u) {
                        const bool ok = i0 + u < e;
                        in[u] = ok && S.label[i0 + u] == j;
                        xv[u] = ok ? S.x[i0 + u][k] : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < kSeqBatch; ++u)
                        if (in[u]) { acc = __fadd_rn(acc, xv[u]); w += 1.f; }
                }
                tot = __fadd_rn(tot, acc);
                wtot += w;
            }
            S.centers_new[j][k] = tot;
            if (k == 0) S.weight[j] = wtot;
        }
        __syncthreads();
        if (t == 0) {
            // empty clusters: move the farthest points (from their old centres) into them
            int n_empty = 0;
            for (int j = 0; j < K; ++j) n_empty += S.weight[j] == 0.f;
            if (n_empty) {
                for (int i = 0; i < n; ++i) {
                    const int l = S.label[i];
                    const float dx = __fsub_rn(S.x[i][0], S.centers[l][0]), dy = __fsub_rn(S.x[i][1], S.centers[l][1]);
                    S.cums[i] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                }
                for (int j = 0; j < K; ++j) {
                    if (S.weight[j] != 0.f) continue;
                    int far = 0;
                    for (int i = 1; i < n; ++i)
                        if (S.cums[i] > S.cums[far]) far = i;
                    if (S.cums[far] == 0.f) break;
                    const int old = S.label[far];
                    for (int k = 0; k < 2; ++k) {
                        S.centers_new[old][k] = __fsub_rn(S.centers_new[old][k], S.x[far][k]);
                        S.centers_new[j][k] = S.x[far][k];
                    }
                    S.weight[j] = 1.f;
                    S.weight[old] -= 1.f;
                    S.cums[far] = -1.f;
                }
            }
            int arg = 0;
            for (int j = 1; j < K; ++j)
                if (S.weight[j] > S.weight[arg]) arg = j;
            float shift_tot = 0.f;
            for (int j = 0; j < K; ++j) {
                if (S.weight[j] > 0.f) {
                    const float alpha = __fdiv_rn(1.0f, S.weight[j]);
                    S.centers_new[j][0] = __fmul_rn(S.centers_new[j][0], alpha);
                    S.centers_new[j][1] = __fmul_rn(S.centers_new[j][1], alpha);
                } else {
                    S.centers_new[j][0] = S.centers_new[arg][0];
                    S.centers_new[j][1] = S.centers_new[arg][1];
                }
            }
            for (int j = 0; j < K; ++j) {
                const float dx = __fsub_rn(S.centers_new[j][0], S.centers[j][0]);
                const float dy = __fsub_rn(S.centers_new[j][1], S.centers[j][1]);
                const float sh = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                shift_tot = __fadd_rn(shift_tot, __fmul_rn(sh, sh));
                S.centers[j][0] = S.centers_new[j][0];
                S.centers[j][1] = S.centers_new[j][1];
            }
            // 0 = continue, 1 = strict convergence, 2 = tolerance convergence
            S.members[0] = S.flag == 0 ? 1 : (shift_tot <= S.tol ? 2 : 0);
        }
        __syncthreads();
        const int stop = S.members[0];
        for (int i = t; i < n; i += kThreadsKM) S.label_old[i] = S.label[i];
        __syncthreads();
        if (stop == 1) { strict = true; break; }
        if (stop == 2) break;
    }
    if (!strict) {                     // labels consistent with the final centres
        if (t < K)
            S.csq[t] = __fadd_rn(__fmul_rn(S.centers[t][0], S.centers[t][0]),
                                 __fmul_rn(S.centers[t][1], S.centers[t][1]));
        __syncthreads();
        for (int i = t; i < n; i += kThreadsKM) {
            int lab = 0;
            float best = 0.f;
            for (int j = 0; j < K; ++j) {
                const float dot = __fadd_rn(__fmul_rn(S.x[i][0], S.centers[j][0]),
                                            __fmul_rn(S.x[i][1], S.centers[j][1]));
                const float d = __fadd_rn(S.csq[j], __fmul_rn(-2.0f, dot));
                if (j == 0 || d < best) { best = d; lab = j; }
            }
            S.label[i] = lab;
        }
        __syncthreads();
    }

    // ---- outputs
    for (int i = t; i < n; i += kThreadsKM) labels_out[i] = S.label[i];
    if (t < K) {
        centers_out[t * 2 + 0] = __fadd_rn(S.centers[t][0], S.mean[0]);
        centers_out[t * 2 + 1] = __fadd_rn(S.centers[t][1], S.mean[1]);
        // members of cluster t in ascending index order (np.where(labels == t)[0])
        int cnt = 0;
        int *mem = member_ws + (long long)t * n;
        for (int i = 0; i < n; ++i)
            if (S.label[i] == t) mem[cnt++] = i;
        S.members[t] = cnt;
        counts_out[t] = cnt;
    }
    __syncthreads();
    if (index_out && T > 0) {
        // keep_ix[:T] when the cluster is large enough, else T draws with replacement
        // (functions/mask.py:216-222; the reference draws with the unseeded numpy RNG)
        for (int q = t; q < K * T; q += kThreadsKM) {
            const int c = q / T, r = q - c * T;
            const int cnt = S.members[c];
            const int *mem = member_ws + (long long)c * n;
            int pick;
            if (cnt >= T) pick = mem[r];
            else if (cnt > 0) {
                int u = (int)(pick_uniform[q] * (float)cnt);
                pick = mem[u < cnt ? u : cnt - 1];
            } else pick = 0;
            index_out[q] = pick;
        }
    }
}

}  // namespace

SCDA_API size_t scda_kmeans_workspace_bytes(int n, int k) { return sizeof(int) * (size_t)n * (size_t)k; }

SCDA_API int scda_kmeans_regions(const float *rois, int roi_stride, int n, int k, int first_center_id,
                                 const double *uniforms, int n_local_trials, int max_iter, float tol,
                                 const float *pick_uniform, int threshold, int *labels, float *centers,
                                 int *counts, long long *index, void *workspace, size_t workspace_bytes,
                                 cudaStream_t stream)
{
    if (!rois || !uniforms || !labels || !centers || !counts || !workspace) return 0;
    if (n < 1 || n > kMaxN || k < 1 || k > kMaxK || k > n || roi_stride < 5) return 0;
    if (n_local_trials < 1 || n_local_trials > kMaxTrials || first_center_id < 0 || first_center_id >= n) return 0;
    if (threshold > 0 && (!index || !pick_uniform)) return 0;
    if (workspace_bytes < scda_kmeans_workspace_bytes(n, k)) return 0;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(KmSmem));
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    kmeans_kernel<<<1, kThreadsKM, sizeof(KmSmem), stream>>>(rois, roi_stride, n, k, first_center_id, uniforms,
                                                              n_local_trials, max_iter, tol, pick_uniform,
                                                              threshold, labels, centers, counts, index,
                                                              (int *)workspace);
    return scda_launch_status();
}
