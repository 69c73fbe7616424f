// Fused loss kernels of the detector and the adversarial phases (HBM / latency bound, tiny):
//
//  * smooth-L1 with sigma on masked predictions, summed —
//    `smooth_l1_loss_with_sigma(pred * mask, target)` of the reference
//    (models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:238-246, called at :54-55 and
//    :64-66): d = pred*mask - target; loss = sum(0.5 s^2 d^2 if |d| < 1/s^2 else |d| - 0.5/s^2).
//    The reference spends ~10 elementwise launches and a reduction on it, forward alone.
//  * sigmoid + per-row binary cross entropy — the adversarial terms of the four-phase update
//    (tools/faster_rcnn_train_val.py:577-600, 655-680, 716-732): `F.binary_cross_entropy(sigmoid(x[k]), y)`
//    for every cluster row k against ONE label row y (soft labels U(0.8,1)/U(0,0.3) or a constant),
//    with torch's clamp of the logs at -100 (aten/native/Loss.cpp) and its backward
//    (p - y) / max(p (1 - p), 1e-12) chained through the sigmoid.
//
// Reductions are block-local trees in a fixed order (one block per output): deterministic.
#include "common.cuh"

namespace {

constexpr int kLT = 1024;

__device__ __forceinline__ float block_sum(float v, float *sh)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    const float r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kLT)
smooth_l1_fwd_kernel(const float *__restrict__ pred, const float *__restrict__ mask,
                     const float *__restrict__ target, long long n, float sigma2, float *__restrict__ out)
{
    __shared__ float sh[32];
    const float thr = 1.f / sigma2, half = 0.5f / sigma2;
    float acc = 0.f;
    auto term = [&](float p, float m, float t) {
        const float d = p * m - t;
        const float a = fabsf(d);
        return a < thr ? d * d * sigma2 * 0.5f : a - half;
    };
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
          reinterpret_cast<uintptr_t>(mask)) & 15) == 0) {
        // 16-byte loads, two in flight per operand: the single block is bound by load latency, not bandwidth
        const long long n4 = n / 4;
        const float4 *p4 = reinterpret_cast<const float4 *>(pred), *t4 = reinterpret_cast<const float4 *>(target);
        const float4 *m4 = reinterpret_cast<const float4 *>(mask);
        const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
        long long i = threadIdx.x;
        for (; i + kLT < n4; i += 2 * kLT) {
            const float4 pa = p4[i], pb = p4[i + kLT], ta = t4[i], tb = t4[i + kLT];
            const float4 ma = mask ? m4[i] : one, mb = mask ? m4[i + kLT] : one;
            acc += term(pa.x, ma.x, ta.x) + term(pa.y, ma.y, ta.y) + term(pa.z, ma.z, ta.z) + term(pa.w, ma.w, ta.w);
            acc += term(pb.x, mb.x, tb.x) + term(pb.y, mb.y, tb.y) + term(pb.z, mb.z, tb.z) + term(pb.w, mb.w, tb.w);
        }
        for (; i < n4; i += kLT) {
            const float4 pa = p4[i], ta = t4[i];
            const float4 ma = mask ? m4[i] : one;
            acc += term(pa.x, ma.x, ta.x) + term(pa.y, ma.y, ta.y) + term(pa.z, ma.z, ta.z) + term(pa.w, ma.w, ta.w);
        }
        done = n4 * 4;
    }
    for (long long i = done + threadIdx.x; i < n; i += kLT)
        acc += term(pred[i], mask ? mask[i] : 1.f, target[i]);
    const float s = block_sum(acc, sh);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256)
smooth_l1_bwd_kernel(const float *__restrict__ pred, const float *__restrict__ mask,
                     const float *__restrict__ target, const float *__restrict__ gout, long long n, float sigma2,
                     float *__restrict__ gpred)
{
    const float thr = 1.f / sigma2;
    const float g = gout[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f;
        const float d = pred[i] * m - target[i];
        const float a = fabsf(d);
        const float dl = a < thr ? sigma2 * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        gpred[i] = g * dl * m;
    }
}

// one block per row k: out[k] = mean_m BCE(sigmoid(x[k, m]), y[m * ystride])
__global__ void __launch_bounds__(kLT)
bce_rows_fwd_kernel(const float *__restrict__ x, const float *__restrict__ y, int ystride, int M,
                    float *__restrict__ out)
{
    __shared__ float sh[32];
    const int k = blockIdx.x;
    float acc = 0.f;
    for (int m = threadIdx.x; m < M; m += kLT) {
        const float p = 1.f / (1.f + expf(-x[(long long)k * M + m]));
        const float t = y[(long long)m * ystride];
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);     // Loss.cu of torch 2.x
        acc += (t - 1.f) * lq - t * lp;
    }
    const float s = block_sum(acc, sh);
    if (threadIdx.x == 0) out[k] = s / (float)M;
}

__global__ void __launch_bounds__(256)
bce_rows_bwd_kernel(const float *__restrict__ x, const float *__restrict__ y, int ystride, int K, int M,
                    const float *__restrict__ gout, float *__restrict__ gx)
{
    const long long n = (long long)K * M;
    const float inv = 1.f / (float)M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / M), m = (int)(i - (long long)k * M);
        const float p = 1.f / (1.f + expf(-x[i]));
        const float t = y[(long long)m * ystride];
        const float q = (1.f - p) * p;
        // binary_cross_entropy_backward then sigmoid_backward, as torch chains them
        const float gp = gout[k] * inv * (p - t) / fmaxf(q, 1e-12f);
        gx[i] = gp * q;
    }
}

// ------------------------------------------------------------------ softmax cross-entropy + top-1 accuracy
// F.cross_entropy(logits, targets, ignore_index) (mean over the rows whose target != ignore_index) and
// `accuracy(output, target, topk=(1,))` of the reference in ONE pass over the logits:
// _add_rpn_loss / _add_rcnn_loss, models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:36-68,
// 249-267 (there: log_softmax + nll_loss + topk + eq + nonzero/index + sum, ~15 launches and a host
// synchronisation in `accuracy`).  One thread per row (C <= 32 classes: 2 for the RPN, 9 for the RCNN head);
// per-block partials (sum of -log p_target, rows counted, rows whose argmax is the target) in a fixed
// order, a one-block finish: deterministic.
constexpr int kCeThreads = 256;

__global__ void __launch_bounds__(kCeThreads)
softmax_ce_partial_kernel(const float *__restrict__ x, long long ld, const long long *__restrict__ target,
                          long long M, int C, long long ignore_index, float *__restrict__ partial)
{
    __shared__ float sh[32];
    float nll = 0.f, cnt = 0.f, hit = 0.f;
    for (long long m = (long long)blockIdx.x * kCeThreads + threadIdx.x; m < M;
         m += (long long)gridDim.x * kCeThreads) {
        const long long t = target[m];
        if (t == ignore_index) continue;
        const float *row = x + m * ld;
        float mx = row[0];
        int am = 0;
        for (int c = 1; c < C; ++c) {
            const float v = row[c];
            if (v > mx) { mx = v; am = c; }
        }
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        const float xt = (t >= 0 && t < C) ? row[t] : 0.f;
        nll += (logf(sum) + mx) - xt;
        cnt += 1.f;
        hit += (am == (int)t) ? 1.f : 0.f;
    }
    nll = block_sum(nll, sh);
    cnt = block_sum(cnt, sh);
    hit = block_sum(hit, sh);
    if (threadIdx.x == 0) {
        partial[3 * blockIdx.x] = nll;
        partial[3 * blockIdx.x + 1] = cnt;
        partial[3 * blockIdx.x + 2] = hit;
    }
}

// out[0] = mean loss (NaN if no row counted, as torch), out[1] = top-1 accuracy in percent, out[2] = rows counted
__global__ void softmax_ce_finish_kernel(const float *__restrict__ partial, int blocks, float *__restrict__ out)
{
    if (threadIdx.x != 0) return;
    float nll = 0.f, cnt = 0.f, hit = 0.f;
    for (int b = 0; b < blocks; ++b) {
        nll += partial[3 * b];
        cnt += partial[3 * b + 1];
        hit += partial[3 * b + 2];
    }
    out[0] = nll / cnt;
    out[1] = hit * (100.f / fmaxf(cnt, 1.f));
    out[2] = cnt;
}

// dx[m, c] = gout * (softmax(x[m])[c] - [c == target[m]]) / rows counted; 0 for ignored rows
__global__ void __launch_bounds__(kCeThreads)
softmax_ce_bwd_kernel(const float *__restrict__ x, long long ld, const long long *__restrict__ target, long long M,
                      int C, long long ignore_index, const float *__restrict__ stats,
                      const float *__restrict__ gout, float *__restrict__ dx, long long lddx)
{
    const float scale = gout[0] / stats[2];
    for (long long m = (long long)blockIdx.x * kCeThreads + threadIdx.x; m < M;
         m += (long long)gridDim.x * kCeThreads) {
        const long long t = target[m];
        float *drow = dx + m * lddx;
        if (t == ignore_index) {
            for (int c = 0; c < C; ++c) drow[c] = 0.f;
            continue;
        }
        const float *row = x + m * ld;
        float mx = row[0];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
        const float inv = 1.f / sum;
        for (int c = 0; c < C; ++c) drow[c] = scale * (expf(row[c] - mx) * inv - (c == (int)t ? 1.f : 0.f));
    }
}

// RPN objectness: 2-way softmax over each anchor's (bg, fg) channel pair of the NCHW class map,
// foreground probability written straight in the proposal stage's anchor order
// (…reweight_cluster.py:153-155 + functions/rpn_proposal.py:44-49: permute to NHWC, view(-1, 2), softmax,
// permute back, permute again, reshape, take column 1: five layout passes and a softmax there).
// cls [B, 2A, H, W] -> score [B, H*W*A], score[b, (h*W + w)*A + a] = softmax(cls[b, 2a : 2a+2, h, w])[1]
__global__ void __launch_bounds__(256)
rpn_fg_score_kernel(const float *__restrict__ cls, int B, int A, long long HW, float *__restrict__ score)
{
    const long long total = (long long)B * HW * A;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i % A);
        const long long hw = (i / A) % HW;
        const long long b = i / (A * HW);
        const float *p = cls + (b * 2 * A + 2 * a) * HW + hw;
        const float l0 = p[0], l1 = p[HW];
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        score[i] = e1 / (e0 + e1);
    }
}

int ew_grid(long long n)
{
    long long b = (n + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

SCDA_API int scda_smooth_l1_sigma_sum_fwd(long long n, const float *pred, const float *mask, const float *target,
                                          float sigma, float *loss_sum, cudaStream_t stream)
{
    if (n <= 0 || !pred || !target || !loss_sum || !(sigma > 0.f)) return 0;
    smooth_l1_fwd_kernel<<<1, kLT, 0, stream>>>(pred, mask, target, n, sigma * sigma, loss_sum);
    return scda_launch_status();
}

SCDA_API int scda_smooth_l1_sigma_sum_bwd(long long n, const float *pred, const float *mask, const float *target,
                                          float sigma, const float *grad_loss, float *grad_pred, cudaStream_t stream)
{
    if (n <= 0 || !pred || !target || !grad_loss || !grad_pred || !(sigma > 0.f)) return 0;
    smooth_l1_bwd_kernel<<<ew_grid(n), 256, 0, stream>>>(pred, mask, target, grad_loss, n, sigma * sigma, grad_pred);
    return scda_launch_status();
}

SCDA_API int scda_bce_sigmoid_rows_fwd(int K, int M, const float *logits, const float *labels, int label_stride,
                                       float *row_mean, cudaStream_t stream)
{
    if (K <= 0 || M <= 0 || !logits || !labels || !row_mean || label_stride < 0) return 0;
    bce_rows_fwd_kernel<<<K, kLT, 0, stream>>>(logits, labels, label_stride, M, row_mean);
    return scda_launch_status();
}

SCDA_API int scda_bce_sigmoid_rows_bwd(int K, int M, const float *logits, const float *labels, int label_stride,
                                       const float *grad_rows, float *grad_logits, cudaStream_t stream)
{
    if (K <= 0 || M <= 0 || !logits || !labels || !grad_rows || !grad_logits || label_stride < 0) return 0;
    bce_rows_bwd_kernel<<<ew_grid((long long)K * M), 256, 0, stream>>>(logits, labels, label_stride, K, M, grad_rows,
                                                                       grad_logits);
    return scda_launch_status();
}

// ---------------------------------------------------------------------------------------------------
// Decoder head: nn.ConvTranspose2d(Cin, Cout, kernel_size=1) + nn.Tanh (the last two layers of each decoder,
// models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:380-383 of the reference: 32 -> 3
// channels over 4 x 256 x 256 pixels).  cuDNN / cuBLAS run this 0.05-GFLOP layer as GEMMs with a 3-wide
// dimension (100 us for the data gradient alone); it is one streaming pass each way:
//   forward : y[p, co] = tanh(b[co] + sum_ci x[p, ci] W[ci, co])                      (x, y channels-last)
//   backward: g = dy * (1 - y^2);  dx[p, ci] = sum_co W[ci, co] g[co];
//             dW[ci, co] = sum_p x[p, ci] g[co],  db[co] = sum_p g[co]  — per-thread register sums over a
//             grid-stride pixel loop, a fixed-order block tree, one partial row per block, then a second
//             kernel adds the block rows in order (deterministic).
namespace {

constexpr int kHeadCin = 32, kHeadCoutMax = 4, kHeadThreads = 256;

template <int kCout>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                float *__restrict__ y, long long P)
{
    __shared__ float sw[kHeadCin * kCout + kCout];
    for (int i = threadIdx.x; i < kHeadCin * kCout + kCout; i += kHeadThreads)
        sw[i] = i < kHeadCin * kCout ? w[i] : (b ? b[i - kHeadCin * kCout] : 0.f);
    __syncthreads();
    for (long long p = (long long)blockIdx.x * kHeadThreads + threadIdx.x; p < P;
         p += (long long)gridDim.x * kHeadThreads) {
        float acc[kCout];
#pragma unroll
        for (int c = 0; c < kCout; ++c) acc[c] = sw[kHeadCin * kCout + c];
        const float4 *xp = reinterpret_cast<const float4 *>(x + p * kHeadCin);
#pragma unroll
        for (int q = 0; q < kHeadCin / 4; ++q) {
            const float4 v = ld_stream_f4(reinterpret_cast<const float *>(xp + q));
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int c = 0; c < kCout; ++c) acc[c] = fmaf(xv[e], sw[(q * 4 + e) * kCout + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < kCout; ++c) y[p * kCout + c] = tanhf(acc[c]);
    }
}

template <int kCout>
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ y,
                const float *__restrict__ dy, float *__restrict__ dx, float *__restrict__ partial, long long P)
{
    constexpr int kNW = kHeadCin * kCout;
    __shared__ float sw[kNW];
    __shared__ float red[kHeadThreads / 32][kNW + kCout];
    for (int i = threadIdx.x; i < kNW; i += kHeadThreads) sw[i] = w[i];
    __syncthreads();
    float aw[kNW], ab[kCout];
#pragma unroll
    for (int i = 0; i < kNW; ++i) aw[i] = 0.f;
#pragma unroll
    for (int c = 0; c < kCout; ++c) ab[c] = 0.f;
    for (long long p = (long long)blockIdx.x * kHeadThreads + threadIdx.x; p < P;
         p += (long long)gridDim.x * kHeadThreads) {
        float g[kCout];
#pragma unroll
        for (int c = 0; c < kCout; ++c) {
            const float yy = y[p * kCout + c];
            g[c] = dy[p * kCout + c] * (1.f - yy * yy);
            ab[c] += g[c];
        }
        const float4 *xp = reinterpret_cast<const float4 *>(x + p * kHeadCin);
#pragma unroll
        for (int q = 0; q < kHeadCin / 4; ++q) {
            const float4 v = ld_stream_f4(reinterpret_cast<const float *>(xp + q));
            const float xv[4] = {v.x, v.y, v.z, v.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < kCout; ++c) {
                    s = fmaf(sw[(q * 4 + e) * kCout + c], g[c], s);
                    aw[(q * 4 + e) * kCout + c] = fmaf(xv[e], g[c], aw[(q * 4 + e) * kCout + c]);
                }
                o[e] = s;
            }
            if (dx) *reinterpret_cast<float4 *>(dx + p * kHeadCin + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kNW; ++i) {
        const float s = warp_sum(aw[i]);
        if (lane == 0) red[wp][i] = s;
    }
#pragma unroll
    for (int c = 0; c < kCout; ++c) {
        const float s = warp_sum(ab[c]);
        if (lane == 0) red[wp][kNW + c] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNW + kCout; i += kHeadThreads) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kHeadThreads / 32; ++k) s += red[k][i];
        partial[(long long)blockIdx.x * (kNW + kCout) + i] = s;
    }
}

__global__ void head_bwd_final_kernel(const float *__restrict__ partial, int blocks, int n, float *__restrict__ dw,
                                      float *__restrict__ db, int nw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < blocks; ++k) s += partial[(long long)k * n + i];
    if (i < nw) dw[i] = s; else if (db) db[i - nw] = s;
}

int head_blocks(long long P)
{
    long long b = (P + kHeadThreads - 1) / kHeadThreads;
    const long long cap = (long long)kNumSMs * 4;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

SCDA_API size_t scda_conv1x1_tanh_workspace_bytes(long long P, int Cin, int Cout)
{
    if (P <= 0 || Cin != kHeadCin || Cout < 1 || Cout > kHeadCoutMax) return 0;
    return sizeof(float) * (size_t)head_blocks(P) * (size_t)(Cin * Cout + Cout);
}

SCDA_API int scda_conv1x1_tanh_fwd(long long P, int Cin, int Cout, const float *x, const float *w, const float *b,
                                   float *y, cudaStream_t stream)
{
    if (P <= 0 || Cin != kHeadCin || Cout < 1 || Cout > kHeadCoutMax || !x || !w || !y) return 0;
    if ((uintptr_t)x % 16) return 0;
    const int g = head_blocks(P);
    switch (Cout) {
    case 1: head_fwd_kernel<1><<<g, kHeadThreads, 0, stream>>>(x, w, b, y, P); break;
    case 2: head_fwd_kernel<2><<<g, kHeadThreads, 0, stream>>>(x, w, b, y, P); break;
    case 3: head_fwd_kernel<3><<<g, kHeadThreads, 0, stream>>>(x, w, b, y, P); break;
    default: head_fwd_kernel<4><<<g, kHeadThreads, 0, stream>>>(x, w, b, y, P); break;
    }
    return scda_launch_status();
}

SCDA_API int scda_conv1x1_tanh_bwd(long long P, int Cin, int Cout, const float *x, const float *w, const float *y,
                                   const float *dy, float *dx, float *dw, float *db, void *workspace,
                                   size_t workspace_bytes, cudaStream_t stream)
{
    if (P <= 0 || Cin != kHeadCin || Cout < 1 || Cout > kHeadCoutMax || !x || !w || !y || !dy || !dw || !workspace)
        return 0;
    if (((uintptr_t)x % 16) || (dx && ((uintptr_t)dx % 16))) return 0;
    if (workspace_bytes < scda_conv1x1_tanh_workspace_bytes(P, Cin, Cout)) return 0;
    const int g = head_blocks(P);
    float *part = (float *)workspace;
    switch (Cout) {
    case 1: head_bwd_kernel<1><<<g, kHeadThreads, 0, stream>>>(x, w, y, dy, dx, part, P); break;
    case 2: head_bwd_kernel<2><<<g, kHeadThreads, 0, stream>>>(x, w, y, dy, dx, part, P); break;
    case 3: head_bwd_kernel<3><<<g, kHeadThreads, 0, stream>>>(x, w, y, dy, dx, part, P); break;
    default: head_bwd_kernel<4><<<g, kHeadThreads, 0, stream>>>(x, w, y, dy, dx, part, P); break;
    }
    const int n = Cin * Cout + Cout;
    head_bwd_final_kernel<<<(n + 127) / 128, 128, 0, stream>>>(part, g, n, dw, db, Cin * Cout);
    return scda_launch_status();
}

SCDA_API size_t scda_softmax_ce_workspace_bytes(long long M)
{
    long long blocks = (M + kCeThreads - 1) / kCeThreads;
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    if (blocks < 1) blocks = 1;
    return (size_t)blocks * 3 * sizeof(float);
}

SCDA_API int scda_softmax_ce_acc_fwd(long long M, int C, const float *logits, long long ld, const long long *targets,
                                     long long ignore_index, float *out3, void *workspace, size_t workspace_bytes,
                                     cudaStream_t stream)
{
    if (M <= 0 || C <= 0 || C > 32 || ld < C || !logits || !targets || !out3 || !workspace) return 0;
    if (workspace_bytes < scda_softmax_ce_workspace_bytes(M)) return 0;
    const int blocks = (int)(scda_softmax_ce_workspace_bytes(M) / (3 * sizeof(float)));
    softmax_ce_partial_kernel<<<blocks, kCeThreads, 0, stream>>>(logits, ld, targets, M, C, ignore_index,
                                                                 (float *)workspace);
    softmax_ce_finish_kernel<<<1, 32, 0, stream>>>((const float *)workspace, blocks, out3);
    return scda_launch_status();
}

SCDA_API int scda_softmax_ce_bwd(long long M, int C, const float *logits, long long ld, const long long *targets,
                                 long long ignore_index, const float *stats3, const float *grad_loss, float *dlogits,
                                 long long lddx, cudaStream_t stream)
{
    if (M <= 0 || C <= 0 || C > 32 || ld < C || lddx < C || !logits || !targets || !stats3 || !grad_loss || !dlogits)
        return 0;
    softmax_ce_bwd_kernel<<<ew_grid(M), kCeThreads, 0, stream>>>(logits, ld, targets, M, C, ignore_index, stats3,
                                                                 grad_loss, dlogits, lddx);
    return scda_launch_status();
}

SCDA_API int scda_rpn_fg_scores(int B, int A, int H, int W, const float *cls_nchw, float *scores, cudaStream_t stream)
{
    if (B <= 0 || A <= 0 || H <= 0 || W <= 0 || !cls_nchw || !scores) return 0;
    rpn_fg_score_kernel<<<ew_grid((long long)B * A * H * W), 256, 0, stream>>>(cls_nchw, B, A, (long long)H * W,
                                                                               scores);
    return scda_launch_status();
}
