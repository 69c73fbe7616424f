// Sigmoid and softmax focal losses, forward and backward, sm_100a.
//
// Replaces the reference's SigmoidFocalLoss{,Gradient}Kernel
// (extensions/_focal_loss/src/cuda/focal_loss_sigmoid_kernel.cu:12-81) and
// SpatialSoftmax / SoftmaxFocalLoss{,GradientWeight,Gradient}Kernel
// (focal_loss_softmax_kernel.cu:12-100).  The per-element formulas keep the
// reference's float/double promotions (double literals in the source) so the
// two agree to rounding of the libdevice transcendentals.
//
// Differences in structure, not in results:
//   - softmax forward is one kernel (row softmax + loss) instead of two, and
//     backward one kernel (weight + gradient) instead of two: the per-row
//     `buff` value is produced and consumed in registers (still written out
//     when the caller passes the buffer, which the Section A ABI requires);
//   - scda_*_focal_loss_sum fuse the reference's Python-side `losses.sum()`
//     (focal_loss.py:46,118) into the same pass: warp shuffle -> one
//     red.global.add per CTA.
#include <float.h>

#include <atomic>

#include "common.cuh"

namespace {

constexpr int kFocalThreads = 256;

struct FocalScale {
    float zn, zp;
};

__device__ __forceinline__ FocalScale focal_scale(float weight_pos, float alpha)
{
    const float Np = (float)fmax((double)weight_pos, 1.0);
    FocalScale s;
    s.zn = (float)((1.0 - (double)alpha) / (double)Np);
    s.zp = alpha / Np;
    return s;
}

// -x*[x>=0] - log(1 + exp(x - 2x*[x>=0])) : log(1 - sigmoid(x)), in the
// reference's double/float mix (focal_loss_sigmoid_kernel.cu:38-41)
__device__ __forceinline__ double log_one_minus_sigmoid(float x)
{
    const double pos = (double)(x >= 0);
    const float e = expf((float)((double)x - 2.0 * (double)x * pos));
    return -1.0 * (double)x * pos - (double)logf((float)(1.0 + (double)e));
}

// powf(x, gamma) for x in [0, 1]; gamma = 2 (the published setting) is one correctly rounded product
// (libdevice powf: ~60 instructions and up to 2 ulp)
__device__ __forceinline__ float pow_gamma(float x, float gamma)
{
    return gamma == 2.f ? __fmul_rn(x, x) : powf(x, gamma);
}

__device__ __forceinline__ float sigmoid_focal_elem(float x, int t, int d, float gamma,
                                                    FocalScale s)
{
    const float c1 = (float)(t == d + 1);
    const float c2 = (float)((t != -1) & (t != d + 1));
    const float p = (float)(1.0 / (1.0 + (double)expf(-x)));
    const float term1 = pow_gamma((float)(1.0 - (double)p), gamma) * logf(fmaxf(p, FLT_MIN));
    const float term2 = (float)((double)pow_gamma(p, gamma) * log_one_minus_sigmoid(x));
    float l = 0.f;
    l += -c1 * term1 * s.zp;
    l += -c2 * term2 * s.zn;
    return l;
}

__device__ __forceinline__ float sigmoid_focal_grad_elem(float x, int t, int d, float gamma,
                                                         FocalScale s)
{
    const float c1 = (float)(t == d + 1);
    const float c2 = (float)((t != -1) & (t != d + 1));
    const float p = (float)(1.0 / (1.0 + (double)expf(-x)));
    const float term1 = (float)((double)powf((float)(1.0 - (double)p), gamma) *
                                (1.0 - (double)p - (double)(p * gamma * logf(fmaxf(p, FLT_MIN)))));
    const float term2 = (float)((double)powf(p, gamma) *
                                (log_one_minus_sigmoid(x) * (1.0 - (double)p) * (double)gamma - (double)p));
    float g = 0.f;
    g += -c1 * s.zp * term1;
    g += -c2 * s.zn * term2;
    return g;
}

// The fused total, in ONE launch and in a fixed order: every CTA leaves its partial in a slot of module
// memory and takes a ticket; the CTA that draws the last ticket adds the partials in index order, overwrites
// loss_sum[0] and re-arms the ticket.  (A memset of loss_sum + one floating-point atomic per CTA was a second
// graph node in front of a 5 us kernel and made the total depend on the arrival order.)  Concurrent calls
// take different slots: the host hands them out round-robin, kFocalSlots calls can be in flight at once.
constexpr int kFocalSlots = 64;
constexpr int kFocalMaxGrid = kNumSMs * 8;
__device__ float g_focal_partial[kFocalSlots][kFocalMaxGrid];
__device__ unsigned int g_focal_ticket[kFocalSlots];

__device__ __forceinline__ void cta_accumulate(float v, float *dst, int slot)
{
    __shared__ float s_part[kFocalThreads / 32];
    __shared__ bool s_last;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kFocalThreads / 32; ++w) t += s_part[w];
        g_focal_partial[slot][blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(&g_focal_ticket[slot], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float t = 0.f;                                     // fixed order: thread k takes partials k, k + 256, ...
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kFocalThreads) t += __ldcg(&g_focal_partial[slot][i]);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kFocalThreads / 32; ++w) tot += s_part[w];
        *dst = tot;
        g_focal_ticket[slot] = 0u;
    }
}

template <bool kSum>
__global__ void __launch_bounds__(kFocalThreads)
sigmoid_focal_fwd_kernel(int N, const float *__restrict__ logits, const int *__restrict__ targets,
                         float weight_pos, float gamma, float alpha, int num_classes,
                         float *__restrict__ losses, float *__restrict__ loss_sum, int slot)
{
    const FocalScale s = focal_scale(weight_pos, alpha);
    float acc = 0.f;
    for (int i = blockIdx.x * kFocalThreads + threadIdx.x; i < N; i += gridDim.x * kFocalThreads) {
        const int row = i / num_classes, d = i - row * num_classes;
        const float l = sigmoid_focal_elem(logits[i], __ldg(targets + row), d, gamma, s);
        if (losses) losses[i] = l;
        acc += l;
    }
    if (kSum) cta_accumulate(acc, loss_sum, slot);
}

__global__ void __launch_bounds__(kFocalThreads)
sigmoid_focal_bwd_kernel(int N, const float *__restrict__ logits, const int *__restrict__ targets,
                         float *__restrict__ dX, float weight_pos, float gamma, float alpha,
                         int num_classes)
{
    const FocalScale s = focal_scale(weight_pos, alpha);
    for (int i = blockIdx.x * kFocalThreads + threadIdx.x; i < N; i += gridDim.x * kFocalThreads) {
        const int row = i / num_classes, d = i - row * num_classes;
        dX[i] = sigmoid_focal_grad_elem(logits[i], __ldg(targets + row), d, gamma, s);
    }
}

__device__ __forceinline__ float softmax_row_z(int label, float alpha, float Np)
{
    return (float)(label == 0) * (1 - alpha) / Np + (float)(label >= 1) * alpha / Np;
}

// thread per row (rows are short: num_classes ~ 2..81)
template <bool kSum>
__global__ void __launch_bounds__(kFocalThreads)
softmax_focal_fwd_kernel(int rows, const float *__restrict__ logits,
                         const int *__restrict__ targets, float weight_pos, float gamma,
                         float alpha, int num_classes, float *__restrict__ losses,
                         float *__restrict__ priors, float *__restrict__ loss_sum, int slot)
{
    const float Np = (float)fmax((double)weight_pos, 1.0);
    float acc = 0.f;
    for (int i = blockIdx.x * kFocalThreads + threadIdx.x; i < rows; i += gridDim.x * kFocalThreads) {
        const float *x = logits + (long long)i * num_classes;
        float *P = priors + (long long)i * num_classes;
        float mx = -FLT_MAX;
        for (int c = 0; c < num_classes; ++c) mx = fmaxf(mx, x[c]);
        float sum = 0.f;
        for (int c = 0; c < num_classes; ++c) {
            const float e = expf(x[c] - mx);
            P[c] = e;
            sum += e;
        }
        const int label = targets[i];
        float pl = 0.f;
        for (int c = 0; c < num_classes; ++c) {
            const float p = __fdiv_rn(P[c], sum);
            P[c] = p;
            if (c == label) pl = p;
        }
        float l = 0.f;
        if (label >= 0)
            l = -(powf((float)(1.0 - (double)pl), gamma) * logf(fmaxf(pl, FLT_MIN))) *
                softmax_row_z(label, alpha, Np);
        if (losses) losses[i] = l;
        acc += l;
    }
    if (kSum) cta_accumulate(acc, loss_sum, slot);
}

__global__ void __launch_bounds__(kFocalThreads)
softmax_focal_bwd_kernel(int rows, const int *__restrict__ targets, float *__restrict__ dX,
                         float weight_pos, float gamma, float alpha, int num_classes,
                         const float *__restrict__ priors, float *__restrict__ buff)
{
    const float Np = (float)fmax((double)weight_pos, 1.0);
    for (int i = blockIdx.x * kFocalThreads + threadIdx.x; i < rows; i += gridDim.x * kFocalThreads) {
        const int label = targets[i];
        const float *P = priors + (long long)i * num_classes;
        float w = 0.f;
        if (label >= 0) {
            const float p = P[label];
            const float onemp = (float)(1.0 - (double)p);
            w = (-powf(onemp, gamma) + gamma * powf(onemp, gamma - 1) * p * logf(fmaxf(p, FLT_MIN))) *
                softmax_row_z(label, alpha, Np);
        }
        if (buff) buff[i] = w;
        const float c1 = (float)(label >= 0);
        for (int c = 0; c < num_classes; ++c)
            dX[(long long)i * num_classes + c] = c1 * w * ((float)(label == c) - P[c]);
    }
}

int focal_grid(long long work)
{
    long long g = (work + kFocalThreads - 1) / kFocalThreads;
    const long long cap = kFocalMaxGrid;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// the fused-sum forms: two CTAs per SM (every CTA costs one ticket and one partial in the final pass)
int focal_sum_grid(long long work)
{
    const int g = focal_grid(work);
    return g > 2 * kNumSMs ? 2 * kNumSMs : g;
}

int focal_slot()
{
    static std::atomic<unsigned> next{0};
    return (int)(next.fetch_add(1u, std::memory_order_relaxed) % kFocalSlots);
}

bool focal_args_ok(int N, int num_classes, const void *a, const void *b, const void *c)
{
    return N >= 0 && num_classes > 0 && N % num_classes == 0 && (N == 0 || (a && b && c));
}

}  // namespace

SCDA_API int SigmoidFocalLossForwardLaucher(const int N, const float *logits, const int *targets,
                                            const float weight_pos, const float gamma,
                                            const float alpha, const int num_classes,
                                            float *losses, cudaStream_t stream)
{
    if (!focal_args_ok(N, num_classes, logits, targets, losses)) return 0;
    if (N == 0) return 1;
    sigmoid_focal_fwd_kernel<false><<<focal_grid(N), kFocalThreads, 0, stream>>>(
        N, logits, targets, weight_pos, gamma, alpha, num_classes, losses, nullptr, 0);
    return scda_launch_status();
}

SCDA_API int scda_sigmoid_focal_loss_sum(const int N, const float *logits, const int *targets,
                                         const float weight_pos, const float gamma,
                                         const float alpha, const int num_classes, float *losses,
                                         float *loss_sum, cudaStream_t stream)
{
    if (!focal_args_ok(N, num_classes, logits, targets, loss_sum) || !loss_sum) return 0;
    if (N == 0) {
        cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(float), stream);
        return e == cudaSuccess ? 1 : -(int)e;
    }
    sigmoid_focal_fwd_kernel<true><<<focal_sum_grid(N), kFocalThreads, 0, stream>>>(
        N, logits, targets, weight_pos, gamma, alpha, num_classes, losses, loss_sum, focal_slot());
    return scda_launch_status();
}

SCDA_API int SigmoidFocalLossBackwardLaucher(const int N, const float *logits, const int *targets,
                                             float *dX_data, const float weight_pos,
                                             const float gamma, const float alpha,
                                             const int num_classes, cudaStream_t stream)
{
    if (!focal_args_ok(N, num_classes, logits, targets, dX_data)) return 0;
    if (N == 0) return 1;
    sigmoid_focal_bwd_kernel<<<focal_grid(N), kFocalThreads, 0, stream>>>(
        N, logits, targets, dX_data, weight_pos, gamma, alpha, num_classes);
    return scda_launch_status();
}

SCDA_API int SoftmaxFocalLossForwardLaucher(const int N, const float *logits, const int *targets,
                                            const float weight_pos, const float gamma,
                                            const float alpha, const int num_classes,
                                            float *losses, float *priors, cudaStream_t stream)
{
    if (!focal_args_ok(N, num_classes, logits, targets, losses) || (N > 0 && !priors)) return 0;
    if (N == 0) return 1;
    const int rows = N / num_classes;
    softmax_focal_fwd_kernel<false><<<focal_grid(rows), kFocalThreads, 0, stream>>>(
        rows, logits, targets, weight_pos, gamma, alpha, num_classes, losses, priors, nullptr, 0);
    return scda_launch_status();
}

SCDA_API int scda_softmax_focal_loss_sum(const int N, const float *logits, const int *targets,
                                         const float weight_pos, const float gamma,
                                         const float alpha, const int num_classes, float *losses,
                                         float *priors, float *loss_sum, cudaStream_t stream)
{
    if (!focal_args_ok(N, num_classes, logits, targets, priors) || !loss_sum) return 0;
    if (N == 0) {
        cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(float), stream);
        return e == cudaSuccess ? 1 : -(int)e;
    }
    const int rows = N / num_classes;
    softmax_focal_fwd_kernel<true><<<focal_sum_grid(rows), kFocalThreads, 0, stream>>>(
        rows, logits, targets, weight_pos, gamma, alpha, num_classes, losses, priors, loss_sum, focal_slot());
    return scda_launch_status();
}

SCDA_API int SoftmaxFocalLossBackwardLaucher(const int N, const float *logits, const int *targets,
                                             float *dX_data, const float weight_pos,
                                             const float gamma, const float alpha,
                                             const int num_classes, const float *priors,
                                             float *buff, cudaStream_t stream)
{
    (void)logits;
    if (!focal_args_ok(N, num_classes, priors, targets, dX_data)) return 0;
    if (N == 0) return 1;
    const int rows = N / num_classes;
    softmax_focal_bwd_kernel<<<focal_grid(rows), kFocalThreads, 0, stream>>>(
        rows, targets, dX_data, weight_pos, gamma, alpha, num_classes, priors, buff);
    return scda_launch_status();
}

SCDA_API int scda_abi_version(void) { return 2; }
