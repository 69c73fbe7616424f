// Adam over one flat fp32 parameter buffer, fused with the data-parallel gradient scale
// and (optionally) the refresh of a bf16 shadow copy for the tensor-core kernels.
//
// Replaces, for each of the four networks, `lr_scheduler.optimizer.step()` of
// torch.optim.Adam(lr, weight_decay=1e-4) (tools/faster_rcnn_train_val.py:305-316,
// 616,635,704,750), which in torch 0.4.1 runs ~10 elementwise kernels per parameter
// tensor (40 tensors for the detector).  Here: one launch per network; p, g, m, v are each
// read once and p, m, v written once (28 B per parameter, HBM bound: 3.8 GB for the
// detector's 136.85 M parameters).
//
// Update rule = torch 0.4.1's (the version the reference pins, README.md:16):
//   g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, __nv_bfloat16 *__restrict__ shadow, long long n, float grad_scale,
            float lr_t, float b1, float b2, float eps, float wd, const float *__restrict__ lr_t_dev)
{
    if (lr_t_dev) lr_t = __ldg(lr_t_dev);     // step size kept on the device (CUDA-graph replays)
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4 *>(p + i);
            const float4 gg = ld_stream_f4(g + i);
            float4 mm = *reinterpret_cast<float4 *>(m + i);
            float4 vv = *reinterpret_cast<float4 *>(v + i);
            float *pa = &pp.x, *ma = &mm.x, *va = &vv.x;
            const float *ga = &gg.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = fmaf(wd, pa[k], ga[k] * grad_scale);
                ma[k] = fmaf(b1, ma[k], (1.f - b1) * gk);
                va[k] = fmaf(b2, va[k], (1.f - b2) * gk * gk);
                pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps);
            }
            *reinterpret_cast<float4 *>(p + i) = pp;
            *reinterpret_cast<float4 *>(m + i) = mm;
            *reinterpret_cast<float4 *>(v + i) = vv;
            if (shadow) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
                uint2 pk;
                pk.x = *reinterpret_cast<unsigned *>(&lo);
                pk.y = *reinterpret_cast<unsigned *>(&hi);
                *reinterpret_cast<uint2 *>(shadow + i) = pk;
            }
        } else {
            for (long long j = i; j < n; ++j) {
                const float gk = fmaf(wd, p[j], g[j] * grad_scale);
                m[j] = fmaf(b1, m[j], (1.f - b1) * gk);
                v[j] = fmaf(b2, v[j], (1.f - b2) * gk * gk);
                p[j] -= lr_t * m[j] / (sqrtf(v[j]) + eps);
                if (shadow) shadow[j] = __float2bfloat16_rn(p[j]);
            }
        }
    }
}

}  // namespace

SCDA_API int scda_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                            void *bf16_shadow, long long n, int step, float lr, float beta1,
                            float beta2, float eps, float weight_decay, float grad_scale,
                            const float *lr_t_dev, cudaStream_t stream)
{
    if (n < 0 || (step < 1 && !lr_t_dev)) return 0;
    if (n == 0) return 1;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return 0;
    if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) return 0;
    if (bf16_shadow && (uintptr_t)bf16_shadow % 8) return 0;
    float lr_t = 0.f;
    if (!lr_t_dev) {
        const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
        lr_t = (float)((double)lr * sqrt(bc2) / bc1);
    }
    long long want = (n / 4 + 255) / 256;
    const int grid = (int)(want < (long long)kNumSMs * 8 ? (want < 1 ? 1 : want) : (long long)kNumSMs * 8);
    adam_kernel<<<grid, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq,
                                          (__nv_bfloat16 *)bf16_shadow, n, grad_scale, lr_t, beta1,
                                          beta2, eps, weight_decay, lr_t_dev);
    return scda_launch_status();
}
