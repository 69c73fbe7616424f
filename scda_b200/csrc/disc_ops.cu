// Discriminator layers that are NOT tensor-core shaped, and the glue of the stride-2 tensor-core convolutions
// (conv_halo.cu kS2, gemm_tc.cu: scda_conv3x3_s2_*).
//
// Replaces, in the image discriminators `GAN_dis_AE` and the feature discriminator `GAN_dis_AE_patch` of the
// reference (models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:270-333,
// common_net.py:205-245 ResDis_cluster, :251-261 LeakyReLUConv2d), the cuDNN / torch calls behind
//   * the first layer, nn.Conv2d(3, 32, 3, stride 2, padding 1) + LeakyReLU: 27 inputs per output — a direct
//     CUDA-core kernel (forward, weight / bias gradient, input gradient);
//   * the 1x1 head nn.Conv2d(128, 1, 1): a per-pixel dot product, whose backward also applies the LeakyReLU
//     gradient of the layer below (the gradient leaves in the dtype / masking the next MMA wants);
//   * nn.BatchNorm2d (training mode: batch statistics over the 4 cluster images, running statistics updated)
//     + LeakyReLU of the feature discriminator, forward and backward;
//   * the weight layout of the stride-2 tensor-core form, [Cout][3x3 taps][row phase, column phase, C], and the
//     fold of its weight gradient back to [Cout][3][3][C].
// All HBM / latency bound and small (<= 8 MB per tensor); reductions are two-stage in a fixed order unless
// noted (red.add of per-block partial sums: reproducible to fp32 rounding, like the bias gradients of nhwc_ops.cu).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// ------------------------------------------------------------------ s2d weight layout
// src bf16 [Cout][3][3][C] -> wd bf16 [Cout][9][4C]: wd[o][(a+1)*3 + (b+1)][(py*2 + px)*C + c] =
// src[o][2a + py + 1][2b + px + 1][c] where that tap exists (a, b in {-1, 0}), else 0
__global__ void __launch_bounds__(256)
s2_weights_kernel(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ wd, int Cout, int C)
{
    const long long total = (long long)Cout * 9 * 4 * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int ph = (int)((i / C) % 4);
        const int tap = (int)((i / (4 * C)) % 9);
        const long long o = i / (36ll * C);
        const int a = tap / 3 - 1, b = tap % 3 - 1, py = ph >> 1, px = ph & 1;
        const int r = 2 * a + py + 1, s = 2 * b + px + 1;
        const bool ok = a <= 0 && b <= 0 && r >= 0 && r < 3 && s >= 0 && s < 3;
        wd[i] = ok ? src[((o * 3 + r) * 3 + s) * C + c] : __float2bfloat16_rn(0.f);
    }
}

// partial fp32 [splits][Cout][9][4C] (taps 0, 1, 3, 4 written) -> dw fp32 [Cout][3][3][C] (+)=
// a thread owns one output and walks the slabs, four independent loads in flight; consecutive threads read
// consecutive addresses of a slab (a CTA: 1 KB per slab step), which is what keeps this pass at HBM speed
__global__ void __launch_bounds__(256)
s2_wgrad_gather_kernel(const float *__restrict__ part, int splits, float *__restrict__ dw, int Cout, int C,
                       int accumulate)
{
    const long long total = (long long)Cout * 9 * C, slab = (long long)Cout * 9 * 4 * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int s = (int)((i / C) % 3), r = (int)((i / (3 * C)) % 3);
        const long long o = i / (9ll * C);
        // r = 2a + py + 1 with a in {-1, 0}: r = 0 -> (a, py) = (-1, 1); r = 1 -> (0, 0); r = 2 -> (0, 1)
        const int a = r == 0 ? -1 : 0, py = r == 1 ? 0 : 1;
        const int b = s == 0 ? -1 : 0, px = s == 1 ? 0 : 1;
        const float *src = part + ((o * 9 + (a + 1) * 3 + (b + 1)) * 4 + py * 2 + px) * C + c;
        float acc = accumulate ? dw[i] : 0.f;
        int k = 0;
        for (; k + 3 < splits; k += 4) {
            const float v0 = __ldg(src + (long long)k * slab), v1 = __ldg(src + (long long)(k + 1) * slab);
            const float v2 = __ldg(src + (long long)(k + 2) * slab), v3 = __ldg(src + (long long)(k + 3) * slab);
            acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; k < splits; ++k) acc += __ldg(src + (long long)k * slab);
        dw[i] = acc;
    }
}

// ------------------------------------------------------------------ first layer: Conv2d(Cin = 3, 32, 3, s 2, p 1) + LeakyReLU
constexpr int kL1Cin = 3, kL1Cout = 32, kL1K = 27;

// x fp32 [N, 3, H, W] with arbitrary strides (NCHW crops or channels-last reconstructions);
// w fp32 [32][3][3][3] (o, r, s, c); y bf16 or fp32 NHWC [N, H/2, W/2, 32].  One thread per output pixel.
template <typename TY>
__global__ void __launch_bounds__(128)
l1_fwd_kernel(const float *__restrict__ x, long long sn, long long sc, long long sh, long long sw, int N, int H, int W,
              const float *__restrict__ w, const float *__restrict__ bias, float slope, TY *__restrict__ y)
{
    __shared__ float sw_[kL1Cout * kL1K + kL1Cout];
    for (int i = threadIdx.x; i < kL1Cout * kL1K; i += blockDim.x) sw_[i] = w[i];
    for (int i = threadIdx.x; i < kL1Cout; i += blockDim.x) sw_[kL1Cout * kL1K + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int j = (int)(p % Wo), i = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    float in[kL1K];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int h = 2 * i + r - 1, ww = 2 * j + s - 1;
            const bool ok = h >= 0 && h < H && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                in[(r * 3 + s) * 3 + c] = ok ? __ldg(x + n * sn + c * sc + h * sh + ww * sw) : 0.f;
        }
    float out[kL1Cout];
#pragma unroll
    for (int o = 0; o < kL1Cout; ++o) {
        float acc = sw_[kL1Cout * kL1K + o];
#pragma unroll
        for (int k = 0; k < kL1K; ++k) acc = fmaf(in[k], sw_[o * kL1K + k], acc);
        out[o] = lrelu(acc, slope);
    }
    if (sizeof(TY) == 2) {
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(y) + p * kL1Cout;
#pragma unroll
        for (int o = 0; o < kL1Cout; o += 8) {
            uint4 pk;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(out[o], out[o + 1]), b1 = __floats2bfloat162_rn(out[o + 2], out[o + 3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(out[o + 4], out[o + 5]), b3 = __floats2bfloat162_rn(out[o + 6], out[o + 7]);
            pk.x = *reinterpret_cast<uint32_t *>(&b0); pk.y = *reinterpret_cast<uint32_t *>(&b1);
            pk.z = *reinterpret_cast<uint32_t *>(&b2); pk.w = *reinterpret_cast<uint32_t *>(&b3);
            *reinterpret_cast<uint4 *>(dst + o) = pk;
        }
    } else {
        float *dst = reinterpret_cast<float *>(y) + p * kL1Cout;
#pragma unroll
        for (int o = 0; o < kL1Cout; o += 4)
            *reinterpret_cast<float4 *>(dst + o) = make_float4(out[o], out[o + 1], out[o + 2], out[o + 3]);
    }
}

// weight + bias gradient: g bf16 NHWC [N, Ho, Wo, 32] = gradient w.r.t. the PRE-activation (the LeakyReLU
// gradient was applied upstream).  Block b owns the pixel range [b * per, (b+1) * per); thread (o, k) with
// k = 0..26 the weight taps and k = 27 the bias; partial[b][o * 28 + k].
constexpr int kL1Tile = 64;
template <typename TG>
__global__ void __launch_bounds__(kL1Cout * 28)
l1_bwd_w_kernel(const float *__restrict__ x, long long sn, long long sc, long long sh, long long sw, int N, int H,
                int W, const TG *__restrict__ g, long long per, float *__restrict__ partial)
{
    __shared__ float sg[kL1Tile][kL1Cout + 1];
    __shared__ float sx[kL1Tile][28];
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(total, p0 + per);
    const int o = threadIdx.x / 28, k = threadIdx.x % 28;
    float acc = 0.f;
    for (long long base = p0; base < p1; base += kL1Tile) {
        const int cnt = (int)min((long long)kL1Tile, p1 - base);
        for (int t = threadIdx.x; t < kL1Tile * kL1Cout; t += blockDim.x) {
            const int pp = t / kL1Cout, oo = t % kL1Cout;
            float v = 0.f;
            if (pp < cnt) {
                if (sizeof(TG) == 2) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(g)[(base + pp) * kL1Cout + oo]);
                else v = reinterpret_cast<const float *>(g)[(base + pp) * kL1Cout + oo];
            }
            sg[pp][oo] = v;
        }
        for (int t = threadIdx.x; t < kL1Tile * 28; t += blockDim.x) {
            const int pp = t / 28, kk = t % 28;
            float v = 0.f;
            if (pp < cnt) {
                if (kk == 27) v = 1.f;
                else {
                    const long long p = base + pp;
                    const int j = (int)(p % Wo), i = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
                    const int c = kk % 3, rs = kk / 3, r = rs / 3, s = rs % 3;
                    const int h = 2 * i + r - 1, ww = 2 * j + s - 1;
                    if (h >= 0 && h < H && ww >= 0 && ww < W) v = __ldg(x + n * sn + c * sc + h * sh + ww * sw);
                }
            }
            sx[pp][kk] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int pp = 0; pp < kL1Tile; ++pp) acc = fmaf(sg[pp][o], sx[pp][k], acc);
        __syncthreads();
    }
    partial[(long long)blockIdx.x * (kL1Cout * 28) + threadIdx.x] = acc;
}

// partial [blocks][32 * 28] -> dw [32][27] (+)=, db [32] (+)=
__global__ void l1_bwd_w_finish_kernel(const float *__restrict__ partial, int blocks, float *__restrict__ dw,
                                       float *__restrict__ db, int accumulate)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kL1Cout * 28) return;
    float acc = 0.f;
    for (int b = 0; b < blocks; ++b) acc += partial[(long long)b * (kL1Cout * 28) + t];
    const int o = t / 28, k = t % 28;
    if (k == 27) {
        if (db) db[o] = accumulate ? db[o] + acc : acc;
    } else {
        dw[o * kL1K + k] = accumulate ? dw[o * kL1K + k] + acc : acc;
    }
}

// input gradient: dx fp32 [N, H, W, 3] channels-last; one thread per input pixel
template <typename TG>
__global__ void __launch_bounds__(128)
l1_bwd_x_kernel(const TG *__restrict__ g, const float *__restrict__ w, int N, int H, int W, float *__restrict__ dx)
{
    __shared__ float sw_[kL1Cout * kL1K];
    for (int i = threadIdx.x; i < kL1Cout * kL1K; i += blockDim.x) sw_[i] = w[i];
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * H * W;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int ww = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int t = h + 1 - r;                 // = 2 i
        if (t < 0 || (t & 1) || (t >> 1) >= Ho) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int u = ww + 1 - s;
            if (u < 0 || (u & 1) || (u >> 1) >= Wo) continue;
            const long long q = (((long long)n * Ho + (t >> 1)) * Wo + (u >> 1)) * kL1Cout;
#pragma unroll 8
            for (int o = 0; o < kL1Cout; ++o) {
                float gv;
                if (sizeof(TG) == 2) gv = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(g)[q + o]);
                else gv = reinterpret_cast<const float *>(g)[q + o];
                const float *wk = sw_ + o * kL1K + (r * 3 + s) * 3;
                a0 = fmaf(gv, wk[0], a0); a1 = fmaf(gv, wk[1], a1); a2 = fmaf(gv, wk[2], a2);
            }
        }
    }
    dx[p * 3] = a0; dx[p * 3 + 1] = a1; dx[p * 3 + 2] = a2;
}

// ------------------------------------------------------------------ 1x1 head: Conv2d(C, 1, 1)
// x bf16 / fp32 [P, C] -> out fp32 [P] = bias + x . w; one warp per pixel
template <typename TX>
__global__ void __launch_bounds__(256)
head_dot_fwd_kernel(const TX *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, long long P,
                    int C, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= P) return;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
        float v;
        if (sizeof(TX) == 2) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(x)[warp * C + c]);
        else v = reinterpret_cast<const float *>(x)[warp * C + c];
        acc = fmaf(v, w[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[warp] = acc + (bias ? bias[0] : 0.f);
}

// g fp32 [P] -> dx [P, C] = g w * (x > 0 ? 1 : slope)  (the gradient w.r.t. the pre-activation of the layer that
// produced x = LeakyReLU(...)), dw [C] += sum_p g x, db += sum g.  Block = 32 pixel lanes x C... one thread per
// (pixel row of the block, channel); partial sums per block meet in dw / db through red.add.
template <typename TX>
__global__ void __launch_bounds__(256)
head_dot_bwd_kernel(const TX *__restrict__ x, const float *__restrict__ w, const float *__restrict__ g, long long P,
                    int C, float slope, TX *__restrict__ dx, float *__restrict__ dw, float *__restrict__ db)
{
    // thread t handles channel t % C of rows (t / C) + k * (256 / C)
    const int c = threadIdx.x % C, rl = threadIdx.x / C, rows = 256 / C;
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(P, p0 + per);
    const float wc = w[c];
    float aw = 0.f, ab = 0.f;
    for (long long p = p0 + rl; p < p1; p += rows) {
        float xv;
        if (sizeof(TX) == 2) xv = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(x)[p * C + c]);
        else xv = reinterpret_cast<const float *>(x)[p * C + c];
        const float gp = g[p];
        const float d = gp * wc * (xv > 0.f ? 1.f : slope);
        if (dx) {
            if (sizeof(TX) == 2) reinterpret_cast<__nv_bfloat16 *>(dx)[p * C + c] = __float2bfloat16_rn(d);
            else reinterpret_cast<float *>(dx)[p * C + c] = d;
        }
        aw = fmaf(gp, xv, aw);
        if (c == 0) ab += gp;
    }
    if (dw) red_add_f32(dw + c, aw);
    if (db && c == 0) red_add_f32(db, ab);
}

// ------------------------------------------------------------------ BatchNorm2d (training) + LeakyReLU
// x fp32 [P, C] (P = N * H * W pixels) -> y = lrelu(gamma * (x - mean) * rstd + beta), batch statistics (biased
// variance) over the P rows; running_mean / running_var updated with `momentum` (unbiased variance), as
// nn.BatchNorm2d.  Two launches each way, grid = (32-channel groups, pixel chunks):
//   partial: a CTA (32 channels x 32 row lanes) reduces its pixel chunk -> partial[chunk][2][C]
//   apply  : every CTA adds the chunk partials of its channels in chunk order (a few hundred floats out of L2:
//            cheaper than a grid-wide barrier or a ticket, and the same bits in every CTA), then transforms its
//            own pixel chunk; the CTAs of chunk 0 write the per-channel outputs.
// (One CTA per 32 channels walking all pixels — 2 to 16 CTAs on 148 SMs — was 40 us forward / up to 180 us
//  backward per layer on the critical path of the discriminator updates.)
constexpr int kBnMaxChunks = 64;

__device__ __forceinline__ void bn_chunk(long long P, int chunks, int chunk, long long *p0, long long *p1)
{
    const long long per = (P + chunks - 1) / chunks;
    *p0 = min(P, per * chunk);
    *p1 = min(P, *p0 + per);
}

__global__ void __launch_bounds__(1024)
bn_fwd_partial_kernel(const float *__restrict__ x, long long P, int C, float *__restrict__ partial)
{
    __shared__ float s1[32][33], s2[32][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    long long p0, p1;
    bn_chunk(P, gridDim.y, blockIdx.y, &p0, &p1);
    const float shift = c < C ? x[c] : 0.f;
    float a = 0.f, b = 0.f;
    if (c < C)
        for (long long p = p0 + rl; p < p1; p += 32) {
            const float d = x[p * C + c] - shift;
            a += d;
            b = fmaf(d, d, b);
        }
    s1[rl][cl] = a;
    s2[rl][cl] = b;
    __syncthreads();
    if (rl == 0 && c < C) {
        for (int k = 1; k < 32; ++k) { a += s1[k][cl]; b += s2[k][cl]; }
        partial[((long long)blockIdx.y * 2 + 0) * C + c] = a;
        partial[((long long)blockIdx.y * 2 + 1) * C + c] = b;
    }
}

template <typename TY>
__global__ void __launch_bounds__(1024)
bn_lrelu_fwd_apply_kernel(const float *__restrict__ x, long long P, int C, const float *__restrict__ gamma,
                          const float *__restrict__ beta, float eps, float slope, float momentum,
                          float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ mean,
                          float *__restrict__ rstd, TY *__restrict__ y, const float *__restrict__ partial, int chunks)
{
    __shared__ float sm[32], sr[32];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    if (rl == 0 && c < C) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < chunks; ++k) {
            a += partial[((long long)k * 2 + 0) * C + c];
            b += partial[((long long)k * 2 + 1) * C + c];
        }
        const float shift = x[c];
        const float inv = 1.f / (float)P;
        const float m = a * inv;
        const float var = fmaxf(b * inv - m * m, 0.f);
        const float mu = shift + m, rs = rsqrtf(var + eps);
        sm[cl] = mu; sr[cl] = rs;
        if (blockIdx.y == 0) {
            mean[c] = mu; rstd[c] = rs;
            if (running_mean) {
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
                const float unb = P > 1 ? var * ((float)P / (float)(P - 1)) : var;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
            }
        }
    }
    __syncthreads();
    if (c >= C) return;
    long long p0, p1;
    bn_chunk(P, gridDim.y, blockIdx.y, &p0, &p1);
    const float mu = sm[cl], ga = gamma[c] * sr[cl], be = beta[c];
    for (long long p = p0 + rl; p < p1; p += 32) {
        const float v = lrelu(fmaf(x[p * C + c] - mu, ga, be), slope);
        if (sizeof(TY) == 2) reinterpret_cast<__nv_bfloat16 *>(y)[p * C + c] = __float2bfloat16_rn(v);
        else reinterpret_cast<float *>(y)[p * C + c] = v;
    }
}

// dy [P, C] (bf16 or fp32) = gradient w.r.t. y; g = dy * lrelu'(gamma xhat + beta);
// dgamma += sum g xhat, dbeta += sum g, dx = gamma rstd (g - mean(g) - xhat mean(g xhat))
template <typename TDy>
__global__ void __launch_bounds__(1024)
bn_bwd_partial_kernel(const float *__restrict__ x, const TDy *__restrict__ dy, long long P, int C,
                      const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ mean,
                      const float *__restrict__ rstd, float slope, float *__restrict__ partial)
{
    __shared__ float s1[32][33], s2[32][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    long long p0, p1;
    bn_chunk(P, gridDim.y, blockIdx.y, &p0, &p1);
    const float mu = c < C ? mean[c] : 0.f, rs = c < C ? rstd[c] : 0.f, ga = c < C ? gamma[c] : 0.f,
                be = c < C ? beta[c] : 0.f;
    float a = 0.f, b = 0.f;
    if (c < C)
        for (long long p = p0 + rl; p < p1; p += 32) {
            const float xh = (x[p * C + c] - mu) * rs;
            float gy;
            if (sizeof(TDy) == 2) gy = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(dy)[p * C + c]);
            else gy = reinterpret_cast<const float *>(dy)[p * C + c];
            const float g = gy * (fmaf(xh, ga, be) > 0.f ? 1.f : slope);
            a += g;
            b = fmaf(g, xh, b);
        }
    s1[rl][cl] = a;
    s2[rl][cl] = b;
    __syncthreads();
    if (rl == 0 && c < C) {
        for (int k = 1; k < 32; ++k) { a += s1[k][cl]; b += s2[k][cl]; }
        partial[((long long)blockIdx.y * 2 + 0) * C + c] = a;
        partial[((long long)blockIdx.y * 2 + 1) * C + c] = b;
    }
}

template <typename TDy, typename TDx>
__global__ void __launch_bounds__(1024)
bn_lrelu_bwd_apply_kernel(const float *__restrict__ x, const TDy *__restrict__ dy, long long P, int C,
                          const float *__restrict__ gamma, const float *__restrict__ beta,
                          const float *__restrict__ mean, const float *__restrict__ rstd, float slope,
                          TDx *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta, int accumulate,
                          const float *__restrict__ partial, int chunks)
{
    __shared__ float sa[32], sb[32];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    if (rl == 0 && c < C) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < chunks; ++k) {
            a += partial[((long long)k * 2 + 0) * C + c];
            b += partial[((long long)k * 2 + 1) * C + c];
        }
        if (blockIdx.y == 0) {
            dbeta[c] = accumulate ? dbeta[c] + a : a;
            dgamma[c] = accumulate ? dgamma[c] + b : b;
        }
        sa[cl] = a / (float)P;
        sb[cl] = b / (float)P;
    }
    __syncthreads();
    if (c >= C || !dx) return;
    long long p0, p1;
    bn_chunk(P, gridDim.y, blockIdx.y, &p0, &p1);
    const float mu = mean[c], rs = rstd[c], ga = gamma[c], be = beta[c];
    const float mg = sa[cl], mgx = sb[cl], k = ga * rs;
    for (long long p = p0 + rl; p < p1; p += 32) {
        const float xh = (x[p * C + c] - mu) * rs;
        float gy;
        if (sizeof(TDy) == 2) gy = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(dy)[p * C + c]);
        else gy = reinterpret_cast<const float *>(dy)[p * C + c];
        const float g = gy * (fmaf(xh, ga, be) > 0.f ? 1.f : slope);
        const float d = k * (g - mg - xh * mgx);
        if (sizeof(TDx) == 2) reinterpret_cast<__nv_bfloat16 *>(dx)[p * C + c] = __float2bfloat16_rn(d);
        else reinterpret_cast<float *>(dx)[p * C + c] = d;
    }
}

int bn_chunks(long long P, int C)
{
    const int groups = (C + 31) / 32;
    long long ch = (2 * kNumSMs + groups - 1) / groups;       // ~two CTAs of 1024 threads per SM
    if (ch > (P + 255) / 256) ch = (P + 255) / 256;           // at least 8 rows per row lane
    if (ch > kBnMaxChunks) ch = kBnMaxChunks;
    return (int)(ch < 1 ? 1 : ch);
}

// global average pool of x fp32 [N, HW, C] -> out [N, C], and its backward written as bf16 / fp32 [N, HW, C]
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const float *__restrict__ x, int N, int HW, int C, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int n = i / C, c = i % C;
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += x[((long long)n * HW + p) * C + c];
    out[i] = acc / (float)HW;
}
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float *__restrict__ g, int N, int HW, int C, T *__restrict__ dx)
{
    const long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C), n = (int)(i / ((long long)HW * C));
        const float v = g[n * C + c] / (float)HW;
        if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16 *>(dx)[i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float *>(dx)[i] = v;
    }
}

int ew_blocks(long long n, int threads)
{
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

SCDA_API int scda_conv_s2_weights(int Cout, int C, const void *w_krsc_bf16, void *wd_bf16, cudaStream_t stream)
{
    if (Cout <= 0 || C <= 0 || !w_krsc_bf16 || !wd_bf16) return 0;
    s2_weights_kernel<<<ew_blocks((long long)Cout * 36 * C, 256), 256, 0, stream>>>(
        (const __nv_bfloat16 *)w_krsc_bf16, (__nv_bfloat16 *)wd_bf16, Cout, C);
    return scda_launch_status();
}

SCDA_API int scda_conv_s2_wgrad_gather(int Cout, int C, const float *partials, int splits, float *dw, int accumulate,
                                       cudaStream_t stream)
{
    if (Cout <= 0 || C <= 0 || !partials || splits < 1 || !dw) return 0;
    s2_wgrad_gather_kernel<<<ew_blocks((long long)Cout * 9 * C, 256), 256, 0, stream>>>(partials, splits, dw, Cout, C,
                                                                                         accumulate);
    return scda_launch_status();
}

SCDA_API int scda_disc_l1_fwd(int N, int H, int W, const float *x, long long sn, long long sc, long long sh,
                              long long sw, const float *w_orsc, const float *bias, float slope, void *y, int y_f32,
                              cudaStream_t stream)
{
    if (N <= 0 || H <= 0 || W <= 0 || ((H | W) & 1) || !x || !w_orsc || !y) return 0;
    const long long total = (long long)N * (H / 2) * (W / 2);
    const int blocks = (int)((total + 127) / 128);
    if (y_f32) l1_fwd_kernel<float><<<blocks, 128, 0, stream>>>(x, sn, sc, sh, sw, N, H, W, w_orsc, bias, slope, (float *)y);
    else l1_fwd_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>(x, sn, sc, sh, sw, N, H, W, w_orsc, bias, slope,
                                                                 (__nv_bfloat16 *)y);
    return scda_launch_status();
}

SCDA_API size_t scda_disc_l1_workspace_bytes(int N, int H, int W)
{
    (void)N; (void)H; (void)W;
    return (size_t)kNumSMs * 2 * kL1Cout * 28 * sizeof(float);
}

SCDA_API int scda_disc_l1_bwd(int N, int H, int W, const float *x, long long sn, long long sc, long long sh,
                              long long sw, const float *w_orsc, const void *g, int g_f32, float *dw, float *db,
                              float *dx, int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    if (N <= 0 || H <= 0 || W <= 0 || ((H | W) & 1) || !x || !w_orsc || !g) return 0;
    if (dw) {
        if (!workspace || workspace_bytes < scda_disc_l1_workspace_bytes(N, H, W)) return 0;
        const long long total = (long long)N * (H / 2) * (W / 2);
        const int blocks = kNumSMs * 2;
        const long long per = ((total + blocks - 1) / blocks + kL1Tile - 1) / kL1Tile * kL1Tile;
        if (g_f32) l1_bwd_w_kernel<float><<<blocks, kL1Cout * 28, 0, stream>>>(x, sn, sc, sh, sw, N, H, W, (const float *)g,
                                                                               per, (float *)workspace);
        else l1_bwd_w_kernel<__nv_bfloat16><<<blocks, kL1Cout * 28, 0, stream>>>(x, sn, sc, sh, sw, N, H, W,
                                                                                 (const __nv_bfloat16 *)g, per,
                                                                                 (float *)workspace);
        l1_bwd_w_finish_kernel<<<(kL1Cout * 28 + 127) / 128, 128, 0, stream>>>((const float *)workspace, blocks, dw, db,
                                                                               accumulate);
    }
    if (dx) {
        const long long total = (long long)N * H * W;
        const int blocks = (int)((total + 127) / 128);
        if (g_f32) l1_bwd_x_kernel<float><<<blocks, 128, 0, stream>>>((const float *)g, w_orsc, N, H, W, dx);
        else l1_bwd_x_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>((const __nv_bfloat16 *)g, w_orsc, N, H, W, dx);
    }
    return scda_launch_status();
}

SCDA_API int scda_head_dot_fwd(long long P, int C, const void *x, int x_f32, const float *w, const float *bias,
                               float *out, cudaStream_t stream)
{
    if (P <= 0 || C <= 0 || !x || !w || !out) return 0;
    const int blocks = (int)((P * 32 + 255) / 256);
    if (x_f32) head_dot_fwd_kernel<float><<<blocks, 256, 0, stream>>>((const float *)x, w, bias, P, C, out);
    else head_dot_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16 *)x, w, bias, P, C, out);
    return scda_launch_status();
}

SCDA_API int scda_head_dot_bwd(long long P, int C, const void *x, int x_f32, const float *w, const float *g,
                               float slope, void *dx, float *dw, float *db, cudaStream_t stream)
{
    // dw / db are ACCUMULATED into (red.add): the caller zeroes them (the optimiser's flat gradient buffer is)
    if (P <= 0 || C <= 0 || C > 256 || 256 % C || !x || !w || !g) return 0;
    const int blocks = (int)((P + 63) / 64 < kNumSMs ? (P + 63) / 64 : kNumSMs);
    if (x_f32) head_dot_bwd_kernel<float><<<blocks, 256, 0, stream>>>((const float *)x, w, g, P, C, slope, (float *)dx, dw, db);
    else head_dot_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16 *)x, w, g, P, C, slope,
                                                                       (__nv_bfloat16 *)dx, dw, db);
    return scda_launch_status();
}

SCDA_API size_t scda_bn_workspace_bytes(long long P, int C)
{
    if (P <= 0 || C <= 0) return 0;
    return sizeof(float) * 2 * (size_t)bn_chunks(P, C) * (size_t)C;
}

SCDA_API int scda_bn_lrelu_fwd(long long P, int C, const float *x, const float *gamma, const float *beta, float eps,
                               float slope, float momentum, float *running_mean, float *running_var, float *mean,
                               float *rstd, void *y, int y_f32, void *workspace, size_t workspace_bytes,
                               cudaStream_t stream)
{
    if (P <= 0 || C <= 0 || !x || !gamma || !beta || !mean || !rstd || !y || !workspace) return 0;
    if (workspace_bytes < scda_bn_workspace_bytes(P, C)) return 0;
    const int chunks = bn_chunks(P, C);
    const dim3 grid((C + 31) / 32, chunks);
    float *partial = (float *)workspace;
    bn_fwd_partial_kernel<<<grid, 1024, 0, stream>>>(x, P, C, partial);
    if (y_f32)
        bn_lrelu_fwd_apply_kernel<float><<<grid, 1024, 0, stream>>>(x, P, C, gamma, beta, eps, slope, momentum,
                                                                    running_mean, running_var, mean, rstd, (float *)y,
                                                                    partial, chunks);
    else
        bn_lrelu_fwd_apply_kernel<__nv_bfloat16><<<grid, 1024, 0, stream>>>(x, P, C, gamma, beta, eps, slope, momentum,
                                                                            running_mean, running_var, mean, rstd,
                                                                            (__nv_bfloat16 *)y, partial, chunks);
    return scda_launch_status();
}

SCDA_API int scda_bn_lrelu_bwd(long long P, int C, const float *x, const void *dy, int dy_f32, const float *gamma,
                               const float *beta, const float *mean, const float *rstd, float slope, void *dx,
                               int dx_f32, float *dgamma, float *dbeta, int accumulate, void *workspace,
                               size_t workspace_bytes, cudaStream_t stream)
{
    if (P <= 0 || C <= 0 || !x || !dy || !gamma || !beta || !mean || !rstd || !dgamma || !dbeta || !workspace) return 0;
    if (workspace_bytes < scda_bn_workspace_bytes(P, C)) return 0;
    const int chunks = bn_chunks(P, C);
    const dim3 grid((C + 31) / 32, chunks);
    float *partial = (float *)workspace;
    if (dy_f32)
        bn_bwd_partial_kernel<float><<<grid, 1024, 0, stream>>>(x, (const float *)dy, P, C, gamma, beta, mean, rstd,
                                                                slope, partial);
    else
        bn_bwd_partial_kernel<__nv_bfloat16><<<grid, 1024, 0, stream>>>(x, (const __nv_bfloat16 *)dy, P, C, gamma, beta,
                                                                        mean, rstd, slope, partial);
#define SCDA_BN_BWD(TDY, TDX)                                                                                         \
    bn_lrelu_bwd_apply_kernel<TDY, TDX><<<grid, 1024, 0, stream>>>(x, (const TDY *)dy, P, C, gamma, beta, mean, rstd, \
                                                                   slope, (TDX *)dx, dgamma, dbeta, accumulate,     \
                                                                   partial, chunks)
    if (dy_f32 && dx_f32) SCDA_BN_BWD(float, float);
    else if (dy_f32) SCDA_BN_BWD(float, __nv_bfloat16);
    else if (dx_f32) SCDA_BN_BWD(__nv_bfloat16, float);
    else SCDA_BN_BWD(__nv_bfloat16, __nv_bfloat16);
#undef SCDA_BN_BWD
    return scda_launch_status();
}

SCDA_API int scda_avgpool_fwd(int N, int HW, int C, const float *x, float *out, cudaStream_t stream)
{
    if (N <= 0 || HW <= 0 || C <= 0 || !x || !out) return 0;
    avgpool_fwd_kernel<<<(N * C + 255) / 256, 256, 0, stream>>>(x, N, HW, C, out);
    return scda_launch_status();
}

SCDA_API int scda_avgpool_bwd(int N, int HW, int C, const float *g, void *dx, int dx_f32, cudaStream_t stream)
{
    if (N <= 0 || HW <= 0 || C <= 0 || !g || !dx) return 0;
    const int blocks = ew_blocks((long long)N * HW * C, 256);
    if (dx_f32) avgpool_bwd_kernel<float><<<blocks, 256, 0, stream>>>(g, N, HW, C, (float *)dx);
    else avgpool_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(g, N, HW, C, (__nv_bfloat16 *)dx);
    return scda_launch_status();
}
