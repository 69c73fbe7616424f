// Epilogue of the tensor-core kernels: 32 fp32 accumulator columns of one output row (already in
// registers, one row per lane) -> bias, ReLU, ReLU-gradient mask, dropout scale -> bf16 / fp32
// store.  The flag set is a template parameter for the combinations the detector uses, so the
// per-element work is a handful of instructions: with run-time flags every element carried ~40
// predicated-off instructions and the epilogue warps, not the tensor pipe, set the pace of the
// short-reduction layers (conv1_x: 9.4 k clocks per tile against 2.3 k of MMA,
// profiles/r1_ncu_halo_g.txt).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tcptx {

enum : int {
    kFlagRelu = 1,        // y = max(y, 0)
    kFlagOutF32 = 2,      // write fp32 instead of bf16
    kFlagMaskPos = 4,     // y = (mask_src > 0) ? y : 0     (ReLU backward fused into dgrad)
    kFlagAccumulate = 8,  // y += previous contents (fp32 output only)
    kFlagMulSrc = 16,     // y *= mul_src            (dropout keep/scale tensor, bf16)
    kFlagBias = 32,       // internal: bias pointer present
    kFlagMaskF32 = 64,    // mask_src holds fp32 (the fp32-parity mode keeps activations in fp32)
    kFlagLeaky = 128,     // y = y > 0 ? y : slope * y            (LeakyReLU forward)
    kFlagMaskLeaky = 256, // with kFlagMaskPos: y = mask_src > 0 ? y : slope * y   (LeakyReLU backward)
};

struct EpiParams {
    const float *bias;               // [N] or null
    void *out;                       // [rows, ldc]
    long long ldc;
    const __nv_bfloat16 *mask_src;   // [rows, ldc]
    const __nv_bfloat16 *mul_src;    // [rows, ldc]
    int flags;                       // run-time flag set (| kFlagBias)
    int N;
    float slope;                     // LeakyReLU negative slope (kFlagLeaky / kFlagMaskLeaky)
};

// kSpec >= 0: the flag set is the compile-time constant kSpec; kSpec < 0: e.flags at run time.
// v: accumulator columns [col0, col0 + 32) of output row `out_row`.
template <int kSpec>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const EpiParams &e, long long out_row,
                                               int col0, bool vec_ok)
{
    const int flags = kSpec >= 0 ? kSpec : e.flags;
    const int ncol = min(32, e.N - col0);
    if (ncol <= 0) return;
    const long long o = out_row * e.ldc + col0;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (ncol == 32 && vec_ok) {
        if (flags & kFlagBias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4 *>(e.bias + col0 + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
        }
        if (flags & kFlagRelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (flags & kFlagLeaky) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * e.slope;
        }
#define SCDA_MASKED(x) ((flags & kFlagMaskLeaky) ? (x) * e.slope : 0.f)     /* value where the mask is <= 0 */
        if ((flags & kFlagMaskPos) && (flags & kFlagMaskF32)) {
            const float *mf = reinterpret_cast<const float *>(e.mask_src) + o;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(mf + j));
                if (!(q.x > 0.f)) f[j] = SCDA_MASKED(f[j]);
                if (!(q.y > 0.f)) f[j + 1] = SCDA_MASKED(f[j + 1]);
                if (!(q.z > 0.f)) f[j + 2] = SCDA_MASKED(f[j + 2]);
                if (!(q.w > 0.f)) f[j + 3] = SCDA_MASKED(f[j + 3]);
            }
        } else if (flags & kFlagMaskPos) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(e.mask_src + o + j));
                const __nv_bfloat16 *qb = reinterpret_cast<const __nv_bfloat16 *>(&q);
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (!(__bfloat162float(qb[t]) > 0.f)) f[j + t] = SCDA_MASKED(f[j + t]);
            }
        }
        if (flags & kFlagMulSrc) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(e.mul_src + o + j));
                const __nv_bfloat16 *qb = reinterpret_cast<const __nv_bfloat16 *>(&q);
#pragma unroll
                for (int t = 0; t < 8; ++t) f[j + t] *= __bfloat162float(qb[t]);
            }
        }
        if (flags & kFlagOutF32) {
            float *dst = reinterpret_cast<float *>(e.out) + o;
            if (flags & kFlagAccumulate) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 old = *reinterpret_cast<const float4 *>(dst + j);
                    *reinterpret_cast<float4 *>(dst + j) =
                        make_float4(old.x + f[j], old.y + f[j + 1], old.z + f[j + 2], old.w + f[j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
        } else {
            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(e.out) + o;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[j], f[j + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                pk.x = *reinterpret_cast<uint32_t *>(&b0);
                pk.y = *reinterpret_cast<uint32_t *>(&b1);
                pk.z = *reinterpret_cast<uint32_t *>(&b2);
                pk.w = *reinterpret_cast<uint32_t *>(&b3);
                *reinterpret_cast<uint4 *>(dst + j) = pk;
            }
        }
        return;
    }
    // ragged tail / unaligned leading dimension: element by element
#pragma unroll 1
    for (int j = 0; j < ncol; ++j) {
        float x = __uint_as_float(v[0]);
        // (register arrays cannot be indexed dynamically without spilling: select by a static scan)
#pragma unroll
        for (int t = 1; t < 32; ++t) x = (t == j) ? __uint_as_float(v[t]) : x;
        if (flags & kFlagBias) x += __ldg(e.bias + col0 + j);
        if (flags & kFlagRelu) x = fmaxf(x, 0.f);
        if (flags & kFlagLeaky) x = x > 0.f ? x : x * e.slope;
        if (flags & kFlagMaskPos) {
            const float m = (flags & kFlagMaskF32) ? reinterpret_cast<const float *>(e.mask_src)[o + j]
                                                   : __bfloat162float(e.mask_src[o + j]);
            if (!(m > 0.f)) x = SCDA_MASKED(x);
        }
        if (flags & kFlagMulSrc) x *= __bfloat162float(e.mul_src[o + j]);
        if (flags & kFlagOutF32) {
            float *dst = reinterpret_cast<float *>(e.out) + o + j;
            *dst = (flags & kFlagAccumulate) ? *dst + x : x;
        } else {
            reinterpret_cast<__nv_bfloat16 *>(e.out)[o + j] = __float2bfloat16_rn(x);
        }
    }
}

#undef SCDA_MASKED

}  // namespace tcptx
