// First convolution of the backbone (conv1_1: 3 -> 64 channels, 3x3, padding 1) straight from the fp32 NCHW
// image: the im2col lives ONLY in shared memory.
//
// Replaces, for models/faster_rcnn/vgg_adver_expansion_cluster.py:101-114 (features[0] + ReLU) of the reference,
// the two-pass form this build used before: nchw_to_nhwc (fp32 NCHW -> bf16 NHWC zero-padded 3 -> 64 channels:
// a 67 MB write per 512 x 1024 image) followed by the 64-channel halo convolution, which spent 21 of every 22
// MMAs on the zero padding (47 us for 1.8 GFLOP of real work).  Here a CTA builds the [128 pixels x 32] bf16
// patch matrix of a pixel tile in shared memory (27 columns k = (kh * 3 + kw) * 3 + ci, 5 zero columns;
// the 128B-swizzled K-major layout the UMMA descriptors of tc_ptx.cuh describe, written with ordinary stores and
// published to the async proxy with fence.proxy.async), multiplies it with the resident [64 x 32] weight tile by
// TWO tcgen05 MMAs (M128 N64 K16) and writes bias + ReLU + bf16 NHWC from TMEM.  HBM traffic = the image once
// (L1 / L2 serve the 9-fold neighbour reuse) + the output once: the kernel is bound by the 67 MB output write.
//
// Warp roles (320 threads, two CTAs per SM): warp 0 MMA issuer, warp 1 TMEM allocator, warps 2-5 patch
// producers (one pixel per thread), warps 6-9 epilogue (TMEM lane quarter = warp & 3).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kStages = 3;
constexpr int kTilePx = 128;
constexpr int kCout = 64;
constexpr int kABytes = kTilePx * 128;          // 128 rows of 128 B (K = 32 uses the first 64 B of a row)
constexpr int kBBytes = kCout * 128;
constexpr int kBarOffset = kStages * kABytes + kBBytes;
constexpr int kNumBars = 2 * kStages + 4;
constexpr int kSmemTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
constexpr int kThreads = 320;

struct FirstParams {
    int NB, H, W, Cin, K;          // K = 9 * Cin <= 32
    const float *x;                // [NB, Cin, H, W] fp32
    const __nv_bfloat16 *w;        // [64][3][3][Cin] bf16
    const float *bias;             // [64]
    __nv_bfloat16 *y;              // [NB, H, W, 64] bf16
    long long pixels;              // NB * H * W
    int tiles;
    int relu;
};

// byte offset of 16-byte chunk c of row r in a 128B-swizzled K-major tile (1024 B aligned base)
__device__ __forceinline__ uint32_t swz(int r, int c)
{
    return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(kThreads, 2) conv_first_kernel(const FirstParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_base = base + kStages * kABytes;
    const uint32_t bar_afull = base + kBarOffset;
    const uint32_t bar_aempty = bar_afull + kStages * 8;
    const uint32_t bar_tfull = bar_aempty + kStages * 8;
    const uint32_t bar_tempty = bar_tfull + 2 * 8;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + kBarOffset + kNumBars * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_afull + s * 8, 4);          // one arrival per producer warp
            mbar_init(bar_aempty + s * 8, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + a * 8, 1);
            mbar_init(bar_tempty + a * 8, 4);         // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32((const void *)tmem_slot)), "r"(2u * kCout) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the weight tile [64 rows = output channels][32 k] (rows of 128 B, swizzled), zero beyond k = K
    for (int i = threadIdx.x; i < kCout * 4; i += kThreads) {
        const int co = i >> 2, c = i & 3;
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            // patch column k = tap * 3 + ci whatever Cin is (the producers unroll three channels per tap)
            const int k0 = c * 8 + h * 2, k1 = k0 + 1;
            const int t0 = k0 / 3, c0 = k0 - t0 * 3, t1 = k1 / 3, c1 = k1 - t1 * 3;
            const float lo = (t0 < 9 && c0 < p.Cin) ? __bfloat162float(p.w[co * p.K + t0 * p.Cin + c0]) : 0.f;
            const float hi = (t1 < 9 && c1 < p.Cin) ? __bfloat162float(p.w[co * p.K + t1 * p.Cin + c1]) : 0.f;
            const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
            pk[h] = *reinterpret_cast<const uint32_t *>(&v);
        }
        *reinterpret_cast<uint4 *>(gen + kStages * kABytes + swz(co, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- MMA issuer: two K = 16 steps per pixel tile
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(128, kCout, 0, 0);
            const uint64_t proto = make_kmajor_desc(0);
            const uint32_t hi = (uint32_t)(proto >> 32);
            const uint32_t a_lo0 = (uint32_t)proto + (base >> 4);
            const uint32_t b_lo = (uint32_t)proto + (b_base >> 4);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % kStages, acc = it & 1;
                mbar_wait(bar_tempty + acc * 8, ((it >> 1) & 1) ^ 1);
                mbar_wait(bar_afull + s * 8, (it / kStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_lo = a_lo0 + s * (kABytes >> 4);
                umma_bf16_lohi(tmem_base + acc * kCout, a_lo, hi, b_lo, hi, idesc, 0u);
                umma_bf16_lohi(tmem_base + acc * kCout, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
                umma_commit(bar_aempty + s * 8);
                umma_commit(bar_tfull + acc * 8);
            }
        }
    } else if (warp >= 2 && warp < 6) {
        // ---- patch producers: thread r owns pixel r of the tile
        const int r = (warp - 2) * 32 + lane;
        const long long plane = (long long)p.H * p.W;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % kStages;
            const long long px = (long long)tile * kTilePx + r;
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.f;
            if (px < p.pixels) {
                const int n = (int)(px / plane);
                const int rem = (int)(px - (long long)n * plane);
                const int h = rem / p.W, w = rem - h * p.W;
                const float *img = p.x + (long long)n * p.Cin * plane;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
                    const bool ok = hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
                    const long long o = (long long)hh * p.W + ww;
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci)
                        if (ci < p.Cin && ok) v[t * 3 + ci] = __ldg(img + ci * plane + o);
                }
            }
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                pk[k] = *reinterpret_cast<const uint32_t *>(&b);
            }
            mbar_wait(bar_aempty + s * 8, ((it / kStages) & 1) ^ 1);
            uint8_t *dst = gen + s * kABytes;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4 *>(dst + swz(r, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull + s * 8);
        }
    } else if (warp >= 6) {
        // ---- epilogue: bias + ReLU + bf16, one pixel (64 channels = 128 B) per lane
        EpiParams e;
        e.bias = p.bias; e.out = p.y; e.ldc = kCout; e.mask_src = nullptr; e.mul_src = nullptr;
        e.flags = 0; e.N = kCout; e.slope = 0.f;
        const int q = warp & 3;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            mbar_wait(bar_tfull + acc * 8, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long px = (long long)tile * kTilePx + q * 32 + lane;
            uint32_t v0[32], v1[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kCout, v0);
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kCout + 32, v1);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + acc * 8);
            if (px < p.pixels) {
                if (p.relu) {
                    epilogue_chunk<kFlagRelu | kFlagBias>(v0, e, px, 0, true);
                    epilogue_chunk<kFlagRelu | kFlagBias>(v1, e, px, 32, true);
                } else {
                    epilogue_chunk<kFlagBias>(v0, e, px, 0, true);
                    epilogue_chunk<kFlagBias>(v1, e, px, 32, true);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kCout)
                     : "memory");
    }
}

}  // namespace

// y[n,h,w,co] = relu?(bias[co] + sum_{kh,kw,ci} x[n,ci,h+kh-1,w+kw-1] * w[co][kh][kw][ci]), x fp32 NCHW (rounded to
// bf16 as it is staged), w bf16 [64][3][3][Cin] (Cin <= 3), y bf16 NHWC.  fp32 accumulation.
SCDA_API int scda_conv3x3_first_nchw(int NB, int H, int W, int Cin, int Cout, const float *x, const void *w_krsc,
                                     const float *bias, void *y, int relu, cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin > 3 || Cout != kCout || !x || !w_krsc || !bias || !y) return 0;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15) return 0;
    FirstParams p = {};
    p.NB = NB; p.H = H; p.W = W; p.Cin = Cin; p.K = 9 * Cin;
    p.x = x; p.w = (const __nv_bfloat16 *)w_krsc; p.bias = bias; p.y = (__nv_bfloat16 *)y;
    p.pixels = (long long)NB * H * W;
    p.tiles = ceil_div(p.pixels, kTilePx);
    p.relu = relu ? 1 : 0;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    const int max_ctas = 2 * num_sms();
    const int grid = p.tiles < max_ctas ? p.tiles : max_ctas;
    conv_first_kernel<<<grid, kThreads, kSmemTotal, stream>>>(p);
    return scda_launch_status();
}
