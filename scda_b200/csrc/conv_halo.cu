// 3x3 convolution (stride 1, zero padding 1, bf16 NHWC, fp32 accumulate in TMEM) with the im2col
// folded into the SHARED-MEMORY STAGING: the input halo of a pixel tile is loaded ONCE per
// 64-channel block and all nine taps are MMA operands read out of that one copy.
//
// Replaces cuDNN behind the reference's nn.Conv2d(k=3, p=1) layers on the hot path (13 VGG convs:
// models/faster_rcnn/vgg_adver_expansion_cluster.py:101-114; RPN 3x3: models/head.py:13-18; the
// decoder's 3x3 convs: models/faster_rcnn/common_net.py:59-80,279-293) and their data gradients.
//
// Why: the per-tap form (gemm_tc.cu, one TMA box per tap) pulls every input pixel through the
// L2->SM fabric nine times and every weight once per 128 pixels; ncu shows it bound there
// (profiles/r1_ncu_conv_f_full.txt: 7-8 TB/s L2->SM, tensor pipe 15-47 %).  Here a CTA owns a
// (8*kSub) x 16 pixel region:
//   A  one 4-D TMA box {64 ch, 8*kSub+2, 18, 1} at {c0, w0-1, h0-1, n}: the halo, rows of 128 B
//      (one pixel x 64 channels), 128B-swizzled, zero-filled outside the image (= the padding).
//      Tap (r,s) of sub-tile j is the SAME bytes seen through a K-major UMMA descriptor whose
//      start address is shifted by (r*HW + s + 8j) rows and whose 8-row-group stride (SBO) is one
//      halo row (HW*128 B): 8 consecutive pixels of an image row are 8 consecutive 128 B rows.
//      The 128B swizzle is a function of the absolute shared-memory address on both the TMA and
//      the MMA side, so a start that is not 1024 B aligned un-swizzles correctly with the
//      descriptor's base_offset field left 0 (measured: scripts/probe_desc.cu,
//      profiles/r1_probe_umma_desc_row_shift.txt — every row shift 0..19 x SBO 1024/1280/2048/2304).
//   B  one weight tile per (tap, channel block), used by the kSub sub-tiles back to back.
// Bytes staged per 128 px x 128 ch x 64 k x 9 taps of work: 295 KB (per-tap form) ->
// 170 KB (kSub=1) -> 94 KB (kSub=2).
//
// Warp roles (128 + 128*kSub threads, one CTA per SM, persistent over tiles):
//   warp 0: B producer   warp 3: A (halo) producer   warp 1: MMA issuer   warp 2: TMEM alloc
//   warps 4..: epilogue, 4 warps per sub-tile (TMEM lane quarters), double-buffered accumulators
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "conv_halo.h"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int kTileH = 16;          // sub-tile = 8 wide x 16 tall = 128 pixels = the MMA's M
constexpr int kSubW = 8;
constexpr int kHaloH = kTileH + 2;

struct HaloParams {
    int H, W, Cred, N;           // Cred: channels of the A tensor (the reduction); N: output channels
    int cblocks;                 // Cred / 64
    int tiles_w, tiles_h, m_tiles, n_tiles;
    const float *bias;
    void *out;                   // [NB*H*W, ldc]
    long long ldc;
    const __nv_bfloat16 *mask_src;
    int flags;
    // stride-2 form (kS2): the convolution runs on the (H, W) = OUTPUT grid over 4 * C virtual input channels
    // (row phase, column phase, channel); c2 = 2 * C = the channels of one row phase
    int c2;
    int tap_mask;                // taps (bit r * 3 + s) that carry non-zero weights
    float slope;                 // LeakyReLU slope of the epilogue
};

// Halo stages: with 64 output channels a tile retires its nine taps in ~2.3 k clocks, about the
// latency of the next tile's 41 KB halo box (324 strided 128 B rows, streamed from HBM at the
// conv1_x resolution), so two stages leave the MMA lane waiting; the narrow B ring of N = 64
// leaves room for a third.
template <int kBlockN>
struct HaloAStages {
    static constexpr int value = kBlockN == 64 ? 3 : 2;
};

template <int kBlockN, int kSub, int kBStages, int kCluster = 1>
struct HaloSmem {
    static constexpr int kAStages = kCluster == 2 ? 3 : HaloAStages<kBlockN>::value;
    static constexpr int kHaloW = kSubW * kSub + 2;
    static constexpr int kABox = kHaloW * kHaloH * 128;                 // bytes one halo box delivers
    static constexpr int kABytes = (kABox + 1023) / 1024 * 1024;
    static constexpr int kBBytes = kBlockN / kCluster * 128;            // a CTA pair holds half of the weight tile each
    static constexpr int kBOffset = kAStages * kABytes;
    static constexpr int kBarOffset = kBOffset + kBStages * kBBytes;
    static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
};

struct HaloTile {
    int n0, img, h0, w0;
    bool valid;                  // (pair) false for the phantom second tile of an odd tile count
};

// tile = tile GROUP index: kCluster consecutive pixel tiles (one per CTA of the cluster) x one N tile
template <int kBlockN, int kSub, int kCluster = 1>
__device__ __forceinline__ HaloTile halo_tile(const HaloParams &p, int tile, int rank = 0)
{
    HaloTile t;
    const int m_group = tile / p.n_tiles;
    const int m_tile = m_group * kCluster + rank;
    t.valid = m_tile < p.m_tiles;
    t.n0 = (tile - m_group * p.n_tiles) * kBlockN;
    const int per_img = p.tiles_h * p.tiles_w;
    t.img = m_tile / per_img;
    const int r = m_tile - t.img * per_img;
    t.h0 = (r / p.tiles_w) * kTileH;
    t.w0 = (r % p.tiles_w) * (kSubW * kSub);
    return t;
}

template <int kBlockN, int kSub, int kSpec, bool kS2Dgrad = false, int kCluster = 1>
__device__ __forceinline__ void halo_epilogue(const HaloParams &p, const EpiParams &e, uint32_t tmem_base,
                                              uint32_t bar_tfull, uint32_t bar_tempty, int ew, int lane,
                                              int rank = 0)
{
    constexpr uint32_t kAccCols = kSub * kBlockN;
    constexpr bool kPair = kCluster == 2;
    const int q = ew & 3;                 // TMEM lane quarter this warp may read
    const int sub = ew >> 2;
    const int total_tiles = ceil_div(p.m_tiles, kCluster) * p.n_tiles;
    const int first_tile = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;
    // (pair) the MMA lane that waits for the drained accumulators lives in the leader CTA
    const uint32_t tempty_dst = kPair ? map_to_cta(bar_tempty, 0) : bar_tempty;
    const bool aligned = ((reinterpret_cast<uintptr_t>(e.out) | reinterpret_cast<uintptr_t>(e.bias) |
                           reinterpret_cast<uintptr_t>(e.mask_src)) & 15) == 0;
    const bool vec_ok = (e.ldc % 8 == 0) && aligned;
    uint32_t ti = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++ti) {
        const HaloTile t = halo_tile<kBlockN, kSub, kCluster>(p, tile, rank);
        const uint32_t acc = ti & 1;
        mbar_wait(bar_tfull + acc * 8, (ti >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = q * 32 + lane;
        const int h = t.h0 + (row >> 3), w = t.w0 + (row & 7) + kSubW * sub;
        const bool row_ok = h < p.H && w < p.W && t.valid;
        long long out_row = ((long long)t.img * p.H + h) * p.W + w;
        int col_base = t.n0;
        if (kS2Dgrad) {
            // data gradient of the stride-2 convolution: N tile = virtual channels (py, px, c) of the input
            // seen as [NB, 2H, W, 2C]; row phase py = n0 / 2C selects the image row 2h + py
            const int py = t.n0 / p.c2;
            col_base = t.n0 - py * p.c2;
            out_row = (((long long)t.img * p.H + h) * 2 + py) * p.W + w;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < kBlockN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols + sub * kBlockN + c0, v);
            if (c0 + 32 >= kBlockN) {
                // this warp's share of the accumulator is in registers: hand the buffer back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (kPair) mbar_arrive_cluster(tempty_dst + acc * 8);
                    else mbar_arrive(bar_tempty + acc * 8);
                }
            }
            if (row_ok) epilogue_chunk<kSpec>(v, e, out_row, col_base + c0, vec_ok);
        }
    }
}

// kS2: the stride-2 form (discriminators: models/faster_rcnn/common_net.py:205-261 of the reference,
// nn.Conv2d(k=3, s=2, p=1)).  out(i, j) = sum_{r,s} x(2i + r - 1, 2j + s - 1) W[r, s] is a 2 x 2-tap stride-1
// convolution over the input's four phase images: the A tensor map views x as [NB, Hin, Win/2, 2C] and walks
// the row dimension with element stride 2, so the halo box of virtual channel block (py, 64 of 2C) is the
// phase image's tile, zero-filled outside; the weights are laid out [Cout][3x3 taps][4C] with the taps of the
// two offsets {-1, 0} (tap_mask) — taps that are structurally zero are neither loaded nor multiplied.
//
// kCluster == 2 is the CTA-PAIR form (tcgen05 cta_group::2, the protocol of tc_gemm_kernel): the two CTAs of a
// cluster own two consecutive pixel tiles of the same N tile and execute ONE 256 x kBlockN MMA per k-step.  Each
// CTA stages its own halo and only HALF of every weight tile; the tensor core of each SM reads the other half out
// of the peer's shared memory.  Bytes staged per SM per 128 px x 256 ch x 64 k x 9 taps of work: 334 KB as two
// lone 128-wide tiles, 167 KB as one paired 256-wide tile — under the ~42 B/clk an SM can pull through the L2
// fabric with every SM loading, which is what held the 256 / 512-channel layers at 65-75 % tensor pipe.
// Both producers complete their bytes on the LEADER's full barriers; only the leader's MMA lane issues; its
// commits arrive on the empty / accumulator-full barriers of BOTH CTAs; the epilogue warps of both CTAs arrive
// on the leader's accumulator-empty barrier.
template <int kBlockN, int kSub, bool kBMn, int kBStages, bool kS2 = false, int kCluster = 1>
__global__ void __launch_bounds__(128 + 128 * kSub, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const HaloParams p)
{
    static_assert(kCluster == 1 || (kCluster == 2 && !kS2), "one CTA or a CTA pair (stride-1 form)");
    constexpr bool kPair = kCluster == 2;
    using L = HaloSmem<kBlockN, kSub, kBStages, kCluster>;
    constexpr int kAStages = L::kAStages;
    constexpr bool kBRes = kBStages == 9;        // nine slots = the resident-weights form
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_afull = base + L::kBarOffset;
    const uint32_t bar_aempty = bar_afull + kAStages * 8;
    const uint32_t bar_bfull = bar_aempty + kAStages * 8;
    const uint32_t bar_bempty = bar_bfull + kBStages * 8;
    const uint32_t bar_tfull = bar_bempty + kBStages * 8;
    const uint32_t bar_tempty = bar_tfull + 2 * 8;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + L::kBarOffset + L::kNumBars * 8);
    constexpr uint32_t kAccCols = kSub * kBlockN;        // TMEM columns of one accumulator set

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = kPair ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int total_tiles = ceil_div(p.m_tiles, kCluster) * p.n_tiles;         // tile groups
    const int first_tile = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(bar_afull + s * 8, 1);
            mbar_init(bar_aempty + s * 8, 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(bar_bfull + s * 8, 1);
            mbar_init(bar_bempty + s * 8, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + a * 8, 1);
            mbar_init(bar_tempty + a * 8, 4 * kSub * kCluster);     // one arrival per epilogue warp (of the pair)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32((const void *)tmem_slot)), "r"(2u * kAccCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32((const void *)tmem_slot)), "r"(2u * kAccCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kPair) cluster_sync_all();       // both CTAs' barriers exist before any remote traffic
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // the barriers the loads complete on: this CTA's, or (pair) the leader's
    const uint32_t afull_dst = kPair ? map_to_cta(bar_afull, 0) : bar_afull;
    const uint32_t bfull_dst = kPair ? map_to_cta(bar_bfull, 0) : bar_bfull;

    if (warp == 3) {
        // ---- A producer: one halo box per (tile, channel block)
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
                const HaloTile t = halo_tile<kBlockN, kSub, kCluster>(p, tile, rank);
                for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
                    const int s = it % kAStages;
                    mbar_wait(bar_aempty + s * 8, ((it / kAStages) & 1) ^ 1);
                    if (leader) mbar_expect_tx(bar_afull + s * 8, L::kABox * kCluster);
                    if (kPair) {
                        tma_load_4d_pair(base + s * L::kABytes, &map_a, afull_dst + s * 8, cb * 64, t.w0 - 1, t.h0 - 1,
                                         t.img);
                    } else if (kS2 && !kBMn) {
                        const int py = (cb * 64) / p.c2;
                        tma_load_4d(base + s * L::kABytes, &map_a, bar_afull + s * 8, cb * 64 - py * p.c2, t.w0 - 1,
                                    2 * (t.h0 - 1) + py, t.img);
                    } else {
                        tma_load_4d(base + s * L::kABytes, &map_a, bar_afull + s * 8, cb * 64, t.w0 - 1, t.h0 - 1,
                                    t.img);
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ---- B producer: one weight tile per (tile, channel block, tap)
        if (elect_one()) {
            uint32_t it = 0;
            // kBRes: one channel block, one N tile -> the nine weight tiles are loaded ONCE and stay
            // in their nine slots for every pixel tile this CTA walks
            constexpr int kHalfN = kBlockN / kCluster;           // weight rows / columns staged by this CTA
            for (int tile = first_tile; tile < (kBRes ? min(total_tiles, first_tile + 1) : total_tiles);
                 tile += tile_step) {
                const HaloTile t = halo_tile<kBlockN, kSub, kCluster>(p, tile, rank);
                const int nb = t.n0 + rank * kHalfN;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    for (int tap = 0; tap < 9; ++tap) {
                        if (kS2 && !((p.tap_mask >> tap) & 1)) continue;
                        const int s = it % kBStages;
                        ++it;
                        if (!kBRes) mbar_wait(bar_bempty + s * 8, (((it - 1) / kBStages) & 1) ^ 1);
                        if (leader) mbar_expect_tx(bar_bfull + s * 8, L::kBBytes * kCluster);
                        const uint32_t b_dst = base + L::kBOffset + s * L::kBBytes;
                        const uint32_t fb = bfull_dst + s * 8;
                        if (kBMn) {
                            // data gradient: B straight from the forward weights W[co][tap][ci] seen
                            // as [co rows][9 * N cols]; reduction index = co (rows), output channel
                            // = ci (contiguous), tap mirrored
#pragma unroll
                            for (int c = 0; c < kHalfN / 64; ++c) {
                                if (kPair) tma_load_2d_pair(b_dst + c * (64 * 128), &map_b, fb,
                                                            (8 - tap) * p.N + nb + c * 64, cb * 64);
                                else tma_load_2d(b_dst + c * (64 * 128), &map_b, fb,
                                                 (8 - tap) * p.N + nb + c * 64, cb * 64);
                            }
                        } else {
                            if (kPair) tma_load_2d_pair(b_dst, &map_b, fb, tap * p.Cred + cb * 64, nb);
                            else tma_load_2d(b_dst, &map_b, fb, tap * p.Cred + cb * 64, nb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        if (elect_one() && leader) {
            constexpr uint32_t idesc = make_idesc(128 * kCluster, kBlockN, 0, kBMn ? 1 : 0);
            constexpr uint32_t kSbo = L::kHaloW * 128;
            // descriptors as (lo, hi) words: hi is constant, lo = start >> 4 (+ LBO field) and only
            // ever gets a compile-time offset added (tap / sub-tile row shift, K advance)
            const uint64_t a_proto = make_kmajor_desc_ex(0, kSbo, 0);
            const uint64_t b_proto = kBMn ? make_mnmajor_desc(0, 64 * 128) : make_kmajor_desc(0);
            const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
            const uint32_t a_lo0 = (uint32_t)a_proto + (base >> 4);
            const uint32_t b_lo0 = (uint32_t)b_proto + ((base + L::kBOffset) >> 4);
            constexpr uint32_t kBk = kBMn ? (16 * 128) >> 4 : 32 >> 4;      // K advance of B per MMA
            uint32_t ait = 0, bit = 0, ti = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++ti) {
                const uint32_t acc = ti & 1;
                mbar_wait(bar_tempty + acc * 8, ((ti >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * kAccCols;
                for (int cb = 0; cb < p.cblocks; ++cb, ++ait) {
                    const uint32_t as = ait % kAStages;
                    mbar_wait(bar_afull + as * 8, (ait / kAStages) & 1);
                    const uint32_t a_lo = a_lo0 + as * (L::kABytes >> 4);
                    uint32_t done_taps = 0;          // (kS2) MMAs already issued for this channel block
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (kS2 && !((p.tap_mask >> tap) & 1)) continue;
                        const uint32_t bs = kBRes ? (uint32_t)tap : bit % kBStages;
                        // (resident: the slot's first and only phase; later waits return at once)
                        mbar_wait(bar_bfull + bs * 8, kBRes ? 0u : (bit / kBStages) & 1);
                        ++bit;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_lo = b_lo0 + bs * (L::kBBytes >> 4);
                        const uint32_t first = kS2 ? (uint32_t)cb | done_taps : (uint32_t)(cb | tap);
                        done_taps = 1;
#pragma unroll
                        for (int sub = 0; sub < kSub; ++sub) {
                            // rows (tap/3) * HW + tap%3 + 8*sub of 128 B each = 8 units of 16 B per row
                            const uint32_t a_off = (uint32_t)((tap / 3) * L::kHaloW + (tap % 3) + kSubW * sub) * 8u;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (kPair) umma_bf16_pair_lohi(tmem_d + sub * kBlockN, a_lo + a_off + k * 2, a_hi,
                                                               b_lo + k * kBk, b_hi, idesc, first | (uint32_t)k);
                                else umma_bf16_lohi(tmem_d + sub * kBlockN, a_lo + a_off + k * 2, a_hi, b_lo + k * kBk,
                                                    b_hi, idesc, first | (uint32_t)k);
                            }
                        }
                        if (!kBRes) {
                            if (kPair) umma_commit_pair(bar_bempty + bs * 8);
                            else umma_commit(bar_bempty + bs * 8);
                        }
                    }
                    if (kPair) umma_commit_pair(bar_aempty + as * 8);
                    else umma_commit(bar_aempty + as * 8);
                }
                if (kPair) umma_commit_pair(bar_tfull + acc * 8);
                else umma_commit(bar_tfull + acc * 8);
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue (flag set resolved at compile time for the combinations the detector uses)
        EpiParams e;
        e.bias = p.bias; e.out = p.out; e.ldc = p.ldc; e.mask_src = p.mask_src; e.mul_src = nullptr;
        e.flags = p.flags | (p.bias ? kFlagBias : 0);
        e.N = p.N;
        e.slope = p.slope;
#define SCDA_HEPI(SPEC)                                                                                          \
    halo_epilogue<kBlockN, kSub, SPEC, kS2 && kBMn, kCluster>(p, e, tmem_base, bar_tfull, bar_tempty, warp - 4, lane, \
                                                              rank)
        if (kS2) {
            // discriminator layers: conv + bias + LeakyReLU; its data gradient through the LeakyReLU below
            if (kBMn) { e.N = p.c2; }
            if (e.flags == (kFlagLeaky | kFlagBias)) SCDA_HEPI(kFlagLeaky | kFlagBias);
            else if (e.flags == (kFlagMaskPos | kFlagMaskLeaky)) SCDA_HEPI(kFlagMaskPos | kFlagMaskLeaky);
            else if (e.flags == kFlagOutF32) SCDA_HEPI(kFlagOutF32);
            else if (e.flags == 0) SCDA_HEPI(0);
            else SCDA_HEPI(-1);
        } else if (e.flags == (kFlagRelu | kFlagBias)) SCDA_HEPI(kFlagRelu | kFlagBias);
        else if (e.flags == kFlagMaskPos) SCDA_HEPI(kFlagMaskPos);
        else if (e.flags == 0) SCDA_HEPI(0);
        // fp32-parity mode (x3_ops.cu): fp32 outputs, fp32 ReLU masks
        else if (e.flags == (kFlagRelu | kFlagBias | kFlagOutF32)) SCDA_HEPI(kFlagRelu | kFlagBias | kFlagOutF32);
        else if (e.flags == (kFlagMaskPos | kFlagMaskF32 | kFlagOutF32)) SCDA_HEPI(kFlagMaskPos | kFlagMaskF32 | kFlagOutF32);
        else if (e.flags == kFlagOutF32) SCDA_HEPI(kFlagOutF32);
        else if (e.flags == (kFlagOutF32 | kFlagBias)) SCDA_HEPI(kFlagOutF32 | kFlagBias);      // (decoder conv -> IN)
        else SCDA_HEPI(-1);
#undef SCDA_HEPI
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kPair) cluster_sync_all();       // no CTA leaves while its peer may still signal or read it
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols)
                         : "memory");
    }
}

template <int kBlockN, int kSub, bool kBMn, int kBStages, bool kS2 = false, int kCluster = 1>
int launch_halo(const CUtensorMap &ma, const CUtensorMap &mb, const HaloParams &p, cudaStream_t stream)
{
    using L = HaloSmem<kBlockN, kSub, kBStages, kCluster>;
    static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
    static_assert(2 * kSub * kBlockN <= 512, "TMEM columns");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<kBlockN, kSub, kBMn, kBStages, kS2, kCluster>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    const long long groups = (long long)ceil_div(p.m_tiles, kCluster) * p.n_tiles;
    const int max_clusters = num_sms() / kCluster;
    const int clusters = (int)(groups < max_clusters ? groups : max_clusters);
    if (kCluster == 1) {
        conv_halo_kernel<kBlockN, kSub, kBMn, kBStages, kS2, 1><<<clusters, 128 + 128 * kSub, L::kTotal, stream>>>(ma, mb, p);
        return scda_launch_status();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCluster);
    cfg.blockDim = dim3(128 + 128 * kSub);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<kBlockN, kSub, kBMn, kBStages, kS2, kCluster>, ma, mb, p);
    if (e != cudaSuccess) return -(int)e;
    return scda_launch_status();
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

}  // namespace

// Tile plan of the halo kernel for a layer: bn = N tile (64 | 128), sub = sub-tiles per CTA
// (1 | 2); returns false where the per-tap kernel of gemm_tc.cu should be used instead.
// scda_conv3x3_set_plan (or SCDA_CONV_HALO / SCDA_HALO_BN / SCDA_HALO_SUB at load) overrides it.
static int g_enabled = -1, g_force_bn = 0, g_force_sub = 0, g_resident = 1, g_sub128 = 1, g_pair = -1;

static void plan_init()
{
    if (g_enabled >= 0) return;
    g_enabled = env_int("SCDA_CONV_HALO", 1) ? 1 : 0;
    g_force_bn = env_int("SCDA_HALO_BN", 0);
    g_force_sub = env_int("SCDA_HALO_SUB", 0);
    g_resident = env_int("SCDA_HALO_RESIDENT", 1);
    g_sub128 = env_int("SCDA_HALO_SUB128", 1) == 2 ? 2 : 1;
    g_pair = env_int("SCDA_HALO_PAIR", -1);
}

// CTA-pair form of the halo kernel: -1 = the measured plan (scda_conv_halo_plan), 0 = never, 1 = wherever legal
SCDA_API int scda_conv3x3_set_pair(int mode)
{
    plan_init();
    if (mode < -1 || mode > 1) return 0;
    g_pair = mode;
    return 1;
}

SCDA_API int scda_conv3x3_set_plan(int halo, int block_n, int sub_tiles)
{
    plan_init();
    if (halo >= 0) g_enabled = halo ? 1 : 0;
    if (block_n >= 0) {
        if (block_n != 0 && block_n != 64 && block_n != 128 && block_n != 256) return 0;
        g_force_bn = block_n;
    }
    if (sub_tiles >= 0) {
        if (sub_tiles > 2) return 0;
        g_force_sub = sub_tiles;
    }
    return 1;
}

// where the pair form measured faster than the lone-CTA plan on B200 (scripts/halobench.py)
static bool pair_pays(int Cred, int Nout, int pb, long long m_pairs)
{
    (void)Cred; (void)Nout; (void)pb; (void)m_pairs;
    return false;
}

bool scda_conv_halo_plan(int NB, int H, int W, int Cred, int Nout, bool dgrad, int *bn, int *sub, int *cluster)
{
    plan_init();
    *cluster = 1;
    // forward: the reduction walks whole 64-channel blocks of a K-major weight row (tap, ci), so Cred
    // must be a multiple of 64; 32 output channels ride in a 64-wide N tile whose upper weight rows are
    // TMA out-of-bounds zeros.  Data gradient: 32 reduction channels (Cout_fwd = 32) are a 64-channel
    // block whose upper half is zero-filled on BOTH operands (activation box and MN-major weight rows).
    if (!g_enabled) return false;
    if (dgrad ? (Cred % 32 || Nout % 64) : (Cred % 64 || Nout % 32)) return false;
    // measured on B200 (profiles/r1_halobench_d_elect.jsonl): with the MMA lane issuing back to back,
    // one sub-tile per CTA (twice the tiles, 256 TMEM columns) beats two for 128-wide N tiles; 64-channel
    // outputs keep two sub-tiles (the weight tile is small, the halo overlap is what is left to save);
    // layers with few pixels fall back to 64-wide N tiles to give ~every SM a tile
    int b = (Nout % 128 == 0) ? 128 : 64;
    const long long th = ceil_div(H, kTileH);
    int s = b == 64 ? 2 : g_sub128;
    long long tiles = (long long)NB * th * ceil_div(W, kSubW * s) * ceil_div(Nout, b);
    if (tiles < (long long)num_sms() * 3 / 4) {
        s = 1;
        if (b == 128) b = 64;
    }
    if (g_force_bn == 64 || (g_force_bn == 128 && Nout % 128 == 0)) b = g_force_bn;
    if (g_force_sub == 1 || g_force_sub == 2) s = g_force_sub;
    // CTA pairs (one 256-pixel x b MMA per k-step, half a weight tile staged per SM).  The data gradient's
    // MN-major weight operand is staged in 64-column chunks, so a pair needs b >= 128 there.
    if (g_pair != 0) {
        const long long m_pairs = ((long long)NB * th * ceil_div(W, kSubW) + 1) / 2;
        const int min_b = dgrad ? 128 : 64;
        int pb = 0;
        if (g_force_bn == 256 && Nout % 256 == 0) pb = 256;
        else if (g_force_bn && g_force_bn >= min_b && Nout % g_force_bn == 0) pb = g_force_bn;
        else if (!g_force_bn) {
            // widest N tile that still gives ~every SM pair a tile group
            for (int cand = 256; cand >= min_b; cand >>= 1)
                if (Nout % cand == 0 && (m_pairs * (Nout / cand) >= (long long)(num_sms() / 2) * 3 / 4 || cand == min_b)) {
                    pb = cand;
                    break;
                }
        }
        const bool want = g_pair == 1 || pair_pays(Cred, Nout, pb, m_pairs);
        if (pb && want && Nout % pb == 0) {
            *bn = pb;
            *sub = 1;
            *cluster = 2;
            return true;
        }
    }
    if (b == 256) b = 128;          // (a forced 256 only exists as a pair)
    *bn = b;
    *sub = s;
    return true;
}

int scda_conv_halo_launch(int NB, int H, int W, int Cred, int Nout, const void *a, const void *w_krsc,
                          const float *bias, void *out, int flags, const void *mask_src, bool dgrad, int bn,
                          int sub, int cluster, cudaStream_t stream)
{
    const int region_w = kSubW * sub;
    HaloParams p = {};
    p.H = H; p.W = W; p.Cred = Cred; p.N = Nout;
    p.cblocks = ceil_div(Cred, 64);
    p.tiles_w = ceil_div(W, region_w);
    p.tiles_h = ceil_div(H, kTileH);
    p.m_tiles = NB * p.tiles_h * p.tiles_w;
    p.n_tiles = ceil_div(Nout, bn);
    p.bias = bias; p.out = out; p.ldc = Nout;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.flags = flags;
    // 64 -> 64 channels (conv1_x and its data gradient): the whole 72 KB of weights stays in shared
    // memory (a B ring of 8 KB tiles is latency bound there: 31 B/clk needed, ~1 us per TMA round trip)
    const bool resident = p.cblocks == 1 && p.n_tiles == 1 && bn == 64 && g_resident != 0;
    CUtensorMap ma, mb;
    cuuint64_t da[4] = {(cuuint64_t)Cred, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)Cred * 2, (cuuint64_t)W * Cred * 2, (cuuint64_t)H * W * Cred * 2};
    cuuint32_t ba[4] = {64, (cuuint32_t)(region_w + 2), (cuuint32_t)kHaloH, 1};
    if (!make_map(&ma, a, 4, da, sa, ba)) return 0;
    if (dgrad) {
        // forward weights [Cred = Cout_fwd rows][9 * Nout cols]
        cuuint64_t db[2] = {(cuuint64_t)9 * Nout, (cuuint64_t)Cred}, sb[1] = {(cuuint64_t)9 * Nout * 2};
        cuuint32_t bb[2] = {64, 64};
        if (!make_map(&mb, w_krsc, 2, db, sb, bb)) return 0;
        if (cluster == 2) {
            if (sub != 1) return 0;
            if (bn == 256) return launch_halo<256, 1, true, 8, false, 2>(ma, mb, p, stream);
            if (bn == 128) return launch_halo<128, 1, true, 8, false, 2>(ma, mb, p, stream);
            return 0;
        }
        if (bn == 64 && resident) return sub == 1 ? launch_halo<64, 1, true, 9>(ma, mb, p, stream)
                                                  : launch_halo<64, 2, true, 9>(ma, mb, p, stream);
        if (bn == 64) return sub == 1 ? launch_halo<64, 1, true, 8>(ma, mb, p, stream)
                                      : launch_halo<64, 2, true, 8>(ma, mb, p, stream);
        return sub == 1 ? launch_halo<128, 1, true, 8>(ma, mb, p, stream)
                        : launch_halo<128, 2, true, 8>(ma, mb, p, stream);
    }
    cuuint64_t db[2] = {(cuuint64_t)9 * Cred, (cuuint64_t)Nout}, sb[1] = {(cuuint64_t)9 * Cred * 2};
    cuuint32_t bb[2] = {64, (cuuint32_t)(bn / cluster)};
    if (!make_map(&mb, w_krsc, 2, db, sb, bb)) return 0;
    if (cluster == 2) {
        if (sub != 1) return 0;
        if (bn == 256) return launch_halo<256, 1, false, 8, false, 2>(ma, mb, p, stream);
        if (bn == 128) return launch_halo<128, 1, false, 8, false, 2>(ma, mb, p, stream);
        if (bn == 64 && resident) return launch_halo<64, 1, false, 9, false, 2>(ma, mb, p, stream);
        if (bn == 64) return launch_halo<64, 1, false, 8, false, 2>(ma, mb, p, stream);
        return 0;
    }
    if (bn == 64 && resident) return sub == 1 ? launch_halo<64, 1, false, 9>(ma, mb, p, stream)
                                              : launch_halo<64, 2, false, 9>(ma, mb, p, stream);
    if (bn == 64) return sub == 1 ? launch_halo<64, 1, false, 8>(ma, mb, p, stream)
                                  : launch_halo<64, 2, false, 8>(ma, mb, p, stream);
    return sub == 1 ? launch_halo<128, 1, false, 8>(ma, mb, p, stream)
                    : launch_halo<128, 2, false, 8>(ma, mb, p, stream);
}

// Stride-2 3x3 convolution (padding 1) and its data gradient on the halo kernel (see conv_halo_kernel, kS2).
//   forward: x bf16 [NB, 2*Ho, 2*Wo, C] (C % 32 == 0), wd bf16 [Cout][3][3][4C] = the s2d weight layout written by
//            scda_conv_s2_weights (disc_ops.cu) -> y [NB, Ho, Wo, Cout]
//   dgrad  : dy bf16 [NB, Ho, Wo, Cout], the same wd -> dx [NB, 2*Ho, 2*Wo, C]
// flags / bias / mask_src / slope: the epilogue of tc_epilogue.cuh (mask_src is indexed like the output).
int scda_conv_halo_s2_launch(int NB, int Ho, int Wo, int C, int Cout, const void *a, const void *wd,
                             const float *bias, void *out, int flags, const void *mask_src, float slope,
                             bool dgrad, cudaStream_t stream)
{
    if (C % 32 || Cout % 32 || NB <= 0 || Ho <= 0 || Wo <= 0) return 0;
    const int c2 = 2 * C, c4 = 4 * C;
    if (c2 % 64) return 0;
    const int Cred = dgrad ? Cout : c4, Nout = dgrad ? c4 : Cout;
    // N tile: 64 unless 128 fits (a data-gradient tile must stay inside one row phase of 2C channels)
    int bn = (Nout % 128 == 0 && (!dgrad || c2 % 128 == 0)) ? 128 : 64;
    const long long th = ceil_div(Ho, kTileH);
    if ((long long)NB * th * ceil_div(Wo, kSubW) * ceil_div(Nout, bn) < (long long)num_sms() * 3 / 4) bn = 64;
    const int sub = 1;
    HaloParams p = {};
    p.H = Ho; p.W = Wo; p.Cred = Cred; p.N = Nout;
    p.cblocks = ceil_div(Cred, 64);
    p.tiles_w = ceil_div(Wo, kSubW * sub);
    p.tiles_h = ceil_div(Ho, kTileH);
    p.m_tiles = NB * p.tiles_h * p.tiles_w;
    p.n_tiles = ceil_div(Nout, bn);
    p.bias = bias; p.out = out;
    p.ldc = dgrad ? c2 : Nout;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.flags = flags;
    p.c2 = c2;
    p.slope = slope;
    // forward taps: offsets (a, b) in {-1, 0}^2 of the 3x3 grid = taps 0, 1, 3, 4; the data gradient walks the
    // mirrored offsets {0, +1}^2 = taps 4, 5, 7, 8 (and reads weight tap 8 - tap)
    p.tap_mask = dgrad ? ((1 << 4) | (1 << 5) | (1 << 7) | (1 << 8)) : ((1 << 0) | (1 << 1) | (1 << 3) | (1 << 4));
    CUtensorMap ma, mb;
    if (dgrad) {
        cuuint64_t da[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)NB};
        cuuint64_t sa[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wo * Cout * 2, (cuuint64_t)Ho * Wo * Cout * 2};
        cuuint32_t ba[4] = {64, (cuuint32_t)(kSubW * sub + 2), (cuuint32_t)kHaloH, 1};
        if (!make_map(&ma, a, 4, da, sa, ba)) return 0;
        cuuint64_t db[2] = {(cuuint64_t)9 * Nout, (cuuint64_t)Cred}, sb[1] = {(cuuint64_t)9 * Nout * 2};
        cuuint32_t bb[2] = {64, 64};
        if (!make_map(&mb, wd, 2, db, sb, bb)) return 0;
        return bn == 64 ? launch_halo<64, 1, true, 8, true>(ma, mb, p, stream)
                        : launch_halo<128, 1, true, 8, true>(ma, mb, p, stream);
    }
    // x seen as [NB, Hin = 2Ho, Wo (pixel pairs), 2C]; rows walked with element stride 2 (one row phase per box)
    cuuint64_t da[4] = {(cuuint64_t)c2, (cuuint64_t)Wo, (cuuint64_t)2 * Ho, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)c2 * 2, (cuuint64_t)Wo * c2 * 2, (cuuint64_t)2 * Ho * Wo * c2 * 2};
    cuuint32_t ba[4] = {64, (cuuint32_t)(kSubW * sub + 2), (cuuint32_t)(2 * kHaloH), 1};
    cuuint32_t es[4] = {1, 1, 2, 1};
    if (!make_map_strided(&ma, a, 4, da, sa, ba, es)) return 0;
    cuuint64_t db[2] = {(cuuint64_t)9 * Cred, (cuuint64_t)Nout}, sb[1] = {(cuuint64_t)9 * Cred * 2};
    cuuint32_t bb[2] = {64, (cuuint32_t)bn};
    if (!make_map(&mb, wd, 2, db, sb, bb)) return 0;
    return bn == 64 ? launch_halo<64, 1, false, 8, true>(ma, mb, p, stream)
                    : launch_halo<128, 1, false, 8, true>(ma, mb, p, stream);
}
