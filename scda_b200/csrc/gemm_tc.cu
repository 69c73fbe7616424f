// tcgen05 tensor-core kernels for sm_100a: bf16 x bf16 -> fp32 (TMEM) GEMM and 3x3
// implicit-GEMM convolution with im2col folded into the TMA staging.
//
// Replaces the cuDNN / cuBLAS calls behind the reference's nn.Conv2d / nn.Linear on the
// detector's hot path (13 VGG convs: models/faster_rcnn/vgg_adver_expansion_cluster.py:101-114;
// RPN head: models/head.py:13-18; fc6/fc7/cls/loc: vgg_adver_expansion_cluster.py:46-60).
//
// One kernel, two A-operand addressing modes:
//   GEMM  C[M,N] = A[M,K] . B[N,K]^T      A tile = 2-D TMA box {64 k, 128 rows}
//   CONV  Y[n,h,w,co] = sum_{r,s,ci} X[n,h+r-1,w+s-1,ci] W[co,r,s,ci]   (NHWC, pad 1)
//         M = pixels of an 8x16 (TH x TW) spatial tile, K = 9 taps x Cin.  For k-block
//         (tap r,s ; channel block c0) the A tile is ONE 4-D TMA box {64 ch, TW, TH, 1}
//         at coordinates {c0, w0+s-1, h0+r-1, n}: the shifted window lands in shared
//         memory already in the K-major 128-byte-swizzled layout the MMA wants, and TMA's
//         out-of-bounds zero fill IS the padding.  No im2col buffer ever exists.
// Both operands are K-major bf16, 128B swizzle; accumulators are 128 x BLOCK_N fp32 in TMEM.
//
// Warp roles (256 threads, 1 CTA / SM):
//   warp 0   : TMA producer (one elected lane), kStages-deep mbarrier ring
//   warp 1   : MMA issuer (one elected lane): 4 x tcgen05.mma (K=16) per 64-wide k-block,
//              tcgen05.commit releases the smem stage / signals the epilogue
//   warp 2   : TMEM allocate / free
//   warps 4-7: epilogue: tcgen05.ld 32 lanes x 32 columns at a time -> bias, ReLU or the
//              ReLU-gradient mask of the layer input, bf16/fp32 pack -> global
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_halo.h"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;        // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 256;


struct TcParams {
    int M, N, K;                 // GEMM view (CONV: M = NB*H*W, K = 9*Cin)
    int num_k_blocks;
    int conv;                    // 0 GEMM, 1 CONV
    int H, W, Cin, TH, TW;       // CONV geometry
    int tiles_w, tiles_h;
    const float *bias;           // [N] or null
    void *out;                   // [M, ldc]
    long long ldc;
    const __nv_bfloat16 *mask_src;   // [M, ldc] (kFlagMaskPos)
    const __nv_bfloat16 *mul_src;    // [M, ldc] (kFlagMulSrc)
    int flags;
};

using namespace tcptx;


template <int kBlockN, int kStages, int kCluster = 1>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
    static constexpr int kBBytes = kBlockN / kCluster * kBlockK * 2;   // a CTA pair holds half of B each
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTileBytes = kStages * kStageBytes;
    static constexpr int kBarOffset = kTileBytes;           // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
    static constexpr int kNumBars = 2 * kStages + 4;
    static constexpr int kTotal = kTileBytes + kNumBars * 8 + 16 + 1024;  // + tmem ptr + align slack
};

struct TileCoord {
    int m_tile, n0;
    int img, h0, w0;
};

__device__ __forceinline__ TileCoord tile_coord(const TcParams &p, int tile, int n_tiles, int block_n,
                                                int cluster, int rank)
{
    TileCoord t;
    const int m_group = tile / n_tiles;
    t.m_tile = m_group * cluster + rank;
    t.n0 = (tile - m_group * n_tiles) * block_n;
    t.img = 0; t.h0 = 0; t.w0 = 0;
    if (p.conv) {
        const int per_img = p.tiles_h * p.tiles_w;
        t.img = t.m_tile / per_img;
        const int r = t.m_tile - t.img * per_img;
        t.h0 = (r / p.tiles_w) * p.TH;
        t.w0 = (r % p.tiles_w) * p.TW;
    }
    return t;
}

template <int kBlockN, int kCluster, int kSpec>
__device__ __forceinline__ void gemm_epilogue(const TcParams &p, const EpiParams &e, uint32_t tmem_base,
                                              uint32_t bar_tfull, uint32_t bar_tempty, uint32_t tempty_remote,
                                              int ew, int lane, int rank, int m_tiles, int n_tiles,
                                              int first_tile, int tile_step, int total_tiles)
{
    constexpr bool kPair = kCluster == 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(e.out) | reinterpret_cast<uintptr_t>(e.bias) |
                           reinterpret_cast<uintptr_t>(e.mask_src) | reinterpret_cast<uintptr_t>(e.mul_src)) & 15) == 0;
    const bool vec_ok = (e.ldc % 8 == 0) && aligned;
    uint32_t ti = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++ti) {
        const TileCoord t = tile_coord(p, tile, n_tiles, kBlockN, kCluster, rank);
        const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
        mbar_wait(bar_tfull + acc * 8, acc_ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = ew * 32 + lane;           // accumulator row = TMEM lane
        long long out_row;
        bool row_ok;
        if (p.conv) {
            const int th = row / p.TW, tw = row - th * p.TW;
            const int h = t.h0 + th, w = t.w0 + tw;
            row_ok = h < p.H && w < p.W && t.m_tile < m_tiles;   // (phantom tile of an odd group)
            out_row = ((long long)t.img * p.H + h) * p.W + w;
        } else {
            out_row = (long long)t.m_tile * kBlockM + row;
            row_ok = out_row < p.M;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < kBlockN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kBlockN + c0), v);
            if (c0 + 32 >= kBlockN) {
                // the whole accumulator of this warp's lanes is in registers: hand the buffer back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (kPair) mbar_arrive_cluster(tempty_remote + acc * 8);
                    else mbar_arrive(bar_tempty + acc * 8);
                }
            }
            if (row_ok) epilogue_chunk<kSpec>(v, e, out_row, t.n0 + c0, vec_ok && (t.n0 % 8 == 0));
        }
    }
}

// Persistent kernel: grid = min(#tile groups, #SMs / kCluster) clusters; cluster i walks tile
// groups i, i + #clusters, ... (n tile fastest).  The accumulator is double-buffered in TMEM
// (2 x kBlockN columns): the epilogue warps drain tile i while the MMA warp is already
// accumulating tile i+1 and the TMA warp is loading ahead of both.
//
// kCluster == 2 is the CTA-PAIR form (tcgen05 cta_group::2): the two CTAs of a cluster (same
// TPC) own two consecutive m tiles of the same n tile and execute ONE 256 x kBlockN MMA per
// k-step.  Each CTA stages its own A tile and only HALF of the B tile; the tensor core of each
// SM reads the other half straight out of the peer's shared memory.  The kernel is bound by
// the bytes an SM can pull through the L2 fabric (~42 B/clk/SM with all SMs loading,
// profiles/r1_ncu_conv_*): per 128 x 256 x 64 block of work a lone CTA stages 48 KB, a paired
// CTA 32 KB.  Protocol: both producers complete their bytes on the LEADER's `full` barrier;
// only the leader's MMA lane issues; its commits arrive on the `empty` / `tmem_full` barriers of
// BOTH CTAs; the epilogue warps of both CTAs arrive on the leader's `tmem_empty`.
template <int kBlockN, int kStages, bool kBMn, int kCluster>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TcParams p, const int m_tiles, const int n_tiles)
{
    static_assert(kCluster == 1 || kCluster == 2, "one CTA or a CTA pair");
    constexpr bool kPair = kCluster == 2;
    using L = SmemLayout<kBlockN, kStages, kCluster>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024 B alignment
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base + L::kBarOffset;
    const uint32_t bar_empty = bar_full + kStages * 8;
    const uint32_t bar_tfull = bar_empty + kStages * 8;     // [2] accumulator ready
    const uint32_t bar_tempty = bar_tfull + 2 * 8;          // [2] accumulator drained
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + L::kBarOffset + L::kNumBars * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = kPair ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int m_groups = (m_tiles + kCluster - 1) / kCluster;
    const int total_tiles = m_groups * n_tiles;              // tile groups
    const int first_tile = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + a * 8, 1);
            mbar_init(bar_tempty + a * 8, 4 * kCluster);   // one arrival per epilogue warp (of the pair)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        const uint32_t ncols = 2 * kBlockN;
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32((const void *)tmem_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32((const void *)tmem_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kPair) cluster_sync_all();       // both CTAs' barriers exist before any remote traffic
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const int cblocks = p.conv ? p.Cin / kBlockK : 0;
            // the barrier the loads complete on: this CTA's, or (pair) the leader's
            const uint32_t full_remote = kPair ? map_to_cta(bar_full, 0) : bar_full;
            constexpr int kHalfN = kBlockN / kCluster;           // B columns staged by this CTA
            uint32_t it = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
                const TileCoord t = tile_coord(p, tile, n_tiles, kBlockN, kCluster, rank);
                const int nb = t.n0 + rank * kHalfN;
                for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(bar_empty + s * 8, ph ^ 1);
                    const uint32_t a_dst = base + s * L::kStageBytes;
                    const uint32_t b_dst = a_dst + L::kABytes;
                    const uint32_t fb = full_remote + s * 8;
                    if (leader) mbar_expect_tx(bar_full + s * 8, L::kStageBytes * kCluster);
                    int b_k = kb * kBlockK, b_col = 0;
                    if (p.conv) {
                        const int tap = kb / cblocks, cb = kb - tap * cblocks;
                        const int r = tap / 3, sx = tap - r * 3;
                        if (kPair)
                            tma_load_4d_pair(a_dst, &map_a, fb, cb * kBlockK, t.w0 + sx - 1, t.h0 + r - 1, t.img);
                        else
                            tma_load_4d(a_dst, &map_a, fb, cb * kBlockK, t.w0 + sx - 1, t.h0 + r - 1, t.img);
                        // K-major weights [n][tap, ci]: column = tap * Cin + ci.  MN-major (data
                        // gradient straight from the forward weights W[co][tap][ci] seen as
                        // [co rows][9*N cols]): reduction index = co (rows), output channel = ci
                        // (contiguous), tap mirrored (r, s) -> (2 - r, 2 - s)
                        b_k = kBMn ? cb * kBlockK : tap * p.Cin + cb * kBlockK;
                        b_col = kBMn ? (8 - tap) * p.N : 0;
                    } else {
                        if (kPair) tma_load_2d_pair(a_dst, &map_a, fb, kb * kBlockK, t.m_tile * kBlockM);
                        else tma_load_2d(a_dst, &map_a, fb, kb * kBlockK, t.m_tile * kBlockM);
                    }
                    if (kBMn) {
                        // B given as [K rows][N contiguous]: one {64 n, 64 k} box per 64-wide n chunk
#pragma unroll
                        for (int c = 0; c < kHalfN / 64; ++c) {
                            if (kPair) tma_load_2d_pair(b_dst + c * (kBlockK * 128), &map_b, fb, b_col + nb + c * 64, b_k);
                            else tma_load_2d(b_dst + c * (kBlockK * 128), &map_b, fb, b_col + nb + c * 64, b_k);
                        }
                    } else {
                        if (kPair) tma_load_2d_pair(b_dst, &map_b, fb, b_k, nb);
                        else tma_load_2d(b_dst, &map_b, fb, b_k, nb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one() && leader) {
            constexpr uint32_t idesc = make_idesc(kBlockM * kCluster, kBlockN, 0, kBMn ? 1 : 0);
            // descriptors as (lo, hi) words: hi constant, lo = (start >> 4) | LBO field; between MMAs
            // only a constant is added to lo (conv_halo.cu has the measurement behind this)
            const uint64_t a_proto = make_kmajor_desc(0);
            const uint64_t b_proto = kBMn ? make_mnmajor_desc(0, kBlockK * 128) : make_kmajor_desc(0);
            const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
            const uint32_t a_lo0 = (uint32_t)a_proto + (base >> 4);
            const uint32_t b_lo0 = (uint32_t)b_proto + ((base + L::kABytes) >> 4);
            constexpr uint32_t kBk = kBMn ? (kUmmaK * 128) >> 4 : (kUmmaK * 2) >> 4;
            uint32_t it = 0, ti = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++ti) {
                const uint32_t acc = ti & 1, acc_ph = (ti >> 1) & 1;
                mbar_wait(bar_tempty + acc * 8, acc_ph ^ 1);     // epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * kBlockN;
                for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
                    const uint32_t s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(bar_full + s * 8, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_lo = a_lo0 + s * (L::kStageBytes >> 4);
                    const uint32_t b_lo = b_lo0 + s * (L::kStageBytes >> 4);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        if (kPair) umma_bf16_pair_lohi(tmem_d, a_lo + k * 2, a_hi, b_lo + k * kBk, b_hi, idesc,
                                                       (uint32_t)(kb | k));
                        else umma_bf16_lohi(tmem_d, a_lo + k * 2, a_hi, b_lo + k * kBk, b_hi, idesc,
                                            (uint32_t)(kb | k));
                    }
                    // frees the stage (in both CTAs of a pair) when these MMAs retire
                    if (kPair) umma_commit_pair(bar_empty + s * 8);
                    else umma_commit(bar_empty + s * 8);
                }
                if (kPair) umma_commit_pair(bar_tfull + acc * 8);     // accumulator complete
                else umma_commit(bar_tfull + acc * 8);
            }
        }
    } else if (warp >= 4) {
        // epilogue, flag set resolved at compile time for the combinations the detector uses
        EpiParams e;
        e.bias = p.bias; e.out = p.out; e.ldc = p.ldc; e.mask_src = p.mask_src; e.mul_src = p.mul_src;
        e.flags = p.flags | (p.bias ? kFlagBias : 0);
        e.N = p.N;
        e.slope = 0.f;
        const uint32_t tempty_remote = kPair ? map_to_cta(bar_tempty, 0) : bar_tempty;
#define SCDA_EPI(SPEC)                                                                                         \
    gemm_epilogue<kBlockN, kCluster, SPEC>(p, e, tmem_base, bar_tfull, bar_tempty, tempty_remote, warp - 4,   \
                                           lane, rank, m_tiles, n_tiles, first_tile, tile_step, total_tiles)
        if (e.flags == (kFlagRelu | kFlagBias)) SCDA_EPI(kFlagRelu | kFlagBias);
        else if (e.flags == (kFlagRelu | kFlagBias | kFlagMulSrc)) SCDA_EPI(kFlagRelu | kFlagBias | kFlagMulSrc);
        else if (e.flags == kFlagMaskPos) SCDA_EPI(kFlagMaskPos);
        else if (e.flags == (kFlagMaskPos | kFlagMulSrc)) SCDA_EPI(kFlagMaskPos | kFlagMulSrc);
        else if (e.flags == 0) SCDA_EPI(0);
        // fp32-parity mode (x3_ops.cu): fp32 outputs, fp32 ReLU masks
        else if (e.flags == (kFlagRelu | kFlagBias | kFlagOutF32)) SCDA_EPI(kFlagRelu | kFlagBias | kFlagOutF32);
        else if (e.flags == (kFlagMaskPos | kFlagMaskF32 | kFlagOutF32)) SCDA_EPI(kFlagMaskPos | kFlagMaskF32 | kFlagOutF32);
        else if (e.flags == kFlagOutF32) SCDA_EPI(kFlagOutF32);
        else SCDA_EPI(-1);
#undef SCDA_EPI
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kPair) cluster_sync_all();       // no CTA leaves while its peer may still signal or read it
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ncols = 2 * kBlockN;
        if (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols)
                         : "memory");
    }
}

// ------------------------------------------------------------------- host side
template <int kBlockN, int kStages, bool kBMn, int kCluster>
int launch_tc_c(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int m_tiles, cudaStream_t stream)
{
    using L = SmemLayout<kBlockN, kStages, kCluster>;
    static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<kBlockN, kStages, kBMn, kCluster>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    const int n_tiles = ceil_div(p.N, kBlockN);
    const long long groups = (long long)ceil_div(m_tiles, kCluster) * n_tiles;
    const int max_clusters = num_sms() / kCluster;
    const int clusters = (int)(groups < max_clusters ? groups : max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCluster);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<kBlockN, kStages, kBMn, kCluster>, ma, mb, p, m_tiles,
                                       n_tiles);
    if (e != cudaSuccess) return -(int)e;
    return scda_launch_status();
}

int dispatch_tc(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int m_tiles, int block_n,
                int cluster, cudaStream_t stream, bool b_mn = false)
{
    // stages: as many as fit in ~200 KB (stage = 16 KB of A + this CTA's share of B)
    if (cluster == 2) {
        if (b_mn) {
            if (block_n == 128) return launch_tc_c<128, 8, true, 2>(ma, mb, p, m_tiles, stream);
            return launch_tc_c<256, 6, true, 2>(ma, mb, p, m_tiles, stream);
        }
        if (block_n == 64) return launch_tc_c<64, 9, false, 2>(ma, mb, p, m_tiles, stream);
        if (block_n == 128) return launch_tc_c<128, 8, false, 2>(ma, mb, p, m_tiles, stream);
        return launch_tc_c<256, 6, false, 2>(ma, mb, p, m_tiles, stream);
    }
    if (b_mn) {
        if (block_n == 64) return launch_tc_c<64, 8, true, 1>(ma, mb, p, m_tiles, stream);
        if (block_n == 128) return launch_tc_c<128, 6, true, 1>(ma, mb, p, m_tiles, stream);
        return launch_tc_c<256, 4, true, 1>(ma, mb, p, m_tiles, stream);
    }
    if (block_n == 64) return launch_tc_c<64, 8, false, 1>(ma, mb, p, m_tiles, stream);
    if (block_n == 128) return launch_tc_c<128, 6, false, 1>(ma, mb, p, m_tiles, stream);
    return launch_tc_c<256, 4, false, 1>(ma, mb, p, m_tiles, stream);
}

// N tile: 64 for narrow outputs; 256 when that still leaves every SM a tile (fewer bytes
// staged per flop), else 128.
int pick_block_n(int N, long long m_tiles = 1 << 30)
{
    if (N <= 64) return 64;
    if (N % 256 == 0 && m_tiles * (N / 256) >= num_sms()) return 256;
    return 128;
}

// Tile plan: N tile width and one CTA (cl = 1) or a CTA pair (cl = 2) per tile group.  The pair
// halves the B bytes each SM stages.  Measured on B200 (profiles/r1_convbench_c_pairs.txt): the
// 256-wide pair wins where the reduction is long (K >= 4096: conv4_2/4_3 58 -> 52 us, fc6 data
// gradient 128 -> 99 us = 1.06 PFLOP/s) and there are enough tile groups for ~all 74 SM pairs;
// it ties at K = 2304 and the 64/128-wide pairs LOSE 5-10 % (cluster launch + cross-CTA barrier
// latency is not amortised), so those keep the one-CTA form.
// SCDA_TC_CLUSTER=1 forces one CTA, =2 forces pairs wherever they are legal (tests, experiments).
struct TilePlan {
    int bn, cl;
};

TilePlan plan_tiles(int N, int K, long long m_tiles, bool b_mn)
{
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("SCDA_TC_CLUSTER");
        forced = e ? atoi(e) : 0;
        if (forced != 1 && forced != 2) forced = 0;
    }
    if (forced != 1 && m_tiles >= 2) {
        const long long pairs = (m_tiles + 1) / 2;
        if (forced == 2) {
            if (N % 256 == 0) return {256, 2};
            if (N > 64) return {128, 2};
            if (!b_mn) return {64, 2};
        } else if (N % 256 == 0 && K >= 4096 && pairs * (N / 256) >= (long long)(num_sms() / 2) * 4 / 5) {
            return {256, 2};
        }
    }
    return {pick_block_n(N, m_tiles), 1};
}

// ---------------------------------------------------------------------------------------
// Weight-gradient kernel: C[Mo, No] (+)= sum_k A[k, Mo] B[k, No], both operands given with
// the REDUCTION index as the row (slow) dimension, i.e. MN-major for the MMA:
//   linear : dW[out, in]      = sum_rows  dY[row, out]      X[row, in]          (2-D boxes)
//   conv   : dW[co, r, s, ci] = sum_pixel dY[pixel, co]     X[pixel + (r-1, s-1), ci]
//            one (tap, co tile, ci tile, k split) per CTA; the shifted X window is one 4-D
//            TMA box per 64-channel chunk, zero-filled outside the image.
// k-block = 128 reduction rows (one 8x16 pixel tile) -> 8 MMAs of K = 16.
// Partial sums of a k split go to their own fp32 slab (deterministic); the caller reduces.
struct WgParams {
    int Mo, No;                  // output tile space (Cout, Cin)
    int conv;
    int H, W, TH, TW, tiles_w, tiles_h;
    int total_k_blocks, k_per_split;
    float *out;                  // [splits][Mo][ldo]
    long long ldo;               // row stride of the output (conv: 9*Cin)
    long long split_stride;
    int accumulate;              // dst += (linear form only)
    // stride-2 convolution (conv_halo.cu, kS2): No = 4C virtual input channels (row phase, column phase,
    // channel), c2 = 2C; the X map views the input as [NB, 2H, W, 2C] with element stride 2 along rows;
    // tap_list holds the ntaps taps of the 3x3 grid that are computed, 4 bits each (0 = all nine in order)
    int s2, c2, ntaps;
    unsigned long long tap_list;
};

template <int kBlockN, int kStages>
struct WgSmem {
    static constexpr int kChunk = 128 * 128;                 // [128 k rows][64 mn] bf16
    static constexpr int kABytes = 2 * kChunk;               // Mo tile 128 = 2 chunks
    static constexpr int kBBytes = (kBlockN / 64) * kChunk;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * kStages + 1) * 8 + 16 + 1024;
};

template <int kBlockN, int kStages>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const WgParams p)
{
    using L = WgSmem<kBlockN, kStages>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base + L::kBarOffset;
    const uint32_t bar_empty = bar_full + kStages * 8;
    const uint32_t bar_tmem = bar_empty + kStages * 8;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + L::kBarOffset + (2 * kStages + 1) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM;                 // Cout offset
    const int n_tiles = (p.No + kBlockN - 1) / kBlockN;
    const int tap_ix = p.conv ? blockIdx.y / n_tiles : 0;
    const int tap = p.ntaps ? (int)((p.tap_list >> (4 * tap_ix)) & 15) : tap_ix;
    const int n0 = (blockIdx.y - tap_ix * n_tiles) * kBlockN;   // Cin offset
    const int split = blockIdx.z;
    const int kb0 = split * p.k_per_split;
    const int kb1 = min(kb0 + p.k_per_split, p.total_k_blocks);
    const int r = tap / 3, sx = tap - r * 3;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        const uint32_t ncols = kBlockN;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32((const void *)tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(bar_empty + s * 8, ph ^ 1);
                const uint32_t a_dst = base + s * L::kStageBytes;
                const uint32_t b_dst = a_dst + L::kABytes;
                mbar_expect_tx(bar_full + s * 8, L::kStageBytes);
                if (p.conv) {
                    const int per_img = p.tiles_h * p.tiles_w;
                    const int img = kb / per_img, t = kb - img * per_img;
                    const int h0 = (t / p.tiles_w) * p.TH, w0 = (t % p.tiles_w) * p.TW;
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_4d(a_dst + c * L::kChunk, &map_a, bar_full + s * 8, m0 + c * 64, w0, h0, img);
                    if (p.s2) {
                        const int py = n0 / p.c2;
#pragma unroll
                        for (int c = 0; c < kBlockN / 64; ++c)
                            tma_load_4d(b_dst + c * L::kChunk, &map_b, bar_full + s * 8, n0 - py * p.c2 + c * 64,
                                        w0 + sx - 1, 2 * (h0 + r - 1) + py, img);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockN / 64; ++c)
                            tma_load_4d(b_dst + c * L::kChunk, &map_b, bar_full + s * 8, n0 + c * 64, w0 + sx - 1,
                                        h0 + r - 1, img);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_2d(a_dst + c * L::kChunk, &map_a, bar_full + s * 8, m0 + c * 64, kb * 128);
#pragma unroll
                    for (int c = 0; c < kBlockN / 64; ++c)
                        tma_load_2d(b_dst + c * L::kChunk, &map_b, bar_full + s * 8, n0 + c * 64, kb * 128);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(kBlockM, kBlockN, 1, 1);
            const uint64_t proto = make_mnmajor_desc(0, L::kChunk);
            const uint32_t proto_lo = (uint32_t)proto, proto_hi = (uint32_t)(proto >> 32);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(bar_full + s * 8, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_lo = proto_lo + ((base + s * L::kStageBytes) >> 4);
                const uint32_t b_lo = a_lo + (L::kABytes >> 4);
#pragma unroll
                for (int k = 0; k < 128 / kUmmaK; ++k)
                    umma_bf16_lohi(tmem_base, a_lo + k * ((kUmmaK * 128) >> 4), proto_hi, b_lo + k * ((kUmmaK * 128) >> 4),
                                   proto_hi, idesc, (uint32_t)(it | k));
                umma_commit(bar_empty + s * 8);
            }
            umma_commit(bar_tmem);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + ew * 32 + lane;
        float *dst_row = p.out + (long long)split * p.split_stride + (long long)row * p.ldo +
                         (p.conv ? (long long)tap * p.No : 0);
        const bool vec_ok = (p.ldo % 4 == 0) && (p.No % 4 == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < kBlockN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
            if (row >= p.Mo) continue;
            const int ncol = min(32, p.No - (n0 + c0));
            if (ncol <= 0) continue;
            float *dst = dst_row + n0 + c0;
            if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < ncol) dst[j] += __uint_as_float(v[j]);
            } else if (ncol == 32 && vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(dst + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                    __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < ncol) dst[j] = __uint_as_float(v[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ncols = kBlockN;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// Convolution weight gradient, three taps per CTA.  The per-tap form above re-reads the dY tile
// and the (shifted) X tile for each of the nine taps: 64 KB staged per 8 MMAs of 128x128x16 =
// 128 B/clk/SM against the ~42 B/clk/SM the L2 fabric delivers.  Here a CTA owns one kernel
// COLUMN sx and accumulates its three taps (r = 0, 1, 2) side by side in TMEM: per 128-pixel
// k-block it stages the dY tile once and ONE X box that is two image rows taller
// ({64 ch, TW, TH + 2, 1} at {ci0, w0 + sx - 1, h0 - 1, n}); tap r reads the same bytes through
// an MN-major descriptor whose start is r image rows (r * TW pixel rows of 128 B) further down —
// the same shifted-descriptor idea as conv_halo.cu, applied to the reduction dimension.
// 72 KB per 24 MMAs = 47 B/clk/SM.
template <int kBlockN, int kStages>
struct Wg3Smem {
    static constexpr int kAChunk = 128 * 128;                // [128 k rows][64 co] bf16
    static constexpr int kABytes = 2 * kAChunk;              // Cout tile 128 = 2 chunks
    static constexpr int kBChunk = 160 * 128;                // up to (TH + 2) * TW = 160 pixel rows x 64 ci
    static constexpr int kBBytes = (kBlockN / 64) * kBChunk;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * kStages + 1) * 8 + 16 + 1024;
    static constexpr uint32_t kTmemCols = 3 * kBlockN <= 256 ? 256 : 512;
};

template <int kBlockN, int kStages>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const WgParams p)
{
    using L = Wg3Smem<kBlockN, kStages>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_full = base + L::kBarOffset;
    const uint32_t bar_empty = bar_full + kStages * 8;
    const uint32_t bar_tmem = bar_empty + kStages * 8;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + L::kBarOffset + (2 * kStages + 1) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM;                 // Cout offset
    const int n_tiles = (p.No + kBlockN - 1) / kBlockN;
    const int sx = blockIdx.y / n_tiles;                 // kernel column 0..2
    const int n0 = (blockIdx.y - sx * n_tiles) * kBlockN;    // Cin offset
    const int split = blockIdx.z;
    const int kb0 = split * p.k_per_split;
    const int kb1 = min(kb0 + p.k_per_split, p.total_k_blocks);
    const uint32_t b_box_bytes = (uint32_t)((p.TH + 2) * p.TW * 128);

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32((const void *)tmem_slot)), "r"(L::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(bar_empty + s * 8, ph ^ 1);
                const uint32_t a_dst = base + s * L::kStageBytes;
                const uint32_t b_dst = a_dst + L::kABytes;
                mbar_expect_tx(bar_full + s * 8, L::kABytes + (kBlockN / 64) * b_box_bytes);
                const int per_img = p.tiles_h * p.tiles_w;
                const int img = kb / per_img, t = kb - img * per_img;
                const int h0 = (t / p.tiles_w) * p.TH, w0 = (t % p.tiles_w) * p.TW;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    tma_load_4d(a_dst + c * L::kAChunk, &map_a, bar_full + s * 8, m0 + c * 64, w0, h0, img);
#pragma unroll
                for (int c = 0; c < kBlockN / 64; ++c)
                    tma_load_4d(b_dst + c * L::kBChunk, &map_b, bar_full + s * 8, n0 + c * 64, w0 + sx - 1, h0 - 1,
                                img);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(kBlockM, kBlockN, 1, 1);
            const uint64_t a_proto = make_mnmajor_desc(0, L::kAChunk), b_proto = make_mnmajor_desc(0, L::kBChunk);
            const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
            const uint32_t row_step = (uint32_t)(p.TW * 128) >> 4;       // one image row of the X box, 16 B units
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(bar_full + s * 8, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_lo = (uint32_t)a_proto + ((base + s * L::kStageBytes) >> 4);
                const uint32_t b_lo = (uint32_t)b_proto + ((base + s * L::kStageBytes + L::kABytes) >> 4);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int k = 0; k < 128 / kUmmaK; ++k)
                        umma_bf16_lohi(tmem_base + r * kBlockN, a_lo + k * ((kUmmaK * 128) >> 4), a_hi,
                                       b_lo + r * row_step + k * ((kUmmaK * 128) >> 4), b_hi, idesc,
                                       (uint32_t)(it | k));
                }
                umma_commit(bar_empty + s * 8);
            }
            umma_commit(bar_tmem);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + ew * 32 + lane;
        const bool vec_ok = (p.ldo % 4 == 0) && (p.No % 4 == 0);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            const int tap = r * 3 + sx;
            float *dst_row = p.out + (long long)split * p.split_stride + (long long)row * p.ldo + (long long)tap * p.No;
#pragma unroll 1
            for (int c0 = 0; c0 < kBlockN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(r * kBlockN + c0), v);
                if (row >= p.Mo) continue;
                const int ncol = min(32, p.No - (n0 + c0));
                if (ncol <= 0) continue;
                float *dst = dst_row + n0 + c0;
                if (ncol == 32 && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(dst + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                        __uint_as_float(v[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncol) dst[j] = __uint_as_float(v[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::kTmemCols)
                     : "memory");
    }
}

template <int kBlockN, int kStages>
int launch_wg3(const CUtensorMap &ma, const CUtensorMap &mb, const WgParams &p, int splits, cudaStream_t stream)
{
    using L = Wg3Smem<kBlockN, kStages>;
    static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad3_kernel<kBlockN, kStages>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    dim3 grid(ceil_div(p.Mo, kBlockM), 3 * ceil_div(p.No, kBlockN), splits);
    tc_wgrad3_kernel<kBlockN, kStages><<<grid, kThreads, L::kTotal, stream>>>(ma, mb, p);
    return scda_launch_status();
}

template <int kBlockN, int kStages>
int launch_wg(const CUtensorMap &ma, const CUtensorMap &mb, const WgParams &p, int taps, int splits,
              cudaStream_t stream)
{
    using L = WgSmem<kBlockN, kStages>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<kBlockN, kStages>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return -(int)e;
        attr_done = true;
    }
    dim3 grid(ceil_div(p.Mo, kBlockM), taps * ceil_div(p.No, kBlockN), splits);
    tc_wgrad_kernel<kBlockN, kStages><<<grid, kThreads, L::kTotal, stream>>>(ma, mb, p);
    return scda_launch_status();
}

}  // namespace

SCDA_API int scda_gemm_bf16_tn(int M, int N, int K, const void *A, long long lda, const void *B, long long ldb,
                               const float *bias, void *C, long long ldc, int flags, const void *mask_src,
                               const void *mul_src, cudaStream_t stream)
{
    if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) return 0;
    if (lda % 8 || ldb % 8 || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return 0;
    if ((flags & kFlagMaskPos) && !mask_src) return 0;
    if ((flags & kFlagMulSrc) && !mul_src) return 0;
    if ((flags & kFlagAccumulate) && !(flags & kFlagOutF32)) return 0;
    const TilePlan tp = plan_tiles(N, K, ceil_div(M, kBlockM), false);
    const int bn = tp.bn, cl = tp.cl;
    CUtensorMap ma, mb;
    cuuint64_t da[2] = {(cuuint64_t)K, (cuuint64_t)M}, sa[1] = {(cuuint64_t)lda * 2};
    cuuint32_t ba[2] = {kBlockK, kBlockM};
    cuuint64_t db[2] = {(cuuint64_t)K, (cuuint64_t)N}, sb[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t bb[2] = {kBlockK, (cuuint32_t)(bn / cl)};
    if (!make_map(&ma, A, 2, da, sa, ba) || !make_map(&mb, B, 2, db, sb, bb)) return 0;
    TcParams p = {};
    p.M = M; p.N = N; p.K = K;
    p.num_k_blocks = ceil_div(K, kBlockK);
    p.conv = 0;
    p.bias = bias; p.out = C; p.ldc = ldc;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.mul_src = (const __nv_bfloat16 *)mul_src;
    p.flags = flags;
    return dispatch_tc(ma, mb, p, ceil_div(M, kBlockM), bn, cl, stream);
}

SCDA_API int scda_conv3x3_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *x, const void *w_krsc,
                                    const float *bias, void *y, int flags, const void *mask_src,
                                    cudaStream_t stream)
{
    if (NB <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !x || !w_krsc || !y) return 0;
    if (Cin % kBlockK) return 0;               // channel blocks of 64 (conv1_1 has its own kernel)
    if ((flags & kFlagMaskPos) && !mask_src) return 0;
    if ((flags & kFlagAccumulate) && !(flags & kFlagOutF32)) return 0;
    {
        int hbn, hsub, hcl;  // halo form (conv_halo.cu): the input tile is staged once for all 9 taps
        if (!(flags & kFlagMulSrc) && scda_conv_halo_plan(NB, H, W, Cin, Cout, false, &hbn, &hsub, &hcl))
            return scda_conv_halo_launch(NB, H, W, Cin, Cout, x, w_krsc, bias, y, flags, mask_src, false, hbn,
                                         hsub, hcl, stream);
    }
    int TW = 16, TH = 8;
    if (W % 16) {                                 // narrow maps: 8 x 16 tile turned around
        if (W % 8 == 0) { TW = 8; TH = 16; } else return 0;
    }
    const int tiles_w = W / TW, tiles_h = ceil_div(H, TH);
    const TilePlan tp = plan_tiles(Cout, 9 * Cin, (long long)NB * tiles_h * tiles_w, false);
    const int bn = tp.bn, cl = tp.cl;
    CUtensorMap ma, mb;
    cuuint64_t da[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t ba[4] = {kBlockK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint64_t db[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout}, sb[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t bb[2] = {kBlockK, (cuuint32_t)(bn / cl)};
    if (!make_map(&ma, x, 4, da, sa, ba) || !make_map(&mb, w_krsc, 2, db, sb, bb)) return 0;
    TcParams p = {};
    p.M = NB * H * W; p.N = Cout; p.K = 9 * Cin;
    p.num_k_blocks = 9 * (Cin / kBlockK);
    p.conv = 1;
    p.H = H; p.W = W; p.Cin = Cin; p.TH = TH; p.TW = TW; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.bias = bias; p.out = y; p.ldc = Cout;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.flags = flags;
    return dispatch_tc(ma, mb, p, NB * tiles_h * tiles_w, bn, cl, stream);
}

SCDA_API int scda_conv3x3_dgrad_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *dy,
                                          const void *w_krsc, void *dx, int flags, const void *mask_src,
                                          cudaStream_t stream)
{
    // dx[n,h,w,ci] = sum_{r,s,co} dy[n,h-(r-1),w-(s-1),co] W[co][r][s][ci]: the forward kernel with
    // A = dy, reduction over (tap, co), and B read from the FORWARD weights as an MN-major operand.
    if (NB <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !dy || !w_krsc || !dx) return 0;
    if (Cin % 64 || Cout % 32) return 0;
    if ((flags & kFlagMaskPos) && !mask_src) return 0;
    if (flags & (kFlagAccumulate | kFlagMulSrc | kFlagRelu)) return 0;
    {
        int hbn, hsub, hcl;
        if (scda_conv_halo_plan(NB, H, W, Cout, Cin, true, &hbn, &hsub, &hcl))
            return scda_conv_halo_launch(NB, H, W, Cout, Cin, dy, w_krsc, nullptr, dx, flags, mask_src, true, hbn,
                                         hsub, hcl, stream);
    }
    if (Cout % kBlockK) return 0;              // (the per-tap form needs whole 64-channel blocks)
    int TW = 16, TH = 8;
    if (W % 16) {
        if (W % 8 == 0) { TW = 8; TH = 16; } else return 0;
    }
    const int tiles_w = W / TW, tiles_h = ceil_div(H, TH);
    const TilePlan tp = plan_tiles(Cin, 9 * Cout, (long long)NB * tiles_h * tiles_w, true);
    const int bn = tp.bn, cl = tp.cl;
    CUtensorMap ma, mb;
    cuuint64_t da[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
    cuuint32_t ba[4] = {kBlockK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint64_t db[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout}, sb[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t bb[2] = {64, kBlockK};
    if (!make_map(&ma, dy, 4, da, sa, ba) || !make_map(&mb, w_krsc, 2, db, sb, bb)) return 0;
    TcParams p = {};
    p.M = NB * H * W; p.N = Cin; p.K = 9 * Cout;
    p.num_k_blocks = 9 * (Cout / kBlockK);
    p.conv = 1;
    p.H = H; p.W = W; p.Cin = Cout; p.TH = TH; p.TW = TW; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.bias = nullptr; p.out = dx; p.ldc = Cin;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.flags = flags;
    return dispatch_tc(ma, mb, p, NB * tiles_h * tiles_w, bn, cl, stream, true);
}

SCDA_API int scda_gemm_bf16_nn(int M, int N, int K, const void *A, long long lda, const void *B, long long ldb,
                               const float *bias, void *C, long long ldc, int flags, const void *mask_src,
                               const void *mul_src, cudaStream_t stream)
{
    // C[M,N] = A[M,K] . B[K,N]   (B row-major with N contiguous: the MN-major operand form)
    if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) return 0;
    if (lda % 8 || ldb % 8 || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return 0;
    if ((flags & kFlagMaskPos) && !mask_src) return 0;
    if ((flags & kFlagMulSrc) && !mul_src) return 0;
    if ((flags & kFlagAccumulate) && !(flags & kFlagOutF32)) return 0;
    const TilePlan tp = plan_tiles(N, K, ceil_div(M, kBlockM), true);
    const int bn = tp.bn, cl = tp.cl;
    CUtensorMap ma, mb;
    cuuint64_t da[2] = {(cuuint64_t)K, (cuuint64_t)M}, sa[1] = {(cuuint64_t)lda * 2};
    cuuint32_t ba[2] = {kBlockK, kBlockM};
    cuuint64_t db[2] = {(cuuint64_t)N, (cuuint64_t)K}, sb[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t bb[2] = {64, kBlockK};
    if (!make_map(&ma, A, 2, da, sa, ba) || !make_map(&mb, B, 2, db, sb, bb)) return 0;
    TcParams p = {};
    p.M = M; p.N = N; p.K = K;
    p.num_k_blocks = ceil_div(K, kBlockK);
    p.bias = bias; p.out = C; p.ldc = ldc;
    p.mask_src = (const __nv_bfloat16 *)mask_src;
    p.mul_src = (const __nv_bfloat16 *)mul_src;
    p.flags = flags;
    return dispatch_tc(ma, mb, p, ceil_div(M, kBlockM), bn, cl, stream, true);
}

SCDA_API int scda_linear_wgrad_bf16(int rows, int Nout, int Kin, const void *dY, long long lddy, const void *X,
                                    long long ldx, float *dW, long long lddw, int accumulate,
                                    cudaStream_t stream)
{
    // dW[Nout, Kin] = dY[rows, Nout]^T . X[rows, Kin]
    if (rows <= 0 || Nout <= 0 || Kin <= 0 || !dY || !X || !dW) return 0;
    if (lddy % 8 || ldx % 8 || ((uintptr_t)dY % 16) || ((uintptr_t)X % 16)) return 0;
    const int bn = Kin <= 64 ? 64 : 128;
    CUtensorMap ma, mb;
    cuuint64_t da[2] = {(cuuint64_t)Nout, (cuuint64_t)rows}, sa[1] = {(cuuint64_t)lddy * 2};
    cuuint64_t db[2] = {(cuuint64_t)Kin, (cuuint64_t)rows}, sb[1] = {(cuuint64_t)ldx * 2};
    cuuint32_t box[2] = {64, 128};
    if (!make_map(&ma, dY, 2, da, sa, box) || !make_map(&mb, X, 2, db, sb, box)) return 0;
    WgParams p = {};
    p.Mo = Nout; p.No = Kin; p.conv = 0;
    p.total_k_blocks = ceil_div(rows, 128);
    p.k_per_split = p.total_k_blocks;
    p.out = dW; p.ldo = lddw; p.split_stride = 0;
    p.accumulate = accumulate ? 1 : 0;
    if (bn == 64) return launch_wg<64, 4>(ma, mb, p, 1, 1, stream);
    return launch_wg<128, 3>(ma, mb, p, 1, 1, stream);
}

// Form of the convolution weight gradient: 1 = three taps per CTA (tc_wgrad3_kernel, the default: the dY tile
// and one taller X box serve a whole kernel column, 47 instead of 125 B/clk/SM of L2 -> shared-memory feed;
// 5-30 % faster per layer and 2.4 % per iteration), 0 = one tap per CTA (tc_wgrad_kernel, 128 TMEM columns).
// SCDA_WGRAD3=0 selects the one-tap form at load.
static int g_wgrad3 = -1;
static bool wgrad_three_taps()
{
    if (g_wgrad3 < 0) {
        const char *e = getenv("SCDA_WGRAD3");
        g_wgrad3 = (e && *e == '0') ? 0 : 1;
    }
    return g_wgrad3 == 1;
}

SCDA_API int scda_conv3x3_wgrad_set_form(int three_taps)
{
    if (three_taps != 0 && three_taps != 1) return 0;
    g_wgrad3 = three_taps;
    return 1;
}

SCDA_API int scda_conv3x3_wgrad_bf16_nhwc_ld(int NB, int H, int W, int Cin, int Cout, const void *x, long long ldx,
                                             const void *dy, long long ldy, float *dw_partials, int splits,
                                             cudaStream_t stream)
{
    // dw_partials: [splits][Cout][3][3][Cin] fp32; the caller sums the slabs.  ldx / ldy: elements between
    // consecutive pixels of x / dy (>= Cin / Cout: the operands may be channel sub-blocks of wider tensors,
    // which is how the fp32-parity mode addresses the hi / lo halves of a split tensor, x3_ops.cu)
    if (NB <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !x || !dy || !dw_partials || splits < 1) return 0;
    if (Cin % 64 || Cout % 32) return 0;      // Cout = 32: the upper half of the 64-wide dY box is TMA zero fill
    if (ldx < Cin || ldy < Cout || ldx % 8 || ldy % 8 || ((uintptr_t)x | (uintptr_t)dy) % 16) return 0;
    int TW = 16, TH = 8;
    if (W % 16) {
        if (W % 8 == 0) { TW = 8; TH = 16; } else return 0;
    }
    if (H % TH) return 0;
    const int tiles_w = W / TW, tiles_h = H / TH;
    const int bn = Cin <= 64 ? 64 : 128;
    CUtensorMap ma, mb;
    cuuint64_t da[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)W * ldy * 2, (cuuint64_t)H * W * ldy * 2};
    cuuint64_t db[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t sb[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    if (!make_map(&ma, dy, 4, da, sa, box) || !make_map(&mb, x, 4, db, sb, box)) return 0;
    WgParams p = {};
    p.Mo = Cout; p.No = Cin; p.conv = 1;
    p.H = H; p.W = W; p.TH = TH; p.TW = TW; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.total_k_blocks = NB * tiles_h * tiles_w;
    if (splits > p.total_k_blocks) return 0;
    p.k_per_split = ceil_div(p.total_k_blocks, splits);
    if (ceil_div(p.total_k_blocks, p.k_per_split) != splits) return 0;   // every slab must be written
    p.out = dw_partials; p.ldo = 9ll * Cin; p.split_stride = (long long)Cout * 9 * Cin;
    // one unsplit pass over a short reduction with enough (tap, tile) CTAs to fill the GPU: the one-tap form
    // (the host planner asks for splits == 1 only there, scda_b200/tc.py: _wgrad_splits)
    const bool short_k = splits == 1 && p.total_k_blocks <= 16 &&
                         9 * ceil_div(Cout, kBlockM) * ceil_div(Cin, bn) >= 96;
    if (wgrad_three_taps() && !short_k) {
        // the X box is two image rows taller: one box per 64-channel chunk serves taps r = 0, 1, 2
        cuuint32_t box3[4] = {64, (cuuint32_t)TW, (cuuint32_t)(TH + 2), 1};
        CUtensorMap mb3;
        if (!make_map(&mb3, x, 4, db, sb, box3)) return 0;
        if (bn == 64) return launch_wg3<64, 4>(ma, mb3, p, splits, stream);
        return launch_wg3<128, 3>(ma, mb3, p, splits, stream);
    }
    if (bn == 64) return launch_wg<64, 4>(ma, mb, p, 9, splits, stream);
    return launch_wg<128, 3>(ma, mb, p, 9, splits, stream);
}

SCDA_API int scda_conv3x3_wgrad_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *x, const void *dy,
                                          float *dw_partials, int splits, cudaStream_t stream)
{
    return scda_conv3x3_wgrad_bf16_nhwc_ld(NB, H, W, Cin, Cout, x, Cin, dy, Cout, dw_partials, splits, stream);
}

// Weight gradient of the stride-2 3x3 convolution (conv_halo.cu, kS2): x bf16 [NB, 2Ho, 2Wo, C], dy bf16
// [NB, Ho, Wo, Cout] -> dw_partials fp32 [splits][Cout][9][4C] in the s2d weight layout; only the four taps
// {0, 1, 3, 4} are written (scda_conv_s2_wgrad_gather of disc_ops.cu folds them back to [Cout][3][3][C]).
SCDA_API int scda_conv3x3_s2_wgrad_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *x, const void *dy,
                                             float *dw_partials, int splits, cudaStream_t stream)
{
    if (NB <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || Cout <= 0 || !x || !dy || !dw_partials || splits < 1) return 0;
    if (C % 32 || Cout % 32) return 0;
    const int c2 = 2 * C, c4 = 4 * C;
    int TW = 16, TH = 8;
    if (Wo % 16) {
        if (Wo % 8 == 0) { TW = 8; TH = 16; } else return 0;
    }
    // (a tile taller than the map is fine: TMA zero-fills the rows outside on BOTH operands)
    const int tiles_w = Wo / TW, tiles_h = ceil_div(Ho, TH);
    const int bn = c2 % 128 == 0 ? 128 : 64;          // a Cin tile stays inside one row phase
    CUtensorMap ma, mb;
    cuuint64_t da[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)NB};
    cuuint64_t sa[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wo * Cout * 2, (cuuint64_t)Ho * Wo * Cout * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint64_t db[4] = {(cuuint64_t)c2, (cuuint64_t)Wo, (cuuint64_t)2 * Ho, (cuuint64_t)NB};
    cuuint64_t sb[3] = {(cuuint64_t)c2 * 2, (cuuint64_t)Wo * c2 * 2, (cuuint64_t)2 * Ho * Wo * c2 * 2};
    cuuint32_t boxb[4] = {64, (cuuint32_t)TW, (cuuint32_t)(2 * TH), 1};
    cuuint32_t es[4] = {1, 1, 2, 1};
    if (!make_map(&ma, dy, 4, da, sa, box) || !make_map_strided(&mb, x, 4, db, sb, boxb, es)) return 0;
    WgParams p = {};
    p.Mo = Cout; p.No = c4; p.conv = 1;
    p.H = Ho; p.W = Wo; p.TH = TH; p.TW = TW; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.total_k_blocks = NB * tiles_h * tiles_w;
    if (splits > p.total_k_blocks) return 0;
    p.k_per_split = ceil_div(p.total_k_blocks, splits);
    if (ceil_div(p.total_k_blocks, p.k_per_split) != splits) return 0;
    p.out = dw_partials; p.ldo = 9ll * c4; p.split_stride = (long long)Cout * 9 * c4;
    p.s2 = 1; p.c2 = c2; p.ntaps = 4;
    p.tap_list = 0ull | (1ull << 4) | (3ull << 8) | (4ull << 12);
    if (bn == 64) return launch_wg<64, 4>(ma, mb, p, 4, splits, stream);
    return launch_wg<128, 3>(ma, mb, p, 4, splits, stream);
}

// the halo convolution's stride-2 form behind the C ABI (conv_halo.cu: scda_conv_halo_s2_launch)
SCDA_API int scda_conv3x3_s2_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *x, const void *wd,
                                       const float *bias, void *y, int flags, float slope, cudaStream_t stream)
{
    if (!x || !wd || !y) return 0;
    if (flags & (kFlagMaskPos | kFlagMulSrc | kFlagAccumulate)) return 0;
    return scda_conv_halo_s2_launch(NB, Ho, Wo, C, Cout, x, wd, bias, y, flags, nullptr, slope, false, stream);
}

SCDA_API int scda_conv3x3_s2_dgrad_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *dy, const void *wd,
                                             void *dx, int flags, const void *mask_src, float slope,
                                             cudaStream_t stream)
{
    if (!dy || !wd || !dx) return 0;
    if ((flags & kFlagMaskPos) && !mask_src) return 0;
    if (flags & (kFlagRelu | kFlagLeaky | kFlagMulSrc | kFlagAccumulate)) return 0;
    return scda_conv_halo_s2_launch(NB, Ho, Wo, C, Cout, dy, wd, nullptr, dx, flags, mask_src, slope, true, stream);
}
