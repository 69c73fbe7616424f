// Probe: how does tcgen05.mma un-swizzle a K-major SWIZZLE_128B operand whose descriptor start
// address is NOT 1024-byte aligned (shifted by whole 128-byte rows), and with a stride between
// 8-row groups (SBO) that is not a multiple of 1024?  The halo form of the 3x3 convolution
// (csrc/conv_halo.cu) relies on the answer.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/probe_desc scripts/probe_desc.cu
// Shared memory holds 512 logical rows of 64 bf16 written the way TMA writes them (16-byte chunk
// c of row i at i*128 + ((c ^ (i & 7)) * 16), region 1024-aligned).  For each (row shift j, SBO,
// base_offset mode) one M128 x N64 x K64 accumulation (4 MMAs, K advance +32 B) is compared
// with the host.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t base_off)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int kRows = 512;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16 *a, const __nv_bfloat16 *b, float *d, int shift, int sbo, int base_mode)
{
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *gen = raw + (base - smem_u32(raw));
    uint8_t *sa = gen;                       // kRows x 128 B
    uint8_t *sb = gen + kRows * 128;         // 64 x 128 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(gen + kRows * 128 + 64 * 128);
    volatile uint32_t *slot = reinterpret_cast<volatile uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kRows * 8; i += 128) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sa + row * 128 + ((c ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(a + row * 64 + c * 8);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sb + row * 128 + ((c ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(b + row * 64 + c * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32((const void *)slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t a0 = base + shift * 128;
        const uint32_t bo = base_mode ? ((a0 >> 7) & 7) : 0;
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc(a0 + k * 32, sbo, bo);
            const uint64_t bd = make_desc(base + kRows * 128 + k * 32, 1024, 0);
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd),
                "r"(make_idesc(128, 64)), "r"((uint32_t)(k != 0))
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(
            smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) d[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

int main()
{
    std::vector<__nv_bfloat16> ha(kRows * 64), hb(64 * 64);
    std::vector<float> fa(kRows * 64), fb(64 * 64);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) {
        float v = (float)(rand() % 17 - 8) / 8.f;
        ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]);
    }
    for (size_t i = 0; i < hb.size(); ++i) {
        float v = (float)(rand() % 13 - 6) / 4.f;
        hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]);
    }
    __nv_bfloat16 *da, *db;
    float *dd;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    const int smem = kRows * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int sbos[] = {1024, 1280, 2304, 2048};
    std::vector<float> hd(128 * 64);
    for (int sbo : sbos)
        for (int mode = 0; mode < 2; ++mode) {
            printf("sbo %4d base_offset=%s :", sbo, mode ? "(addr>>7)&7" : "0");
            for (int shift = 0; shift < 20; ++shift) {
                cudaMemset(dd, 0, 128 * 64 * 4);
                probe_kernel<<<1, 128, smem>>>(da, db, dd, shift, sbo, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hd.data(), dd, 128 * 64 * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m) {
                    const int row = shift + (m / 8) * (sbo / 128) + (m % 8);
                    for (int n = 0; n < 64; ++n) {
                        float ref = 0.f;
                        for (int k = 0; k < 64; ++k) ref += fa[row * 64 + k] * fb[n * 64 + k];
                        if (fabsf(ref - hd[m * 64 + n]) > 1e-3f) ++bad;
                    }
                }
                printf(" %s", bad ? "x" : "ok");
            }
            printf("\n");
        }
    return 0;
}
