// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128 per CTA) for cta_group 1/2, A from shared memory (SS) or
// tensor memory (TS), N = 64/128/256, with no other shared-memory traffic.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}

template <int CG, bool TS, int BN>
__global__ void __launch_bounds__(128, 1) k_rate(long long *out, int reps, int rnd, int mode)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + 32768;            // A: [128][64] x2 ; B: up to [256][64] x2
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ volatile int stop_flag;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (32768 + 131072) / 4; i += 128) {
        // rnd != 0: pseudo-random fp16 pairs in (-1, 1) (realistic switching activity); else constant 1.0h
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const uint32_t lo = 0x3000u | (h & 0x8BFFu), hi = 0x3000u | ((h >> 16) & 0x8BFFu);
        ((uint32_t *)smem)[i] = rnd ? (lo | (hi << 16)) : 0x3c003c00u;
    }
    if (threadIdx.x == 0) { stop_flag = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3))); asm volatile("fence.mbarrier_init.release.cluster;"); asm volatile("fence.proxy.async.shared::cta;"); }
    if (warp == 0) {
        if (CG == 2) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
        else { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); } else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    constexpr int BNL = BN / CG;
    const bool dual = (mode & 64) != 0;
    if ((warp == 1 || (dual && warp == 2)) && rank == 0) {
        const uint32_t dacc = tmem + (warp == 2 ? 128u : 0u);
        uint64_t *mybar = warp == 2 ? &bar3 : &bar;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        long long t0 = 0, t1 = 0;
        if (lane == 0) {
            t0 = clock64();
            for (int r = 0; r < reps; r++) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int stg = (mode & 1) ? (r & 3) : 0;
                    const uint64_t ad = make_sdesc(smem_u32(sA) + (k >> 2) * (128 * 128) + (k & 3) * 32);
                    const uint64_t bd = make_sdesc(smem_u32(sB) + stg * 16384 + (k >> 2) * (BNL * 128) + (k & 3) * 32);
                    const uint32_t at = tmem + 256 + stg * 64 + k * 8;
                    if (CG == 2) {
                        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dacc), "r"(at), "l"(bd), "r"(idesc) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dacc), "l"(ad), "l"(bd), "r"(idesc) : "memory");
                    } else {
                        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dacc), "r"(at), "l"(bd), "r"(idesc) : "memory");
                        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dacc), "l"(ad), "l"(bd), "r"(idesc) : "memory");
                    }
                }
                if (mode & 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                if (mode & 8) {
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                }
                if (mode & 16) {
                    uint32_t ok, acc = 0;
#pragma unroll
                    for (int q = 0; q < 6; q++) { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(q & 1) : "memory"); acc += ok; }
                    if (acc == 12345) out[1] = acc;
                }
                if (mode & 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            }
            if (false) {}
            if (CG == 2) asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 1;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(smem_u32(mybar)) : "memory");
            else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mybar)) : "memory");
        }
        __syncwarp();
        mbar_wait(smem_u32(mybar), 0);
        if (lane == 0) { t1 = clock64(); out[blockIdx.x * 2 + (warp == 2)] = t1 - t0; }
    }
    if ((mode & 4) && (warp == 2 || warp == 3)) {
        // hammer TMEM with stores to columns 256+ of this warp's lane quadrant while the MMAs run
        uint32_t v = threadIdx.x;
        while (!stop_flag) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + 256 + 64), "r"(v) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            v++;
        }
    }
    if (warp == 1 && rank == 0 && lane == 0) stop_flag = 1;   // (dual mode is only used without the store hammer)
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); } else __syncthreads();
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem)); else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

template <int CG, bool TS, int BN>
void run(const char *name, int nblocks, int rnd, int mode)
{
    long long *d; cudaMalloc(&d, 8 * 512); cudaMemset(d, 0, 8 * 512);
    const int reps = 512;
    size_t smem = 1024 + 32768 + 131072;
    cudaFuncSetAttribute(k_rate<CG, TS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int it = 0; it < 2; it++) cudaLaunchKernelEx(&cfg, k_rate<CG, TS, BN>, d, reps, rnd, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[512]; cudaMemcpy(h, d, 8 * 512, cudaMemcpyDeviceToHost);
    printf("%-14s mode=%d rnd=%d blocks=%3d : %s cycles/MMA = %.1f   (MMA = 128x%dx16 per CTA%s)\n", name, mode, rnd, nblocks, cudaGetErrorString(e), (double)(h[0] > h[1] ? h[0] : h[1]) / (reps * 8 * ((mode & 64) ? 2 : 1)), BN, CG == 2 ? ", pair 256 rows" : "");
    cudaFree(d);
}

int main()
{
    for (int mode : {0, 57, 64, 64 + 57}) {
        run<1, true, 128>("cg1 TS N=128", 128, 1, mode);
        run<2, true, 128>("cg2 TS N=128", 128, 1, mode);
    }
    return 0;
}
