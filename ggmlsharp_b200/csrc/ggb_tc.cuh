// tcgen05 / TMA / mbarrier building blocks shared by the batched mul_mat kernels (ggb_gemm.cu: one tile per CTA pair,
// ggb_gemm_grouped.cu: persistent CTA pairs walking the tiles of a whole batch of nodes).  Inline PTX for sm_100a only.
#pragma once
#include "ggb_internal.h"

#include <cuda.h>

namespace ggb {
namespace tc {

constexpr int BM = 128;            // weight rows per tile  (MMA M, TMEM lanes)
constexpr int BK = 128;            // K per pipeline step   (4 quant blocks; two 64-wide swizzle atoms)
constexpr int NDQ_WARPS = 16;
constexpr int NTHREADS = (4 + NDQ_WARPS) * 32;


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)      // non-blocking probe
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row (1024 B) swizzle atoms stacked along M/N.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)0 << 16;                            // leading byte offset: unused for a single swizzled K atom
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier that may live in the peer CTA (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// 2-CTA TMA load: bytes land in this CTA's shared memory, completion is signalled on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar)      // arrives on the barrier at this offset in BOTH CTAs of the pair
{
    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// One 32-weight block -> 64 bytes of the swizzle-128B K-major A tile.  `row` points at this thread's 128-byte row of its
// sub-tile, `cx` is the row's swizzle XOR (r & 7), `half` selects the 32-K half of the 64-wide sub-tile, `rot` the STS order.
// (w & mask) | magic in ONE LOP3 (immLut 0xEA = (a & b) | c); the constants are passed in registers so ptxas cannot split it
__device__ __forceinline__ uint32_t and_or(uint32_t w, uint32_t mask, uint32_t magic)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "r"(mask), "r"(magic));
    return d;
}


// ---- A operand in tensor memory: nibble -> fp16 expansion and tcgen05.st (used by the Q4_0 / Q4_1 kernels) ----
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, bool cg2)
{
    if (cg2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// 8 weights of one 32-bit nibble word -> 4 half2 in K order (0,4) (1,5) (2,6) (3,7), scaled: (q-8)*d [+ m']
template <int TYPE>
__device__ __forceinline__ void dequant_word(uint32_t w, __half2 d2, __half2 m2, uint32_t mk_lo, uint32_t mk_hi, uint32_t mg_lo, uint32_t mg_hi, uint32_t *out)
{
    const __half2 o_lo = __float2half2_rn(1032.0f), o_hi = __float2half2_rn(72.0f);     // 1024+8, 64+8
    const uint32_t ws = w >> 8;
    uint32_t v0 = and_or(w, mk_lo, mg_lo), v1 = and_or(w, mk_hi, mg_hi), v2 = and_or(ws, mk_lo, mg_lo), v3 = and_or(ws, mk_hi, mg_hi);
    __half2 h0 = *reinterpret_cast<__half2 *>(&v0), h1 = *reinterpret_cast<__half2 *>(&v1);
    __half2 h2 = *reinterpret_cast<__half2 *>(&v2), h3 = *reinterpret_cast<__half2 *>(&v3);
    if (TYPE == GGML_TYPE_Q4_0) {
        h0 = __hmul2(__hsub2(h0, o_lo), d2); h1 = __hmul2(__hsub2(h1, o_hi), d2);
        h2 = __hmul2(__hsub2(h2, o_lo), d2); h3 = __hmul2(__hsub2(h3, o_hi), d2);
    } else {
        h0 = __hfma2(__hsub2(h0, o_lo), d2, m2); h1 = __hfma2(__hsub2(h1, o_hi), d2, m2);
        h2 = __hfma2(__hsub2(h2, o_lo), d2, m2); h3 = __hfma2(__hsub2(h3, o_hi), d2, m2);
    }
    out[0] = *reinterpret_cast<uint32_t *>(&h0); out[1] = *reinterpret_cast<uint32_t *>(&h1);
    out[2] = *reinterpret_cast<uint32_t *>(&h2); out[3] = *reinterpret_cast<uint32_t *>(&h3);
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}


} // namespace tc
} // namespace ggb
