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

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, bool cg2)
{
    if (cg2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
// bytes one K step (128 weights) of one weight row occupies in its reference layout = the TMA box width of the raw ring:
// Q4_0 4 x 20, Q4_2 8 x 10; Q4_1 / Q5_1 4 x 24; Q8_0 4 x 36; Q5_0 4 x 22
// Q5_0's 4 x 22 = 88 bytes are neither a legal box width nor a legal box origin for odd K steps (both multiples of 16), so its
// box is 96 bytes wide and starts at 88 ks rounded down to 16: the step's bytes begin at offset 0 (even ks) or 8 (odd ks) of the
// shared-memory row -- read with LDS.64 -- and the 8 spare bytes belong to a neighbouring K step (or are TMA's out-of-bounds zeros).
template <int TYPE> struct RawRow {
    static constexpr int BYTES = (TYPE == GGML_TYPE_Q4_0 || TYPE == GGML_TYPE_Q4_2) ? 80 : TYPE == GGML_TYPE_Q8_0 ? 144 : 96;   // box width = shared-memory row
    __device__ static __forceinline__ int box_x(int ks) { return TYPE == GGML_TYPE_Q5_0 ? (88 * ks) & ~15 : BYTES * ks; }        // byte coordinate of the box
    __device__ static __forceinline__ int skew(int ks) { return TYPE == GGML_TYPE_Q5_0 ? (ks & 1) << 3 : 0; }                    // where the step starts inside the row
};
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
// this thread's raw words of one K step: LDS.128 over the whole row, or (Q5_0) eleven LDS.64 from the skewed start
template <int TYPE>
__device__ __forceinline__ void load_raw_row(uint32_t addr, int ks, uint32_t *w)
{
    if (TYPE == GGML_TYPE_Q5_0) {
        const uint32_t a = addr + (uint32_t)RawRow<TYPE>::skew(ks);
#pragma unroll
        for (int i = 0; i < 11; i++) { const uint2 t = lds64(a + i * 8); w[2 * i] = t.x; w[2 * i + 1] = t.y; }
    } else {
#pragma unroll
        for (int i = 0; i < RawRow<TYPE>::BYTES / 16; i++) { const uint4 t = lds128(addr + i * 16); w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w; }
    }
}

// 8 weights of one 32-bit nibble word -> 4 half2 in K order (0,4) (1,5) (2,6) (3,7), scaled: (q-c)*d [+ m'], c = 8 (16 for Q5_1).
// Q5_1: hb = the 8 fifth bits of these weights (qh >> 8*word); bit e belongs to element e and lands on the 16s place of its half:
// bit 4 / 20 in the "1024 + n" halves (elements 0,4 / 2,6), bit 8 / 24 in the "64 + n" halves whose ulp is 1/16 (elements 1,5 / 3,7).
// Each pair is moved by one multiply ((hb & 0x11) * (2^4 + 2^16) etc.: the cross terms fall outside the mask).
template <int TYPE>
__device__ __forceinline__ void dequant_word(uint32_t w, uint32_t hb, __half2 d2, __half2 m2, uint32_t mk_lo, uint32_t mk_hi, uint32_t mg_lo, uint32_t mg_hi, uint32_t *out)
{
    constexpr float C = (TYPE == GGML_TYPE_Q5_1 || TYPE == GGML_TYPE_Q5_0) ? 16.0f : 8.0f;
    const __half2 o_lo = __float2half2_rn(1024.0f + C), o_hi = __float2half2_rn(64.0f + C);
    const uint32_t ws = w >> 8;
    uint32_t v0 = and_or(w, mk_lo, mg_lo), v1 = and_or(w, mk_hi, mg_hi), v2 = and_or(ws, mk_lo, mg_lo), v3 = and_or(ws, mk_hi, mg_hi);
    if (TYPE == GGML_TYPE_Q5_1 || TYPE == GGML_TYPE_Q5_0) {
        v0 |= ((hb & 0x11u) * 0x00010010u) & 0x00100010u;
        v1 |= ((hb & 0x22u) * 0x00080080u) & 0x01000100u;
        v2 |= ((hb & 0x44u) * 0x00004004u) & 0x00100010u;
        v3 |= ((hb & 0x88u) * 0x00020020u) & 0x01000100u;
    }
    __half2 h0 = *reinterpret_cast<__half2 *>(&v0), h1 = *reinterpret_cast<__half2 *>(&v1);
    __half2 h2 = *reinterpret_cast<__half2 *>(&v2), h3 = *reinterpret_cast<__half2 *>(&v3);
    if (TYPE == GGML_TYPE_Q4_0 || TYPE == GGML_TYPE_Q4_2 || TYPE == GGML_TYPE_Q5_0) {
        h0 = __hmul2(__hsub2(h0, o_lo), d2); h1 = __hmul2(__hsub2(h1, o_hi), d2);
        h2 = __hmul2(__hsub2(h2, o_lo), d2); h3 = __hmul2(__hsub2(h3, o_hi), d2);
    } else {
        h0 = __hfma2(__hsub2(h0, o_lo), d2, m2); h1 = __hfma2(__hsub2(h1, o_hi), d2, m2);
        h2 = __hfma2(__hsub2(h2, o_lo), d2, m2); h3 = __hfma2(__hsub2(h3, o_hi), d2, m2);
    }
    out[0] = *reinterpret_cast<uint32_t *>(&h0); out[1] = *reinterpret_cast<uint32_t *>(&h1);
    out[2] = *reinterpret_cast<uint32_t *>(&h2); out[3] = *reinterpret_cast<uint32_t *>(&h3);
}

// Group j (32 weights) of the raw words w[] of one K step of one row -> 16 words of fp16 pairs.
//   Q4_0  words 5j: [f32 d][qs x 4]                          (q-8)*d
//   Q4_1  words 6j: [f32 d][f32 m][qs x 4]                   (q-8)*d + (m + 8d)      (recentred so the product stays small)
//   Q4_2  words 5j: [h d0 | qs0 0-1][qs0 2-5][qs0 6-7 | h d1][qs1 0-3][qs1 4-7]      two 16-weight blocks, fp16 scales used as they are
//   Q5_1  words 6j: [h d | h m][qh][qs x 4]                  (q5-16)*d + (m + 16d)
// rs = 2^-ew of this weight row (ggb_internal.h: launch_weight_rowexp): the block scale -- float32 in Q4_0 / Q4_1 / Q8_0, fp16 in the
// others -- is multiplied by it in float BEFORE it is rounded to the fp16 MMA operand, so a row of tiny or huge scales keeps its
// mantissa (the reference multiplies d0 * d1 * sumi in float32, Ggml.cs:1158) and (q - c) * d never overflows fp16.
template <int TYPE>
__device__ __forceinline__ void dequant_group(const uint32_t *w, int j, float rs, uint32_t mk_lo, uint32_t mk_hi, uint32_t mg_lo, uint32_t mg_hi, uint32_t *out)
{
    if (TYPE == GGML_TYPE_Q4_0 || TYPE == GGML_TYPE_Q4_1) {
        const uint32_t *wb = TYPE == GGML_TYPE_Q4_0 ? &w[5 * j] : &w[6 * j];
        const __half2 d2 = __float2half2_rn(__uint_as_float(wb[0]) * rs);
        __half2 m2 = __float2half2_rn(0.0f);
        if (TYPE == GGML_TYPE_Q4_1) m2 = __float2half2_rn(fmaf(8.0f, __uint_as_float(wb[0]), __uint_as_float(wb[1])) * rs);
        const uint32_t *qw = TYPE == GGML_TYPE_Q4_0 ? wb + 1 : wb + 2;
#pragma unroll
        for (int i = 0; i < 4; i++) dequant_word<TYPE>(qw[i], 0u, d2, m2, mk_lo, mk_hi, mg_lo, mg_hi, out + 4 * i);
    } else if (TYPE == GGML_TYPE_Q4_2) {
        const uint32_t *wb = &w[5 * j];
        const __half2 d2a = __float2half2_rn(__half2float(__ushort_as_half((unsigned short)(wb[0] & 0xFFFFu))) * rs);
        const __half2 d2b = __float2half2_rn(__half2float(__ushort_as_half((unsigned short)(wb[2] >> 16))) * rs), z = __float2half2_rn(0.0f);
        dequant_word<TYPE>(__funnelshift_r(wb[0], wb[1], 16), 0u, d2a, z, mk_lo, mk_hi, mg_lo, mg_hi, out);
        dequant_word<TYPE>(__funnelshift_r(wb[1], wb[2], 16), 0u, d2a, z, mk_lo, mk_hi, mg_lo, mg_hi, out + 4);
        dequant_word<TYPE>(wb[3], 0u, d2b, z, mk_lo, mk_hi, mg_lo, mg_hi, out + 8);
        dequant_word<TYPE>(wb[4], 0u, d2b, z, mk_lo, mk_hi, mg_lo, mg_hi, out + 12);
    } else if (TYPE == GGML_TYPE_Q5_0) {
        // 22-byte blocks: group j starts at byte 22j = word 5.5j -- even j word aligned [h d | qh lo][qh hi | qs 0-1][qs ..] ...,
        // odd j in the upper half of a word [.. | h d][qh][qs x 4] (the same two decodes as the GEMV, ggb_sib_math.cuh)
        const uint32_t *wb = &w[(11 * j) >> 1];
        uint32_t dd, qh, q[4];
        if (j & 1) {
            dd = wb[0] >> 16; qh = wb[1];
#pragma unroll
            for (int i = 0; i < 4; i++) q[i] = wb[2 + i];
        } else {
            dd = wb[0] & 0xFFFFu; qh = __funnelshift_r(wb[0], wb[1], 16);
#pragma unroll
            for (int i = 0; i < 4; i++) q[i] = __funnelshift_r(wb[1 + i], wb[2 + i], 16);
        }
        const __half2 d2 = __float2half2_rn(__half2float(__ushort_as_half((unsigned short)dd)) * rs), z = __float2half2_rn(0.0f);
#pragma unroll
        for (int i = 0; i < 4; i++) dequant_word<TYPE>(q[i], (qh >> (8 * i)) & 0xFFu, d2, z, mk_lo, mk_hi, mg_lo, mg_hi, out + 4 * i);
    } else if (TYPE == GGML_TYPE_Q8_0) {
        // words 9j: [f32 d][32 x int8].  Bytes k of two consecutive words are elements k and k+4 of a group of 8: PRMT doubles them
        // into the two halves and ONE LOP3 (immLut 0x6A = b ? a ^ c : c, b = 0x00FF00FF, c = 0x64806480) masks, flips the sign
        // bit and ORs the fp16 magic in: half = 1024 + (q + 128), exact; minus 1152, times d.
        const uint32_t *wb = &w[9 * j];
        const __half2 d2 = __float2half2_rn(__uint_as_float(wb[0]) * rs), o = __float2half2_rn(1152.0f);
        uint32_t mk = 0x00FF00FFu, mg = 0x64806480u;
        asm volatile("" : "+r"(mk), "+r"(mg));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t wa = wb[1 + 2 * i], wc = wb[2 + 2 * i];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t t = __byte_perm(wa, wc, (uint32_t)(k | (k << 4) | ((4 + k) << 8) | ((4 + k) << 12))), v;
                asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(v) : "r"(t), "r"(mk), "r"(mg));
                const __half2 h = __hmul2(__hsub2(*reinterpret_cast<__half2 *>(&v), o), d2);
                out[4 * i + k] = *reinterpret_cast<const uint32_t *>(&h);
            }
        }
    } else {
        const uint32_t *wb = &w[6 * j];
        const float df = __half2float(__ushort_as_half((unsigned short)(wb[0] & 0xFFFFu))), mf = __half2float(__ushort_as_half((unsigned short)(wb[0] >> 16)));
        const __half2 d2 = __float2half2_rn(df * rs);
        const __half2 m2 = __float2half2_rn(fmaf(16.0f, df, mf) * rs);
        const uint32_t qh = wb[1];
#pragma unroll
        for (int i = 0; i < 4; i++) dequant_word<TYPE>(wb[2 + i], (qh >> (8 * i)) & 0xFFu, d2, m2, mk_lo, mk_hi, mg_lo, mg_hi, out + 4 * i);
    }
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
