// Single-token (N < 16) mul_mat: a persistent, bandwidth-bound GEMV over a LIST of graph nodes.
//
// Replaces the COMPUTE phase of ggml_compute_forward_mul_mat_{f32,f16_f32,q_f32} (Ggml.cs:6127-6164,
// 6390-6425, 6662-6699) and the dots ggml_vec_dot_{f32,f16,q4_0_q8_0,q4_1_q8_1} (Ggml.cs:2631-2651,
// 1124-1201).  The reference splits src0 rows over OS threads (dr = ceil(nr/nth)); here rows are split
// over every warp of a one-CTA-per-SM grid, and one launch walks all nodes of a batch.
//
// Data movement: weight rows are contiguous byte ranges (nb01 bytes each), so a producer warp streams
// tiles of rows HBM -> shared memory with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a ring
// of stages guarded by full/empty mbarriers (k_gemv_fast); rows that are not 16-byte multiples take
// k_gemv, which stages with plain loads.  Lanes read 20-byte Q4_0 / 24-byte Q4_1 blocks from shared
// memory (80 / 48-byte units with LDS.128, or single blocks with a lane stride of 5 / 6 words: both
// bank-conflict free), so the raw reference block layout is kept in HBM and the unaligned-vector-load
// problem of 20-byte blocks never reaches the memory system.
// The activation vector(s) were quantized by k_act_batch exactly as quantize_row_q8_0/q8_1 do and sit
// in shared memory for the whole node.
//
// Arithmetic per Q4_0 block is the reference's: an exact int32 block sum (dp4a), then
// sumf += (d0*d1) * (float)sumi in float.  Only the cross-block summation order differs (lane-strided
// partial sums + a shuffle tree instead of one sequential chain).
//
// The sibling formats Q4_2 / Q5_0 / Q5_1 / Q8_0 (ggml_vec_dot_q4_2_q8_0 ... q8_0_q8_0, Ggml.cs:1204-1380) run through
// the same two kernels: their per-block arithmetic lives in ggb_sib_math.cuh, only the unit geometry differs (Q4_2: 8 blocks =
// 80 B, Q5_0: 4 blocks = 88 B read with LDS.64, Q5_1: 2 blocks = 48 B, Q8_0: 4 blocks = 144 B; all lane strides conflict-free).
//
// Roofline: HBM.  Algorithmic bytes per node = M*row_bytes (weights) + 4*K*N (x) + 4*M*N (y).
#include "ggb_internal.h"
#include "ggb_sib_math.cuh"
#include "ggb_act_q8.cuh"

#ifndef GGB_GEMV_PART
#define GGB_GEMV_PART 0
#endif

#include <algorithm>
#include <cstdlib>

namespace ggb {

namespace {

constexpr int NWARPS = 8;
constexpr int MAX_DEPTH = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ---- per-type partial dot of one staged chunk against NC staged activation columns ----

template <int NC>
__device__ __forceinline__ void dot_q4_0(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes, int kb,
                                         int lane, float (&acc)[NC])
{
    const int nblk = nbytes / 20, b0 = off_bytes / 20;
    for (int i = lane; i < nblk; i += 32) {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(w + i * 20);
        const float d0 = __uint_as_float(wp[0]);
        int lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const uint32_t q = wp[1 + j]; lo[j] = (int)(q & 0x0F0F0F0Fu); hi[j] = (int)((q >> 4) & 0x0F0F0F0Fu); }
        const int bi = b0 + i;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint8_t *xc = xs + c * xcol_bytes;
            const int4 ev = *reinterpret_cast<const int4 *>(xc + bi * 16);
            const int4 od = *reinterpret_cast<const int4 *>(xc + kb * 16 + bi * 16);
            const int2 ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + bi * 8);
            int s = ds.y * -8;                                   // sum (q-8)*p = sum q*p - 8*sum p
            s = __dp4a(lo[0], ev.x, s); s = __dp4a(hi[0], od.x, s);
            s = __dp4a(lo[1], ev.y, s); s = __dp4a(hi[1], od.y, s);
            s = __dp4a(lo[2], ev.z, s); s = __dp4a(hi[2], od.z, s);
            s = __dp4a(lo[3], ev.w, s); s = __dp4a(hi[3], od.w, s);
            acc[c] = __fadd_rn(acc[c], __fmul_rn(__fmul_rn(d0, __int_as_float(ds.x)), (float)s));   // Ggml.cs:1158
        }
    }
}

template <int NC>
__device__ __forceinline__ void dot_q4_1(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes, int kb,
                                         int lane, float (&acc)[NC])
{
    const int nblk = nbytes / 24, b0 = off_bytes / 24;
    for (int i = lane; i < nblk; i += 32) {
        const uint2 *wp = reinterpret_cast<const uint2 *>(w + i * 24);
        const uint2 dm = wp[0], qa = wp[1], qb = wp[2];
        const float d0 = __uint_as_float(dm.x), m0 = __uint_as_float(dm.y);
        const uint32_t q[4] = {qa.x, qa.y, qb.x, qb.y};
        int lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { lo[j] = (int)(q[j] & 0x0F0F0F0Fu); hi[j] = (int)((q[j] >> 4) & 0x0F0F0F0Fu); }
        const int bi = b0 + i;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint8_t *xc = xs + c * xcol_bytes;
            const int4 ev = *reinterpret_cast<const int4 *>(xc + bi * 16);
            const int4 od = *reinterpret_cast<const int4 *>(xc + kb * 16 + bi * 16);
            const int2 ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + bi * 8);
            int s = 0;
            s = __dp4a(lo[0], ev.x, s); s = __dp4a(hi[0], od.x, s);
            s = __dp4a(lo[1], ev.y, s); s = __dp4a(hi[1], od.y, s);
            s = __dp4a(lo[2], ev.z, s); s = __dp4a(hi[2], od.z, s);
            s = __dp4a(lo[3], ev.w, s); s = __dp4a(hi[3], od.w, s);
            // sum_j (d0*q_j + m0) * (d1*p_j)  (Ggml.cs:1190-1196)  =  d0*d1*sum q_j p_j  +  m0*d1*sum p_j
            const float d1 = __int_as_float(ds.x);
            acc[c] += (d0 * d1) * (float)s + (m0 * d1) * (float)ds.y;
        }
    }
}

template <int NC>
__device__ __forceinline__ void dot_f16(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes,
                                        int lane, float (&acc)[NC])
{
    const int nvec = nbytes >> 4, v0 = off_bytes >> 4;
    for (int i = lane; i < nvec; i += 32) {
        const uint4 wv = *reinterpret_cast<const uint4 *>(w + i * 16);
        const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
        float2 wf[4];
#pragma unroll
        for (int j = 0; j < 4; j++) wf[j] = __half22float2(wh[j]);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint4 xv = *reinterpret_cast<const uint4 *>(xs + c * xcol_bytes + (v0 + i) * 16);
            const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 xf = __half22float2(xh[j]);
                acc[c] = fmaf(wf[j].x, xf.x, acc[c]);            // the f16*f16 product is exact in f32 (Ggml.cs:2647)
                acc[c] = fmaf(wf[j].y, xf.y, acc[c]);
            }
        }
    }
    const int tail = (nbytes & 15) >> 1;                         // K % 8 elements
    if (lane < tail) {
        const int e = (nvec << 3) + lane;
        const float wv = __half2float(reinterpret_cast<const __half *>(w)[e]);
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] = fmaf(wv, __half2float(reinterpret_cast<const __half *>(xs + c * xcol_bytes + off_bytes)[e]), acc[c]);
    }
}

template <int NC>
__device__ __forceinline__ void dot_f32(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes,
                                        int lane, float (&acc)[NC])
{
    const int nvec = nbytes >> 4, v0 = off_bytes >> 4;
    for (int i = lane; i < nvec; i += 32) {
        const float4 wv = *reinterpret_cast<const float4 *>(w + i * 16);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + c * xcol_bytes + (v0 + i) * 16);
            // float product, then accumulate (Ggml.cs:2636); the accumulator is f32 here, f64 in the reference
            acc[c] += __fmul_rn(wv.x, xv.x); acc[c] += __fmul_rn(wv.y, xv.y);
            acc[c] += __fmul_rn(wv.z, xv.z); acc[c] += __fmul_rn(wv.w, xv.w);
        }
    }
    const int tail = (nbytes & 15) >> 2;
    if (lane < tail) {
        const int e = (nvec << 2) + lane;
        const float wv = reinterpret_cast<const float *>(w)[e];
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] += __fmul_rn(wv, reinterpret_cast<const float *>(xs + c * xcol_bytes + off_bytes)[e]);
    }
}

// sibling formats on the plain-staging path: a lane owns one 32-element group (20 / 22 / 24 / 36 bytes, only 2-byte aligned in the
// stage), gathers it as 16-bit words and runs the shared per-group dot; activations in the linear (bps = 1) Q8P layout
template <int TYPE, int NC>
__device__ __forceinline__ void dot_sib(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes, int kb,
                                        int lane, float (&acc)[NC])
{
    constexpr int G = sib::Grp<TYPE>::G, NW = (G + 3) / 4;
    const int ngrp = nbytes / G, g0 = off_bytes / G;
    for (int i = lane; i < ngrp; i += 32) {
        const unsigned short *hp = reinterpret_cast<const unsigned short *>(w + i * G);
        uint32_t wd[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) wd[k] = (uint32_t)hp[2 * k] | (2 * k + 1 < G / 2 ? (uint32_t)hp[2 * k + 1] << 16 : 0u);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint8_t *xc = xs + c * xcol_bytes;
            sib::XBlk x;
            x.ev = *reinterpret_cast<const int4 *>(xc + (g0 + i) * 16);
            x.od = *reinterpret_cast<const int4 *>(xc + kb * 16 + (g0 + i) * 16);
            x.ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + (g0 + i) * 8);
            acc[c] = sib::dot_group_words<TYPE>(wd, x, acc[c]);
        }
    }
}

template <int TYPE, int NC>
__device__ __forceinline__ void dot_chunk(const GemvHdr &b, const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs,
                                          int lane, float (&acc)[NC])
{
    if constexpr (TYPE == GGML_TYPE_Q4_0) dot_q4_0<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, b.K / GGB_QK, lane, acc);
    else if constexpr (TYPE == GGML_TYPE_Q4_1) dot_q4_1<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, b.K / GGB_QK, lane, acc);
    else if constexpr (TYPE == GGML_TYPE_F16) dot_f16<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, lane, acc);
    else if constexpr (TYPE == GGML_TYPE_F32) dot_f32<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, lane, acc);
    else dot_sib<TYPE, NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, b.K / GGB_QK, lane, acc);
}

// Generic path for rows that are not 16-byte multiples / not 16-byte aligned (e.g. K = 4128, byte-offset views): the same
// math, but every warp stages its own row (or K-chunk of a long row) into shared memory with plain loads.  One group = one
// weight row; groups are dealt to warps round-robin inside a CTA's contiguous range.
template <int TYPE, int NC, int CAP>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_gemv(const __grid_constant__ GemvBatchT<CAP> b)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xbytes = (NC * b.xcol_bytes + 127) & ~127;
    uint8_t *xs = smem;
    uint8_t *stage = smem + xbytes + (size_t)warp * b.stage_bytes;

    const int g_begin = (int)((long long)b.total_groups * blockIdx.x / gridDim.x);
    const int g_end = (int)((long long)b.total_groups * (blockIdx.x + 1) / gridDim.x);

    asm volatile("griddepcontrol.wait;" ::: "memory");           // activations come from k_act_batch
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (g_begin >= g_end) return;
    int n = 0;
    while (g_begin >= b.node[n].g0 + b.node[n].ngroups) n++;
    for (; n < b.n_nodes && b.node[n].g0 < g_end; n++) {
        const GemvNode &nd = b.node[n];
        __syncthreads();                                         // everyone is done with the previous node's x
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(nd.xq);
            uint4 *dst = reinterpret_cast<uint4 *>(xs);
            for (int i = threadIdx.x; i < (NC * b.xcol_bytes) >> 4; i += NWARPS * 32) dst[i] = src[i];
        }
        __syncthreads();
        const int seg_lo = max(g_begin, nd.g0), seg_hi = min(g_end, nd.g0 + nd.ngroups);
        int g = seg_lo + ((warp - (seg_lo - g_begin) % NWARPS + NWARPS) % NWARPS);
        for (; g < seg_hi; g += NWARPS) {
            const int row = g - nd.g0;
            float acc[NC];
#pragma unroll
            for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
            for (int c = 0; c < b.nchunk; c++) {
                const int nbytes = min(b.chunk_bytes, b.row_bytes - c * b.chunk_bytes);
                const uint8_t *src = nd.W + (long long)row * b.nb01 + (long long)c * b.chunk_bytes;
                __syncwarp();                                    // previous chunk fully consumed
                if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 3) == 0)
                    for (int i = lane; i < nbytes >> 2; i += 32) reinterpret_cast<uint32_t *>(stage)[i] = reinterpret_cast<const uint32_t *>(src)[i];
                else
                    for (int i = lane; i < nbytes >> 1; i += 32) reinterpret_cast<uint16_t *>(stage)[i] = reinterpret_cast<const uint16_t *>(src)[i];
                __syncwarp();
                dot_chunk<TYPE, NC>(b, stage, c * b.chunk_bytes, nbytes, xs, lane, acc);
            }
#pragma unroll
            for (int cc = 0; cc < NC; cc++) {
                const float v = warp_sum(acc[cc]);
                if (lane == 0) {
                    float *yp = nd.y + (long long)cc * nd.ldy + row;
                    *yp = v;
                    for (int p = 0; p < b.n_peers; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + b.peer_delta[p]) = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path: rows are whole 16-byte-aligned "units" (Q4_0: 4 blocks = 80 B, Q4_1: 2 blocks = 48 B,
// F16/F32: 16 B), so a lane reads a unit with LDS.128 (lane strides of 20 / 12 / 4 words are
// bank-conflict free) and every bulk copy is legal.  16 warps per CTA, stage/phase counters instead
// of div/mod, and -- for single-token Q4 nodes with K <= 4096 -- the lane's slice of the quantized
// activation vector lives in registers for the whole node.
// Activation layout for Q4 types here is "unit-major": block b = bps*u + j is stored at plane index
// j*S + u (S = units per row), so lanes reading consecutive units hit consecutive 16-byte slots.
// ------------------------------------------------------------------------------------------------

constexpr int NWF = 15;      // consumer warps; +1 producer warp = 16 warps = 512 threads -> 128 registers per thread

using sib::XBlk;             // one Q8P activation block: even quants, odd quants, {d, sum}

__device__ __forceinline__ int q4_isum(const uint32_t q0, const uint32_t q1, const uint32_t q2, const uint32_t q3, const XBlk &x, int s)
{
    s = __dp4a((int)(q0 & 0x0F0F0F0Fu), x.ev.x, s); s = __dp4a((int)((q0 >> 4) & 0x0F0F0F0Fu), x.od.x, s);
    s = __dp4a((int)(q1 & 0x0F0F0F0Fu), x.ev.y, s); s = __dp4a((int)((q1 >> 4) & 0x0F0F0F0Fu), x.od.y, s);
    s = __dp4a((int)(q2 & 0x0F0F0F0Fu), x.ev.z, s); s = __dp4a((int)((q2 >> 4) & 0x0F0F0F0Fu), x.od.z, s);
    s = __dp4a((int)(q3 & 0x0F0F0F0Fu), x.ev.w, s); s = __dp4a((int)((q3 >> 4) & 0x0F0F0F0Fu), x.od.w, s);
    return s;
}

__device__ __forceinline__ XBlk load_xblk(const uint8_t *xc, int kb, int idx)
{
    XBlk x;
    x.ev = *reinterpret_cast<const int4 *>(xc + idx * 16);
    x.od = *reinterpret_cast<const int4 *>(xc + kb * 16 + idx * 16);
    x.ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + idx * 8);
    return x;
}

// one Q4_0 unit (4 blocks, 80 bytes at w) against 4 activation blocks
__device__ __forceinline__ float q4_0_unit(const uint8_t *w, const XBlk (&x)[4], float acc)
{
    const uint4 *wp = reinterpret_cast<const uint4 *>(w);
    const uint4 a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
    // words: d0 q q q q | d1 q q q q | d2 q q q q | d3 q q q q
    int s;
    s = q4_isum(a0.y, a0.z, a0.w, a1.x, x[0], x[0].ds.y * -8);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__uint_as_float(a0.x), __int_as_float(x[0].ds.x)), (float)s));
    s = q4_isum(a1.z, a1.w, a2.x, a2.y, x[1], x[1].ds.y * -8);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__uint_as_float(a1.y), __int_as_float(x[1].ds.x)), (float)s));
    s = q4_isum(a2.w, a3.x, a3.y, a3.z, x[2], x[2].ds.y * -8);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__uint_as_float(a2.z), __int_as_float(x[2].ds.x)), (float)s));
    s = q4_isum(a4.x, a4.y, a4.z, a4.w, x[3], x[3].ds.y * -8);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__uint_as_float(a3.w), __int_as_float(x[3].ds.x)), (float)s));
    return acc;
}

// one Q4_1 unit (2 blocks, 48 bytes at w) against 2 activation blocks
__device__ __forceinline__ float q4_1_unit(const uint8_t *w, const XBlk (&x)[4], float acc)
{
    const uint4 *wp = reinterpret_cast<const uint4 *>(w);
    const uint4 a0 = wp[0], a1 = wp[1], a2 = wp[2];
    // words: d0 m0 q q | q q d1 m1 | q q q q
    int s = q4_isum(a0.z, a0.w, a1.x, a1.y, x[0], 0);
    float d1 = __int_as_float(x[0].ds.x);
    acc += (__uint_as_float(a0.x) * d1) * (float)s + (__uint_as_float(a0.y) * d1) * (float)x[0].ds.y;
    s = q4_isum(a2.x, a2.y, a2.z, a2.w, x[1], 0);
    d1 = __int_as_float(x[1].ds.x);
    acc += (__uint_as_float(a1.z) * d1) * (float)s + (__uint_as_float(a1.w) * d1) * (float)x[1].ds.y;
    return acc;
}

// sibling formats: the unit's words are fetched with the widest aligned loads its size allows (Q5_0's 88 bytes: 11 x LDS.64,
// lane stride 22 words, conflict-free per half-warp; the others LDS.128 with lane strides of 20 / 12 / 36 words); the arithmetic
// on those words is ggb_sib_math.cuh's dot_unit_words
template <int TYPE>
__device__ __forceinline__ float sib_unit(const uint8_t *w, const XBlk (&x)[4], float acc)
{
    constexpr int NWORDS = sib::Unit<TYPE>::BYTES / 4;
    uint32_t v[NWORDS];
    if constexpr (sib::Unit<TYPE>::BYTES % 16 == 0) {
        const uint4 *wp = reinterpret_cast<const uint4 *>(w);
#pragma unroll
        for (int i = 0; i < NWORDS / 4; i++) { const uint4 a = wp[i]; v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w; }
    } else {
        const uint2 *wp = reinterpret_cast<const uint2 *>(w);
#pragma unroll
        for (int i = 0; i < NWORDS / 2; i++) { const uint2 a = wp[i]; v[2 * i] = a.x; v[2 * i + 1] = a.y; }
    }
    return sib::dot_unit_words<TYPE>(v, x, acc);
}

template <int TYPE> struct UnitTraits { static constexpr int BYTES = 16, BPS = 0; static constexpr bool Q = false; };
template <> struct UnitTraits<GGML_TYPE_Q4_0> { static constexpr int BYTES = 80, BPS = 4; static constexpr bool Q = true; };
template <> struct UnitTraits<GGML_TYPE_Q4_1> { static constexpr int BYTES = 48, BPS = 2; static constexpr bool Q = true; };
template <> struct UnitTraits<GGML_TYPE_Q4_2> { static constexpr int BYTES = 80, BPS = 4; static constexpr bool Q = true; };
template <> struct UnitTraits<GGML_TYPE_Q5_0> { static constexpr int BYTES = 88, BPS = 4; static constexpr bool Q = true; };
template <> struct UnitTraits<GGML_TYPE_Q5_1> { static constexpr int BYTES = 48, BPS = 2; static constexpr bool Q = true; };
template <> struct UnitTraits<GGML_TYPE_Q8_0> { static constexpr int BYTES = 144, BPS = 4; static constexpr bool Q = true; };

template <int TYPE>
__device__ __forceinline__ float unit_dot(const uint8_t *w, const XBlk (&x)[4], float acc)
{
    if constexpr (TYPE == GGML_TYPE_Q4_0) return q4_0_unit(w, x, acc);
    else if constexpr (TYPE == GGML_TYPE_Q4_1) return q4_1_unit(w, x, acc);
    else if constexpr (UnitTraits<TYPE>::Q) return sib_unit<TYPE>(w, x, acc);
    else return acc;
}

// partial dot of `nunits` staged units starting at unit u0 of the row
template <int TYPE, int NC, bool XREG>
__device__ __forceinline__ void dot_units(const uint8_t *w, int u0, int nunits, int S, int kb, const uint8_t *xs, int xcol_bytes,
                                          const XBlk (&xr)[4], int lane, float (&acc)[NC])
{
    constexpr int UB = UnitTraits<TYPE>::BYTES, BPS = UnitTraits<TYPE>::BPS;
    if (UnitTraits<TYPE>::Q) {
        if (XREG) {                                               // NC == 1, S <= 32: this lane owns unit `lane`
            if (lane < nunits) acc[0] = unit_dot<TYPE>(w + lane * UB, xr, acc[0]);
        } else {
            for (int u = lane; u < nunits; u += 32) {
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    XBlk x[4];
#pragma unroll
                    for (int j = 0; j < BPS; j++) x[j] = load_xblk(xs + c * xcol_bytes, kb, j * S + u0 + u);
                    acc[c] = unit_dot<TYPE>(w + u * UB, x, acc[c]);
                }
            }
        }
    } else if (TYPE == GGML_TYPE_F16) {
        dot_f16<NC>(w, u0 * 16, nunits * 16, xs, xcol_bytes, lane, acc);
    } else {
        dot_f32<NC>(w, u0 * 16, nunits * 16, xs, xcol_bytes, lane, acc);
    }
}

// CTA = NWF consumer warps + 1 producer warp.  A stage holds one "tile": TR = NWF*rpw consecutive rows
// (or, for rows longer than one copy, one K-chunk of NWF rows); the producer warp's lanes issue one
// bulk copy per row onto the stage's full-barrier and run `depth` stages ahead; consumer warp w owns
// rows w*rpw .. w*rpw+rpw-1 of every tile and releases the stage through its empty-barrier.
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int TYPE, int NC, bool XREG, int CAP>
__global__ void __launch_bounds__((NWF + 1) * 32, 1) k_gemv_fast(const __grid_constant__ GemvBatchT<CAP> b)
{
    constexpr int UB = UnitTraits<TYPE>::BYTES, BPS = UnitTraits<TYPE>::BPS;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int depth = b.depth, stage_bytes = b.stage_bytes, rpw = b.rs, nchunk = b.nchunk, row_bytes = b.row_bytes, chunk_bytes = b.chunk_bytes;
    const int TR = NWF * rpw;
    const int xbytes = (NC * b.xcol_bytes + 127) & ~127;
    uint8_t *xs = smem;
    uint8_t *stages = smem + xbytes;
    const uint32_t stage0 = smem_u32(stages);
    const uint32_t full0 = smem_u32(smem + xbytes + (size_t)depth * stage_bytes), empty0 = full0 + 8 * MAX_DEPTH;

    const int t_begin = (int)((long long)b.total_groups * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)b.total_groups * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NWF); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NWF) {
        // ===== producer warp: weights normally do not depend on the preceding kernels, so no PDL wait here.  When they do (src0 is an
        // earlier node's result: the writer may still be running, because the kernel between it and this one released its dependents
        // before it waited itself), the copies are held back like the activation reads =====
        if (b.wait_w) asm volatile("griddepcontrol.wait;" ::: "memory");
        const long long nb01 = b.nb01;
        int n = 0, st = 0; uint32_t ph = 0;
        for (int t = t_begin; t < t_end; t++) {
            while (t >= b.node[n].g0 + b.node[n].ngroups) n++;
            const int row0 = (t - b.node[n].g0) * TR;
            const int rows = min(TR, b.node[n].M - row0);
            const uint8_t *src = b.node[n].W + (long long)row0 * nb01;
            for (int c = 0; c < nchunk; c++) {
                const int cb = nchunk == 1 ? row_bytes : min(chunk_bytes, row_bytes - c * chunk_bytes);
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                if (lane == 0) mbar_expect_tx(full0 + 8 * st, (uint32_t)(rows * cb));
                __syncwarp();
                const uint32_t dst = stage0 + st * stage_bytes;
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s(dst + r * cb, src + (long long)r * nb01 + (long long)c * chunk_bytes, (uint32_t)cb, full0 + 8 * st);
                if (++st == depth) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ===== consumer warps =====
    asm volatile("griddepcontrol.wait;" ::: "memory");           // activations come from k_act_batch
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int xcol_bytes = b.xcol_bytes, kb = b.K / GGB_QK;
    const int S = row_bytes / UB, chunk_units = chunk_bytes / UB;
    int n = -1, st = 0; uint32_t ph = 0;
    int nt_end = 0, nM = 0, nldy = 0, nt0 = 0;
    float *ny = nullptr;
    XBlk xr[4];
    for (int t = t_begin; t < t_end; t++) {
        if (t >= nt_end) {
            // next node: (re)load this node's activations
            do { n++; } while (t >= b.node[n].g0 + b.node[n].ngroups);
            nt0 = b.node[n].g0; nt_end = nt0 + b.node[n].ngroups; nM = b.node[n].M; nldy = b.node[n].ldy; ny = b.node[n].y;
            if (XREG) {
                const uint8_t *xq = b.node[n].xq;
                if (b.fuse_x) {
                    // xq is the F32 row: the consumer warps quantize it into shared memory exactly as k_act_batch would have staged
                    // it (one definition, ggb_act_q8.cuh) -- 16 KB from L2 per CTA instead of a launch in front of every small level
                    asm volatile("bar.sync 1, %0;" ::"n"(NWF * 32) : "memory");      // every lane has taken its registers from the previous row
                    for (int base = 0; base < kb; base += NWF * 4) {                 // 8 lanes per block; uniform trip count (shuffles)
                        const int col = base + (int)(threadIdx.x >> 3), sub = lane & 7;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col < kb) v = __ldcg(reinterpret_cast<const float4 *>(xq + (long long)col * 128 + sub * 16));
                        uint32_t ev, od; float d; int sm;
                        q8_block_sub8(v, sub, TYPE == GGML_TYPE_Q4_2, ev, od, d, sm);
                        if (col < kb) q8_block_store(xs, kb, BPS, col, sub, ev, od, d, sm);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(NWF * 32) : "memory");
                    xq = xs;
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j < BPS) {
                        xr[j] = lane < S ? load_xblk(xq, kb, j * S + lane) : XBlk{make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), make_int2(0, 0)};
                    }
            } else {
                asm volatile("bar.sync 1, %0;" ::"n"(NWF * 32) : "memory");      // all consumers are done with the previous x
                if (UnitTraits<TYPE>::Q && NC == 1 && b.fuse_x) {
                    // long rows (more than 32 units, K-chunked stages): the same self-staging as above, the blocks stay in shared memory
                    const uint8_t *xf = b.node[n].xq;
                    for (int base = 0; base < kb; base += NWF * 4) {
                        const int col = base + (int)(threadIdx.x >> 3), sub = lane & 7;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col < kb) v = __ldcg(reinterpret_cast<const float4 *>(xf + (long long)col * 128 + sub * 16));
                        uint32_t ev, od; float d; int sm;
                        q8_block_sub8(v, sub, TYPE == GGML_TYPE_Q4_2, ev, od, d, sm);
                        if (col < kb) q8_block_store(xs, kb, BPS, col, sub, ev, od, d, sm);
                    }
                } else {
                    const uint4 *src = reinterpret_cast<const uint4 *>(b.node[n].xq);
                    uint4 *dst = reinterpret_cast<uint4 *>(xs);
                    for (int i = threadIdx.x; i < (NC * xcol_bytes) >> 4; i += NWF * 32) dst[i] = src[i];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NWF * 32) : "memory");
            }
        }
        const int row_base = (t - nt0) * TR + warp * rpw;
        const int nrows = min(rpw, nM - row_base);               // <= 0: this warp has no row in a ragged last tile
        float acc[NC];
        for (int c = 0; c < nchunk; c++) {
            mbar_wait(full0 + 8 * st, ph);
            const uint8_t *stage = stages + (size_t)st * stage_bytes;
            if (nchunk == 1) {
                for (int r = 0; r < nrows; r++) {
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
                    dot_units<TYPE, NC, XREG>(stage + (warp * rpw + r) * row_bytes, 0, S, S, kb, xs, xcol_bytes, xr, lane, acc);
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) {
                        const float v = warp_sum(acc[cc]);
                        if (lane == 0) {
                            float *yp = ny + (long long)cc * nldy + row_base + r;
                            *yp = v;
                            for (int p = 0; p < b.n_peers; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + b.peer_delta[p]) = v;
                        }
                    }
                }
            } else if (nrows > 0) {
                if (c == 0) {
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
                }
                const int cbytes = min(chunk_bytes, row_bytes - c * chunk_bytes);
                dot_units<TYPE, NC, false>(stage + warp * cbytes, c * chunk_units, cbytes / UB, S, kb, xs, xcol_bytes, xr, lane, acc);
                if (c == nchunk - 1) {
#pragma unroll
                    for (int cc = 0; cc < NC; cc++) {
                        const float v = warp_sum(acc[cc]);
                        if (lane == 0) {
                            float *yp = ny + (long long)cc * nldy + row_base;
                            *yp = v;
                            for (int p = 0; p < b.n_peers; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + b.peer_delta[p]) = v;
                        }
                    }
                }
            }
            // Release the stage only after its bytes have been CONSUMED: the accumulators are named as inputs of an (empty) asm
            // so the compiler cannot sink the dot-product math below the arrive; in-order issue then guarantees every LDS of
            // this stage has returned before the barrier is signalled (the same WAR hazard bit the GEMM, profiles/README.md).
#pragma unroll
            for (int cc = 0; cc < NC; cc++) asm volatile("" ::"f"(acc[cc]) : "memory");
            __syncwarp();                                        // every lane has finished reading this stage
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
            if (++st == depth) { st = 0; ph ^= 1; }
        }
    }
}

template <int TYPE, int CAP>
int launch_fast_typed(const GemvBatchT<CAP> &b, size_t smem, int grid, cudaStream_t s, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((NWF + 1) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    constexpr bool Q = UnitTraits<TYPE>::Q;
    const bool xreg = Q && b.ncols == 1 && b.nchunk == 1 && b.row_bytes / UnitTraits<TYPE>::BYTES <= 32;
#define GGB_FAST_CASE(NCV, XR) { \
        static PerDeviceOnce attr_once; \
        if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(k_gemv_fast<TYPE, NCV, XR, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); } \
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_gemv_fast<TYPE, NCV, XR, CAP>, b)); }
    if (xreg) { if constexpr (Q) GGB_FAST_CASE(1, true) }
    else switch (b.ncols) {
        case 1: GGB_FAST_CASE(1, false) break;
        case 2: GGB_FAST_CASE(2, false) break;
        case 4: GGB_FAST_CASE(4, false) break;
        case 8: GGB_FAST_CASE(8, false) break;
        default: return set_error(GGB_E_INVALID, "gemv: ncols=%d", b.ncols);
    }
#undef GGB_FAST_CASE
    count_launch();
    return GGB_OK;
}

template <int TYPE, int CAP>
int launch_typed(const GemvBatchT<CAP> &b, size_t smem, int grid, cudaStream_t s, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
#define GGB_GEMV_CASE(NCV) case NCV: { \
        static PerDeviceOnce attr_once; \
        if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(k_gemv<TYPE, NCV, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); } \
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_gemv<TYPE, NCV, CAP>, b)); } break;
    switch (b.ncols) {
        GGB_GEMV_CASE(1) GGB_GEMV_CASE(2) GGB_GEMV_CASE(4) GGB_GEMV_CASE(8)
    default: return set_error(GGB_E_INVALID, "gemv: ncols=%d", b.ncols);
    }
#undef GGB_GEMV_CASE
    count_launch();
    return GGB_OK;
}

constexpr int X_BUDGET = 64 * 1024;        // activation columns resident in shared memory
static const int W_BUDGET = [] { const char *e = getenv("GGB200_GEMV_WBUDGET"); return e ? atoi(e) : 200 * 1024; }();       // weight stages of all warps (sweep: benchmarks/gemv_stage_sweep.sh; 160 -> 200 KB: +5 % F32, +10 % F16 K=11008)
static const int STAGE_MAX = [] { const char *e = getenv("GGB200_GEMV_STAGE_MAX"); return e ? atoi(e) : 4096; }();           // bytes of one bulk copy (one row, several short rows, or a K-chunk of a long row)

} // namespace

// This file is compiled TWICE (csrc/Makefile): GGB_GEMV_PART 0 instantiates the kernels of the four headline weight types plus
// the planning / dispatch code, GGB_GEMV_PART 1 those of the sibling formats -- 144 kernel instantiations in one translation unit
// made it the long pole of the build.
template <int CAP>
int launch_gemv_typed_sib(const GemvBatchT<CAP> &b, size_t smem, int grid, cudaStream_t s, bool pdl);

#if GGB_GEMV_PART == 1
template <int CAP>
int launch_gemv_typed_sib(const GemvBatchT<CAP> &b, size_t smem, int grid, cudaStream_t s, bool pdl)
{
    if (b.async) switch (b.type) {
    case GGML_TYPE_Q4_2: return launch_fast_typed<GGML_TYPE_Q4_2, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q5_0: return launch_fast_typed<GGML_TYPE_Q5_0, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q5_1: return launch_fast_typed<GGML_TYPE_Q5_1, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q8_0: return launch_fast_typed<GGML_TYPE_Q8_0, CAP>(b, smem, grid, s, pdl);
    default: return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d is not on this path (Ggml.cs:6719-6742)", b.type);
    }
    switch (b.type) {
    case GGML_TYPE_Q4_2: return launch_typed<GGML_TYPE_Q4_2, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q5_0: return launch_typed<GGML_TYPE_Q5_0, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q5_1: return launch_typed<GGML_TYPE_Q5_1, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q8_0: return launch_typed<GGML_TYPE_Q8_0, CAP>(b, smem, grid, s, pdl);
    default: return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d is not on this path (Ggml.cs:6719-6742)", b.type);
    }
}
template int launch_gemv_typed_sib<GGB_SMALL_BATCH_NODES>(const GemvBatchT<GGB_SMALL_BATCH_NODES> &, size_t, int, cudaStream_t, bool);
template int launch_gemv_typed_sib<GGB_MAX_BATCH_NODES>(const GemvBatchT<GGB_MAX_BATCH_NODES> &, size_t, int, cudaStream_t, bool);
#else
extern template int launch_gemv_typed_sib<GGB_SMALL_BATCH_NODES>(const GemvBatchT<GGB_SMALL_BATCH_NODES> &, size_t, int, cudaStream_t, bool);
extern template int launch_gemv_typed_sib<GGB_MAX_BATCH_NODES>(const GemvBatchT<GGB_MAX_BATCH_NODES> &, size_t, int, cudaStream_t, bool);

// One persistent CTA per SM -- minus GGB200_GEMV_SPARE_SMS (default 0).  Measured (benchmarks/coresidency_probe.py): a small kernel on
// another stream is NOT placed beside a resident GEMV CTA even when registers, threads and shared memory would allow it; it gets an SM
// only when a GEMV CTA exits.  Leaving a few SMs to the side kernels (staging, result copy, peer push) is what lets them overlap.
int gemv_num_ctas()
{
    static const int spare = [] { const char *e = getenv("GGB200_GEMV_SPARE_SMS"); return e ? atoi(e) : 0; }();
    return std::max(1, device_sm_count() - std::max(0, spare));
}
int64_t gemv_x_budget() { return X_BUDGET + 32 * 1024; }
int gemv_group_rows(const GemvHdr &b) { return b.async ? NWF * b.rs : b.rs; }
static int unit_bytes_async(int type)
{
    switch (type) { case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_2: return 80; case GGML_TYPE_Q4_1: case GGML_TYPE_Q5_1: return 48;
                    case GGML_TYPE_Q5_0: return 88; case GGML_TYPE_Q8_0: return 144; default: return 16; }
}
int gemv_act_bps(const GemvHdr &b)
{
    if (!b.async) return 1;
    switch (b.type) { case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_2: case GGML_TYPE_Q5_0: case GGML_TYPE_Q8_0: return 4;
                      case GGML_TYPE_Q4_1: case GGML_TYPE_Q5_1: return 2; default: return 1; }
}
// the fast kernel, a quantized type, one activation column: its consumer warps can build the Q8 blocks of the row themselves
bool gemv_can_fuse_x(const GemvHdr &b)
{
    if (!b.async || !is_q_weight(b.type) || b.ncols != 1) return false;
    const int ub = unit_bytes_async(b.type);
    return ub > 0 && b.row_bytes % ub == 0;
}

// Fills the shape-dependent fields of b (everything but the node list).  ncols in {1,2,4,8}.
int gemv_plan(GemvHdr &b, int type, int64_t K, int64_t nb01, int ncols, const void *Wbase)
{
    const int blck = blck_size(type);
    if (K <= 0 || K % blck) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld is not a multiple of the block size %d (Ggml.cs:6694)", (long long)K, blck);
    const long long row_bytes = (long long)(K / blck) * (long long)type_size(type);
    const long long xcol = (long long)act_row_bytes(type, K);
    if (row_bytes > (1ll << 30) || xcol * ncols > gemv_x_budget())
        return set_error(GGB_E_UNSUPPORTED, "mul_mat: K=%lld too large for the shared-memory resident activation vector", (long long)K);
    b.type = type; b.ncols = ncols; b.K = (int)K; b.row_bytes = (int)row_bytes; b.nb01 = nb01; b.xcol_bytes = (int)xcol;
    const int unit_async = unit_bytes_async(type);
    const int unit_sync = is_q_weight(type) ? (int)q32_bytes(type) : 16;
    // 16-byte aligned rows made of whole units -> TMA bulk staging + the fast kernel; anything else -> plain-load staging
    b.async = ((reinterpret_cast<uintptr_t>(Wbase) & 15) == 0 && (nb01 & 15) == 0 && (row_bytes & 15) == 0 && row_bytes % unit_async == 0) ? 1 : 0;
    // K-chunks of a long row are bulk copies too, so a chunk must also be a multiple of 16 bytes (Q5_0's 88-byte unit is not)
    const int unit = b.async ? (unit_async % 16 ? unit_async * 2 : unit_async) : unit_sync;
    if (b.async) {
        // fast kernel: stage = 16*rpw whole rows, or one K-chunk of 16 rows
        if (row_bytes <= STAGE_MAX) {
            b.nchunk = 1; b.chunk_bytes = (int)row_bytes;
            b.rs = (int)std::max<long long>(1, std::min<long long>(8, 2048 / row_bytes));     // rows per consumer warp per stage
            b.stage_bytes = (int)(NWF * b.rs * row_bytes);
        } else {
            const long long units = row_bytes / unit;
            long long nchunk = (row_bytes + STAGE_MAX - 1) / STAGE_MAX;
            const long long cu = (units + nchunk - 1) / nchunk;
            b.chunk_bytes = (int)(cu * unit);
            b.nchunk = (int)((row_bytes + b.chunk_bytes - 1) / b.chunk_bytes);
            b.rs = 1;
            b.stage_bytes = NWF * b.chunk_bytes;
        }
        const long long xb = (xcol * ncols + 127) & ~127ll;
        long long budget = 227 * 1024 - xb - 2 * MAX_DEPTH * 8 - 1024;
        if (budget > W_BUDGET) budget = W_BUDGET;
        int d = (int)(budget / b.stage_bytes);
        if (d > MAX_DEPTH) d = MAX_DEPTH;
        if (d < 2) return set_error(GGB_E_UNSUPPORTED, "mul_mat: shape needs more shared memory than one SM has");
        b.depth = d;
        return GGB_OK;
    }
    if (row_bytes <= STAGE_MAX) {
        b.nchunk = 1;
        b.chunk_bytes = (int)row_bytes;
        b.rs = 1;
        b.stage_bytes = (int)align_up((size_t)row_bytes, 16);
    } else {
        const long long units = (row_bytes + unit - 1) / unit;
        long long nchunk = (row_bytes + STAGE_MAX - 1) / STAGE_MAX;
        const long long cu = (units + nchunk - 1) / nchunk;
        b.chunk_bytes = (int)(cu * unit);
        b.nchunk = (int)((row_bytes + b.chunk_bytes - 1) / b.chunk_bytes);
        b.rs = 1;
        b.stage_bytes = (int)align_up((size_t)b.chunk_bytes, 16);
    }
    const long long xbytes = (xcol * ncols + 127) & ~127ll;
    long long wbudget = 227 * 1024 - xbytes - NWARPS * MAX_DEPTH * 8 - 1024;
    if (wbudget > W_BUDGET) wbudget = W_BUDGET;
    int depth = 1;
    if ((long long)NWARPS * b.stage_bytes > wbudget) return set_error(GGB_E_UNSUPPORTED, "mul_mat: shape needs more shared memory than one SM has");
    b.depth = depth;
    return GGB_OK;
}


// ================================================================ the decode program (ggb_internal.h: DpProgram)
namespace {

constexpr int DP_NT = NWF * 32;                                   // consumer threads

// consumers of all CTAs: everything written before is visible to everyone after (the cooperative-groups grid.sync pattern: CTA barrier,
// one thread counts + spins, CTA barrier).  The counter only grows: barrier k completes at k * gridDim.x arrivals.
__device__ __forceinline__ void dp_grid_barrier(unsigned *bar, unsigned target)
{
    asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");
    if (threadIdx.x == 0) {
        // release: this CTA's writes (ordered before by the CTA barrier) are visible before the count; acquire: what the other CTAs
        // released is visible to everyone behind the second CTA barrier.  No separate fences: each would cost another L2 round trip.
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");
}

__device__ __forceinline__ void dp_stamp(const DpProgram &P, int si, int k)
{
    if (P.trace && threadIdx.x == 0 && (blockIdx.x & 31) == 0 && (blockIdx.x >> 5) < 4 && si < 64) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.trace[(((blockIdx.x >> 5) * 64) + si) * 8 + k] = t;
    }
}

// a mul_mat result, and -- when the next node of the graph is the SILU of it -- that node's result as well: the one lookup per row
// disappears among the tiles, where the same SILU as a row op of the next step would have every CTA gather the whole row's entries
__device__ __forceinline__ void dp_store_y(const DpProgram &P, const DpNode &nd, int row, float v)
{
    if (nd.y2 != nd.y) nd.y[row] = v;
    if (nd.y2) nd.y2[row] = __half2float(__ushort_as_half(__ldg(P.silu_table + __half_as_ushort(__float2half_rn(v)))));
}

// one step's tiles of this CTA against the staged row in xs; the arithmetic of a row is k_gemv_fast's (dot_units, then warp_sum)
template <int TYPE>
__device__ __forceinline__ void dp_consume(const DpProgram &P, const DpStep &sp, const uint8_t *xs, const uint8_t *stages, uint32_t full0, uint32_t empty0,
                                           int warp, int lane, int &st, uint32_t &ph)
{
    constexpr int UB = UnitTraits<TYPE>::BYTES;
    const int rpw = sp.rs, TR = NWF * rpw, nchunk = sp.nchunk, row_bytes = sp.row_bytes, chunk_bytes = sp.chunk_bytes;
    const int kb = sp.K / GGB_QK, S = row_bytes / UB, chunk_units = chunk_bytes / UB;
    const int t_begin = (int)((long long)sp.total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)sp.total_tiles * (blockIdx.x + 1) / gridDim.x);
    const XBlk xr[4] = {};
    int n = sp.node0;
    for (int t = t_begin; t < t_end; t++) {
        while (t >= P.node[n].tile0 + P.node[n].ntiles) n++;
        const DpNode &nd = P.node[n];
        const int row_base = (t - nd.tile0) * TR + warp * rpw;
        const int nrows = min(rpw, nd.M - row_base);
        float acc[1];
        for (int c = 0; c < nchunk; c++) {
            mbar_wait(full0 + 8 * st, ph);
            const uint8_t *stage = stages + (size_t)st * P.slot_bytes;
            if (nchunk == 1) {
                for (int r = 0; r < nrows; r++) {
                    acc[0] = 0.0f;
                    dot_units<TYPE, 1, false>(stage + (warp * rpw + r) * row_bytes, 0, S, S, kb, xs, 0, xr, lane, acc);
                    const float v = warp_sum(acc[0]);
                    if (lane == 0) dp_store_y(P, nd, row_base + r, v);
                }
            } else if (nrows > 0) {
                if (c == 0) acc[0] = 0.0f;
                const int cbytes = min(chunk_bytes, row_bytes - c * chunk_bytes);
                dot_units<TYPE, 1, false>(stage + warp * cbytes, c * chunk_units, cbytes / UB, S, kb, xs, 0, xr, lane, acc);
                if (c == nchunk - 1) {
                    const float v = warp_sum(acc[0]);
                    if (lane == 0) dp_store_y(P, nd, row_base, v);
                }
            }
            asm volatile("" ::"f"(acc[0]) : "memory");           // the stage is released only after its bytes have been consumed (see k_gemv_fast)
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
            if (++st == P.depth) { st = 0; ph ^= 1; }
        }
    }
}

__global__ void __launch_bounds__((NWF + 1) * 32, 1) k_decode_program(const __grid_constant__ DpProgram P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    float *v = reinterpret_cast<float *>(smem);                  // the running row
    uint8_t *xq = smem + P.v_bytes;                              // ... as the step's GEMVs read it: Q8 blocks / halves (F32 weights read v itself)
    uint8_t *stages = xq + P.xs_bytes;
    const uint32_t stage0 = smem_u32(stages);
    const uint32_t full0 = smem_u32(stages + (size_t)P.depth * P.slot_bytes), empty0 = full0 + 8 * MAX_DEPTH;
    __shared__ double part[8];

    if (tid == 0) {
        for (int s = 0; s < P.depth; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NWF); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NWF) {
        // ===== producer warp: the weight tiles of every step, in step order, as far ahead as the ring allows -- across the grid barriers,
        // the consumers' prologues and their waits for other CTAs.  (The builder takes only weights nothing in this launch writes.) =====
        int st = 0; uint32_t ph = 0;
        for (int si = 0; si < P.n_steps; si++) {
            const DpStep &sp = P.step[si];
            if (!sp.nnodes) continue;
            const int TR = NWF * sp.rs, nchunk = sp.nchunk, row_bytes = sp.row_bytes, chunk_bytes = sp.chunk_bytes;
            const int t_begin = (int)((long long)sp.total_tiles * blockIdx.x / gridDim.x);
            const int t_end = (int)((long long)sp.total_tiles * (blockIdx.x + 1) / gridDim.x);
            int n = sp.node0;
            for (int t = t_begin; t < t_end; t++) {
                while (t >= P.node[n].tile0 + P.node[n].ntiles) n++;
                const DpNode &nd = P.node[n];
                const int row0 = (t - nd.tile0) * TR;
                const int rows = min(TR, nd.M - row0);
                const uint8_t *src = nd.W + (long long)row0 * nd.nb01;
                for (int c = 0; c < nchunk; c++) {
                    const int cb = nchunk == 1 ? row_bytes : min(chunk_bytes, row_bytes - c * chunk_bytes);
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    if (lane == 0) mbar_expect_tx(full0 + 8 * st, (uint32_t)(rows * cb));
                    __syncwarp();
                    const uint32_t dst = stage0 + st * P.slot_bytes;
                    for (int r = lane; r < rows; r += 32)
                        bulk_g2s(dst + r * cb, src + (long long)r * nd.nb01 + (long long)c * chunk_bytes, (uint32_t)cb, full0 + 8 * st);
                    if (++st == P.depth) { st = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    int st = 0; uint32_t ph = 0; unsigned nbar = 0;
    for (int si = 0; si < P.n_steps; si++) {
        const DpStep &sp = P.step[si];
        dp_stamp(P, si, 0);
        // ---- results of earlier steps go to the host arena: one warp per tensor, this CTA's slice of it (coalesced 16-byte stores that
        //      nobody waits for) ----
        for (int ci = sp.copy0 + warp; ci < sp.copy0 + sp.ncopies; ci += NWF) {
            const DpCopy &cp = P.copy[ci];
            const int per = (cp.n4 + (int)gridDim.x - 1) / (int)gridDim.x;
            const int lo = (int)blockIdx.x * per, hi = min(cp.n4, lo + per);
            for (int i = lo + lane; i < hi; i += 32) cp.dst[i] = __ldcg(cp.src + i);
        }
        // ---- the step's row ops.  float4 i of the running row belongs to thread i mod DP_NT in every op, so element-wise ops need
        //      no barrier between them; operands other CTAs wrote (before the last grid barrier) are read past L1 (__ldcg) ----
        for (int oi = sp.op0; oi < sp.op0 + sp.nops; oi++) {
            const DpOp &op = P.op[oi];
            const int n = op.n, n4 = n >> 2;                                              // (the builder takes rows of whole, aligned float4s only)
            const int per = (n4 + (int)gridDim.x - 1) / (int)gridDim.x;
            const int lo = (int)blockIdx.x * per, hi = min(n4, lo + per);                 // this CTA's slice of the result tensor, in float4s
            float4 *v4 = reinterpret_cast<float4 *>(v);
            const float4 *a4 = reinterpret_cast<const float4 *>(op.a), *b4 = reinterpret_cast<const float4 *>(op.b);
            float4 *d4 = reinterpret_cast<float4 *>(op.dst);
            constexpr int UNR = 4;                                                        // loads in flight per thread and operand: a latency, not a loop of them
            float scale = op.scalar;
            if (op.op == DP_RMS_NORM) {
                if (a4) {
                    for (int base = tid; base < n4; base += DP_NT * UNR) {
                        float4 x[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; u++) if (base + u * DP_NT < n4) x[u] = __ldcg(a4 + base + u * DP_NT);
#pragma unroll
                        for (int u = 0; u < UNR; u++) if (base + u * DP_NT < n4) v4[base + u * DP_NT] = x[u];
                    }
                }
                // k_rms_norm_f32's reduction, order included (256 threads stride the row, float products summed in double, shuffle tree,
                // eight partials added in order), so that the two routes agree to the bit (Ggml.cs:5858-5921)
                asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");               // the running row is complete
                if (tid < 256) {
                    double sum = 0.0;
                    for (int i = tid; i < n; i += 256) { const float x = v[i]; sum += (double)__fmul_rn(x, x); }
#pragma unroll
                    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                    if (lane == 0) part[warp] = sum;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");
                double t = 0.0;
                for (int w = 0; w < 8; w++) t += part[w];
                const float mean = (float)(t / (double)n);
                scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, 1e-6f)));
                a4 = nullptr;                                                             // the row is in v now
            }
            for (int base = tid; base < n4; base += DP_NT * UNR) {
                float4 x[UNR], y[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int i = base + u * DP_NT;
                    if (i < n4) {
                        x[u] = a4 ? __ldcg(a4 + i) : v4[i];
                        if (op.op == DP_ADD || op.op == DP_MUL) y[u] = __ldcg(b4 + i);
                    }
                }
                if (op.op == DP_SILU) {
                    // (float)table_silu_f16[(Half)x] (Ggml.cs:2736-2746): all lookups of the batch in flight together
                    unsigned short h[UNR][4];
#pragma unroll
                    for (int u = 0; u < UNR; u++) {
                        if (base + u * DP_NT < n4) {
                            h[u][0] = __ldg(P.silu_table + __half_as_ushort(__float2half_rn(x[u].x)));
                            h[u][1] = __ldg(P.silu_table + __half_as_ushort(__float2half_rn(x[u].y)));
                            h[u][2] = __ldg(P.silu_table + __half_as_ushort(__float2half_rn(x[u].z)));
                            h[u][3] = __ldg(P.silu_table + __half_as_ushort(__float2half_rn(x[u].w)));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; u++)
                        if (base + u * DP_NT < n4)
                            x[u] = make_float4(__half2float(__ushort_as_half(h[u][0])), __half2float(__ushort_as_half(h[u][1])),
                                               __half2float(__ushort_as_half(h[u][2])), __half2float(__ushort_as_half(h[u][3])));
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int i = base + u * DP_NT;
                    if (i < n4) {
                        float4 r = x[u];
                        if (op.op == DP_ADD) { r.x = __fadd_rn(r.x, y[u].x); r.y = __fadd_rn(r.y, y[u].y); r.z = __fadd_rn(r.z, y[u].z); r.w = __fadd_rn(r.w, y[u].w); }
                        else if (op.op == DP_MUL) { r.x = __fmul_rn(r.x, y[u].x); r.y = __fmul_rn(r.y, y[u].y); r.z = __fmul_rn(r.z, y[u].z); r.w = __fmul_rn(r.w, y[u].w); }
                        else if (op.op == DP_SCALE || op.op == DP_RMS_NORM) { r.x = __fmul_rn(r.x, scale); r.y = __fmul_rn(r.y, scale); r.z = __fmul_rn(r.z, scale); r.w = __fmul_rn(r.w, scale); }
                        v4[i] = r;
                        if (d4 && i >= lo && i < hi) d4[i] = r;
                    }
                }
            }
        }
        dp_stamp(P, si, 1);
        if (sp.nnodes) {
            // ---- the row as this step's weight type multiplies it (k_act_batch's conversions; Q8 blocks through ggb_act_q8.cuh) ----
            asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");
            const int kb = sp.K / GGB_QK;
            const uint8_t *xs = xq;
            if (sp.type == GGML_TYPE_F32) xs = reinterpret_cast<const uint8_t *>(v);
            else if (sp.type == GGML_TYPE_F16) {
                for (int i = tid; i < sp.K; i += DP_NT) reinterpret_cast<__half *>(xq)[i] = __float2half_rn(v[i]);
            } else {
                const int bps = (sp.type == GGML_TYPE_Q4_1 || sp.type == GGML_TYPE_Q5_1) ? 2 : 4;
                const bool q4_2 = sp.type == GGML_TYPE_Q4_2;                          // (its two 16-element blocks need their own half sums)
                for (int base = 0; base < kb; base += DP_NT / 2) {                        // two lanes per block; uniform trip count (shuffles)
                    const int col = base + (tid >> 1), h = tid & 1;
                    float e[16];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float4 x4 = col < kb ? *reinterpret_cast<const float4 *>(v + col * GGB_QK + h * 16 + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        e[4 * i] = x4.x; e[4 * i + 1] = x4.y; e[4 * i + 2] = x4.z; e[4 * i + 3] = x4.w;
                    }
                    uint32_t ev[2], od[2]; float d; int sm;
                    q8_block_half16(e, q4_2, ev, od, d, sm);
                    if (col < kb) q8_block_half16_store(xq, kb, bps, col, h, ev, od, d, sm);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(DP_NT) : "memory");
            dp_stamp(P, si, 2);
            switch (sp.type) {
            case GGML_TYPE_Q4_0: dp_consume<GGML_TYPE_Q4_0>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_Q4_1: dp_consume<GGML_TYPE_Q4_1>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_Q4_2: dp_consume<GGML_TYPE_Q4_2>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_Q5_0: dp_consume<GGML_TYPE_Q5_0>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_Q5_1: dp_consume<GGML_TYPE_Q5_1>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_Q8_0: dp_consume<GGML_TYPE_Q8_0>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            case GGML_TYPE_F16: dp_consume<GGML_TYPE_F16>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            default: dp_consume<GGML_TYPE_F32>(P, sp, xs, stages, full0, empty0, warp, lane, st, ph); break;
            }
        }
        dp_stamp(P, si, 3);
        if (si + 1 < P.n_steps) dp_grid_barrier(P.bar, ++nbar * gridDim.x);
        dp_stamp(P, si, 4);
    }
}

} // namespace

int64_t decode_program_row_max() { return 12288; }
int decode_program_tile_rows(const DpStep &st) { return NWF * st.rs; }

bool decode_program_plan_step(DpStep &st, int type, int64_t K, int64_t nb01, const void *W)
{
    if (!is_q_weight(type) && type != GGML_TYPE_F16 && type != GGML_TYPE_F32) return false;
    if (K <= 0 || K > decode_program_row_max() || K % 4) return false;
    GemvHdr h = {};
    if (gemv_plan(h, type, K, nb01, 1, W) || !h.async) return false;          // (a refusal leaves its text in the error buffer; the caller falls back and nobody reads it)
    st.type = type; st.K = (int)K; st.row_bytes = h.row_bytes; st.rs = h.rs; st.nchunk = h.nchunk; st.chunk_bytes = h.chunk_bytes; st.stage_bytes = h.stage_bytes;
    return true;
}

bool decode_program_finish(DpProgram &p)
{
    int64_t vmax = 0, xmax = 0, slot = 0;
    for (int i = 0; i < p.n_steps; i++) {
        const DpStep &st = p.step[i];
        for (int o = st.op0; o < st.op0 + st.nops; o++) vmax = std::max<int64_t>(vmax, p.op[o].n);
        if (!st.nnodes) continue;
        vmax = std::max<int64_t>(vmax, st.K);
        if (st.type != GGML_TYPE_F32) xmax = std::max<int64_t>(xmax, (int64_t)act_row_bytes(st.type, st.K));
        slot = std::max<int64_t>(slot, st.stage_bytes);
    }
    if (vmax > decode_program_row_max()) return false;
    p.v_bytes = (int)align_up((size_t)vmax * 4, 128);
    p.xs_bytes = (int)align_up((size_t)xmax, 128);
    p.slot_bytes = (int)align_up((size_t)std::max<int64_t>(slot, 128), 128);
    const int64_t room = 227 * 1024 - 1024 - p.v_bytes - p.xs_bytes - 2 * MAX_DEPTH * 8;
    p.depth = (int)std::min<int64_t>(MAX_DEPTH, room / p.slot_bytes);
    return p.depth >= 2;
}

// the grid barrier needs every CTA resident at once: a device (or an MPS partition) that cannot launch one 512-thread CTA with the
// full shared memory per SM cooperatively does not get the program
bool decode_program_available()
{
    static int state[16] = {};                                    // per device: 0 unknown, 1 yes, -1 no
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return false; }
    if (state[dev] == 0) {
        int coop = 0, per_sm = 0;
        bool ok = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop != 0;
        ok = ok && cudaFuncSetAttribute(k_decode_program, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_program, (NWF + 1) * 32, 227 * 1024 - 512) == cudaSuccess && per_sm >= 1;
        cudaGetLastError();
        state[dev] = ok ? 1 : -1;
    }
    return state[dev] > 0;
}

int launch_decode_program(const DpProgram &p, cudaStream_t s)
{
    static PerDeviceOnce once;
    if (once.need()) GGB_CUDA(cudaFuncSetAttribute(k_decode_program, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    const size_t smem = (size_t)p.v_bytes + p.xs_bytes + (size_t)p.depth * p.slot_bytes + 2 * MAX_DEPTH * 8;
    GGB_CUDA(cudaMemsetAsync(p.bar, 0, sizeof(unsigned), s));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)device_sm_count());             // one CTA per SM, all resident at once: the grid barrier needs it
    cfg.blockDim = dim3((NWF + 1) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_decode_program, p));
    count_launch();
    return GGB_OK;
}

template <int CAP>
static int launch_gemv_cap(const GemvBatchT<CAP> &b, cudaStream_t s, bool pdl)
{
    const size_t xbytes = ((size_t)b.ncols * b.xcol_bytes + 127) & ~(size_t)127;
    int grid = gemv_num_ctas();
    // async: groups are tiles and every CTA takes whole tiles; sync: groups are rows, one per warp
    const size_t smem = b.async ? xbytes + (size_t)b.depth * b.stage_bytes + 2 * MAX_DEPTH * 8
                                : xbytes + (size_t)NWARPS * b.stage_bytes;
    const int want = b.async ? b.total_groups : (b.total_groups + NWARPS - 1) / NWARPS;
    if (grid > want) grid = want;
    if (is_sibling_q(b.type)) return launch_gemv_typed_sib<CAP>(b, smem, grid, s, pdl);
    if (b.async) switch (b.type) {
    case GGML_TYPE_Q4_0: return launch_fast_typed<GGML_TYPE_Q4_0, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q4_1: return launch_fast_typed<GGML_TYPE_Q4_1, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_F16: return launch_fast_typed<GGML_TYPE_F16, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_F32: return launch_fast_typed<GGML_TYPE_F32, CAP>(b, smem, grid, s, pdl);
    default: return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d is not on this path (Ggml.cs:6719-6742)", b.type);
    }
    switch (b.type) {
    case GGML_TYPE_Q4_0: return launch_typed<GGML_TYPE_Q4_0, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q4_1: return launch_typed<GGML_TYPE_Q4_1, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_F16: return launch_typed<GGML_TYPE_F16, CAP>(b, smem, grid, s, pdl);
    case GGML_TYPE_F32: return launch_typed<GGML_TYPE_F32, CAP>(b, smem, grid, s, pdl);
    default: return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d is not on this path (Ggml.cs:6719-6742)", b.type);
    }
}

int launch_gemv_batch(const GemvBatch &b, cudaStream_t s, bool pdl)
{
    if (b.n_nodes <= 0 || b.total_groups <= 0) return GGB_OK;
    if (b.n_nodes <= GGB_SMALL_BATCH_NODES) {
        // few nodes: a 1.5 KB parameter block instead of 10 KB (large kernel parameters add microseconds to every launch)
        static thread_local GemvBatchT<GGB_SMALL_BATCH_NODES> sb;
        static_cast<GemvHdr &>(sb) = static_cast<const GemvHdr &>(b);
        for (int i = 0; i < b.n_nodes; i++) sb.node[i] = b.node[i];
        return launch_gemv_cap(sb, s, pdl);
    }
    return launch_gemv_cap(b, s, pdl);
}
#endif

} // namespace ggb
