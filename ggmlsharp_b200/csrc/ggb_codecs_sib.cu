// Row codecs of the sibling weight formats: Q4_2, Q5_0, Q5_1 and Q8_0 as a weight type.
//
//   quantize_row_q4_2_reference_impl  Ggml.cs:547-590     dequantize_row_q4_2  Ggml.cs:992-1022
//   quantize_row_q5_0_reference_impl  Ggml.cs:609-653     dequantize_row_q5_0  Ggml.cs:1025-1061
//   quantize_row_q5_1_reference_impl  Ggml.cs:672-714     dequantize_row_q5_1  Ggml.cs:1064-1101
//   quantize_row_q8_0_reference_impl  Ggml.cs:733-762     dequantize_row_q8_0  Ggml.cs:1104-1122
//
// Block layouts (TypeDefinitions.cs:249-282): q4_2 = {fp16 d; 8 nibble bytes} per 16 elements, q5_0 = {fp16 d; u32 qh; 16 nibble
// bytes}, q5_1 = {fp16 d; fp16 m; u32 qh; 16 nibble bytes}, q8_0 = {f32 d; 32 int8}.  fp16 scales cross the boundary as IEEE
// binary16 bit patterns (the reference's `(ushort)(Half)d` numeric cast is defect D9, see oracle/ggb_oracle.c) and Q8 quants are
// int8 (defect D4).  Same bit-exactness rules as ggb_codecs.cu: explicit round-to-nearest intrinsics, no FMA contraction.
//
// Everything here works on a "group" = the 32 consecutive elements that pair with one activation block (two Q4_2 blocks, one
// block of the others; 20 / 22 / 24 / 36 bytes).  Groups are only 2-byte aligned in memory (10- and 22-byte blocks), so block
// bytes move as 16-bit words.  Roofline: HBM, 4 B + 0.625 / 0.6875 / 0.75 / 1.125 B per element.
#include "ggb_internal.h"
#include "ggb_sib_math.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace ggb {

namespace {

using namespace sib;      // quantize_group / dequantize_group / h_val: ggb_sib_math.cuh (also compiled for the host by tests/emul)
template <int TYPE> using Sib = Grp<TYPE>;

template <int TYPE>
__global__ void __launch_bounds__(128) k_quantize_sib(const float *__restrict__ x, long long ldx, uint8_t *__restrict__ y, long long ngrp, int kb)
{
    constexpr int G = Sib<TYPE>::G;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += (long long)gridDim.x * blockDim.x) {
        const long long row = g / kb;
        const float4 *p = reinterpret_cast<const float4 *>(x + row * ldx + (g - row * kb) * GGB_QK);
        float e[32];
#pragma unroll
        for (int i = 0; i < 8; i++) { const float4 v = __ldg(p + i); e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w; }
        uint32_t o[G / 2];
        quantize_group<TYPE>(e, o);
        unsigned short *dst = reinterpret_cast<unsigned short *>(y + g * G);
#pragma unroll
        for (int i = 0; i < G / 2; i++) dst[i] = (unsigned short)o[i];
    }
}

// Same data movement as k_quantize_q4_tiles (ggb_codecs.cu): every WARP runs its own cp.async pipeline over tiles of 32 groups
// = 4 KB of contiguous source (eight coalesced 16-byte-per-lane copies into shared rows padded to 144 B, three tiles in flight),
// lane l quantizes group l from eight conflict-free LDS.128, and the 32 x G output bytes leave through a staging row as one
// contiguous run of 4-byte lane-consecutive stores.  (The first version -- k_quantize_sib, one thread per group with direct
// 128-bit loads that touch 32 different lines per instruction and 2-byte scattered stores -- reached 34-54 % of the copy peak.)
constexpr int ST_WARPS = 4, ST_STAGES = 3, ST_ROW = 144;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int TYPE>
__global__ void __launch_bounds__(ST_WARPS * 32) k_quantize_sib_tiles(const float *__restrict__ x, long long ldx, uint8_t *__restrict__ y,
                                                                      long long ngrp, int kb)
{
    constexpr int G = Sib<TYPE>::G, OSTAGE = (32 * G + 15) / 16 * 16;
    extern __shared__ uint4 st_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *wbase = reinterpret_cast<uint8_t *>(st_smem) + (size_t)warp * (ST_STAGES * 32 * ST_ROW + OSTAGE);
    const uint32_t in0 = (uint32_t)__cvta_generic_to_shared(wbase);
    uint8_t *ostage = wbase + ST_STAGES * 32 * ST_ROW;
    const long long ntiles = (ngrp + 31) >> 5;
    const long long gw = (long long)blockIdx.x * ST_WARPS + warp, nw = (long long)gridDim.x * ST_WARPS;
    const bool dense = ldx == (long long)kb * GGB_QK;
    const int sub = lane & 7, q = lane >> 3;

    auto src_of = [&](long long g) -> const float * {
        if (dense) return x + g * GGB_QK;
        const long long row = g / kb;
        return x + row * ldx + (g - row * kb) * GGB_QK;
    };
    auto issue = [&](long long tile, int stage) {
        if (tile < ntiles) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const long long g = tile * 32 + i * 4 + q;
                if (g < ngrp) cp_async16(in0 + (uint32_t)(stage * 32 * ST_ROW + (i * 4 + q) * ST_ROW + sub * 16), src_of(g) + sub * 4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");     // always commit so the group count stays uniform
    };

    issue(gw, 0);
    issue(gw + nw, 1);
    int stage = 0;
    for (long long tile = gw; tile < ntiles; tile += nw) {
        issue(tile + 2 * nw, stage == 0 ? 2 : stage - 1);       // the stage consumed in the previous iteration
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        if (tile * 32 + lane < ngrp) {
            float e[32];
            const uint4 *row = reinterpret_cast<const uint4 *>(wbase + stage * 32 * ST_ROW + lane * ST_ROW);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 v = row[i];
                e[4 * i] = __uint_as_float(v.x); e[4 * i + 1] = __uint_as_float(v.y); e[4 * i + 2] = __uint_as_float(v.z); e[4 * i + 3] = __uint_as_float(v.w);
            }
            uint32_t o[G / 2];
            quantize_group<TYPE>(e, o);
            if (G % 4 == 0) {                                    // 20 / 24 / 36-byte groups are word aligned in the staging row
                uint32_t *d = reinterpret_cast<uint32_t *>(ostage + lane * G);
#pragma unroll
                for (int i = 0; i < G / 4; i++) d[i] = o[2 * i] | (o[2 * i + 1] << 16);
            } else {
                unsigned short *d = reinterpret_cast<unsigned short *>(ostage + lane * G);
#pragma unroll
                for (int i = 0; i < G / 2; i++) d[i] = (unsigned short)o[i];
            }
        }
        __syncwarp();
        // nvalid groups x G bytes, contiguous in y (the tile base is a multiple of 32*G, so 4-byte aligned whenever y is)
        const int nvalid = (int)(ngrp - tile * 32 < 32 ? ngrp - tile * 32 : 32), nbytes = nvalid * G;
        uint8_t *yo = y + tile * 32 * G;
#pragma unroll
        for (int j = 0; j < (8 * G + 31) / 32; j++) {
            const int wi = j * 32 + lane;
            if (wi * 4 + 4 <= nbytes) reinterpret_cast<uint32_t *>(yo)[wi] = reinterpret_cast<const uint32_t *>(ostage)[wi];
        }
        if ((nbytes & 2) && lane == 0) reinterpret_cast<unsigned short *>(yo)[nbytes / 2 - 1] = reinterpret_cast<const unsigned short *>(ostage)[nbytes / 2 - 1];
        __syncwarp();                                            // ostage and this input stage are free again
        stage = stage == ST_STAGES - 1 ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// One thread per 16 output bytes: lane sub = t & 7 of a group expands elements 4*sub .. 4*sub+3, so a warp store is 512 contiguous
// bytes; the small scale / qh / quant loads of four independent slices are all issued before any of them is used.
template <int TYPE>
__global__ void __launch_bounds__(256) k_dequantize_sib(const uint8_t *__restrict__ x, float *__restrict__ y, long long ngrp)
{
    constexpr int G = Sib<TYPE>::G, U = 4;
    const long long total = ngrp * 8;
    for (long long base = (long long)blockIdx.x * (256 * U) + threadIdx.x; base < total; base += (long long)gridDim.x * (256 * U)) {
        uint32_t dm[U], qh[U], q[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long t = base + u * 256;
            dm[u] = 0; qh[u] = 0; q[u] = 0;
            if (t >= total) continue;
            const int sub = (int)(t & 7);
            const unsigned short *g = reinterpret_cast<const unsigned short *>(x + (t >> 3) * G);
            if (TYPE == GGML_TYPE_Q4_2) {
                const unsigned short *b = g + 5 * (sub >> 2);
                dm[u] = __ldg(b); q[u] = __ldg(b + 1 + (sub & 3));
            } else if (TYPE == GGML_TYPE_Q5_0) {
                dm[u] = __ldg(g); qh[u] = (uint32_t)__ldg(g + 1 + (sub >> 2)); q[u] = __ldg(g + 3 + sub);       // the half of qh this slice needs
            } else if (TYPE == GGML_TYPE_Q5_1) {
                dm[u] = (uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16); qh[u] = (uint32_t)__ldg(g + 2 + (sub >> 2)); q[u] = __ldg(g + 4 + sub);
            } else {
                dm[u] = (uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16);
                q[u] = (uint32_t)__ldg(g + 2 + 2 * sub) | ((uint32_t)__ldg(g + 3 + 2 * sub) << 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long t = base + u * 256;
            if (t >= total) continue;
            const int sub = (int)(t & 7);
            float4 o;
            if (TYPE == GGML_TYPE_Q4_2) {
                const float d = h_val(dm[u]);
                o.x = __fmul_rn((float)((int)(q[u] & 15u) - 8), d); o.y = __fmul_rn((float)((int)((q[u] >> 4) & 15u) - 8), d);
                o.z = __fmul_rn((float)((int)((q[u] >> 8) & 15u) - 8), d); o.w = __fmul_rn((float)((int)(q[u] >> 12) - 8), d);
            } else if (TYPE == GGML_TYPE_Q5_0 || TYPE == GGML_TYPE_Q5_1) {
                const float d = h_val(dm[u] & 0xFFFFu);
                const uint32_t h = qh[u] >> (4 * (sub & 3));                    // bits of elements 4*sub .. 4*sub+3
                const int n0 = (int)((q[u] & 15u) | ((h & 1u) << 4)), n1 = (int)(((q[u] >> 4) & 15u) | (((h >> 1) & 1u) << 4));
                const int n2 = (int)(((q[u] >> 8) & 15u) | (((h >> 2) & 1u) << 4)), n3 = (int)((q[u] >> 12) | (((h >> 3) & 1u) << 4));
                if (TYPE == GGML_TYPE_Q5_0) {
                    o.x = __fmul_rn((float)(n0 - 16), d); o.y = __fmul_rn((float)(n1 - 16), d);
                    o.z = __fmul_rn((float)(n2 - 16), d); o.w = __fmul_rn((float)(n3 - 16), d);
                } else {                                         // product rounded, then sum rounded
                    const float m = h_val(dm[u] >> 16);
                    o.x = __fadd_rn(__fmul_rn((float)n0, d), m); o.y = __fadd_rn(__fmul_rn((float)n1, d), m);
                    o.z = __fadd_rn(__fmul_rn((float)n2, d), m); o.w = __fadd_rn(__fmul_rn((float)n3, d), m);
                }
            } else {
                const float d = __uint_as_float(dm[u]);
                o.x = __fmul_rn((float)(int)(int8_t)(q[u] & 0xFFu), d); o.y = __fmul_rn((float)(int)(int8_t)((q[u] >> 8) & 0xFFu), d);
                o.z = __fmul_rn((float)(int)(int8_t)((q[u] >> 16) & 0xFFu), d); o.w = __fmul_rn((float)(int)(int8_t)(q[u] >> 24), d);
            }
            reinterpret_cast<float4 *>(y)[t] = o;
        }
    }
}

// ggml_compute_forward_add_q_f32 (Ggml.cs:4797-4906) for the sibling types: group-local, so one thread rebuilds the 32 floats of
// its group in registers (dequantize exactly as above, then one float add: ggml_vec_acc_f32, Ggml.cs:2591-2594) and requantizes
// them.  The whole group is read before anything is written, so dst may alias src0 (ggml_add_inplace).
template <int TYPE>
__global__ void __launch_bounds__(128) k_add_q_sib(const uint8_t *__restrict__ q, const float *__restrict__ x, uint8_t *__restrict__ y, long long ngrp)
{
    constexpr int G = Sib<TYPE>::G;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += (long long)gridDim.x * blockDim.x) {
        const unsigned short *src = reinterpret_cast<const unsigned short *>(q + g * G);
        uint32_t w[G / 2];
#pragma unroll
        for (int i = 0; i < G / 2; i++) w[i] = src[i];
        float e[32];
        dequantize_group<TYPE>(w, e);
        const float4 *xp = reinterpret_cast<const float4 *>(x + g * GGB_QK);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = __ldg(xp + i);
            e[4 * i] = __fadd_rn(e[4 * i], v.x); e[4 * i + 1] = __fadd_rn(e[4 * i + 1], v.y);
            e[4 * i + 2] = __fadd_rn(e[4 * i + 2], v.z); e[4 * i + 3] = __fadd_rn(e[4 * i + 3], v.w);
        }
        uint32_t o[G / 2];
        quantize_group<TYPE>(e, o);
        unsigned short *dst = reinterpret_cast<unsigned short *>(y + g * G);
#pragma unroll
        for (int i = 0; i < G / 2; i++) dst[i] = (unsigned short)o[i];
    }
}

// Batched (tensor-core) path of the sibling formats: weights expanded once per call to dense fp16 [M][K] holding the
// reference's dequantized value rounded to half (the same operand precision the in-kernel Q4_0 / Q4_1 dequant feeds the
// MMAs), then the F16 tcgen05 GEMM runs on it.  One thread per 8 consecutive elements (one 16-byte store).
// (Q4_0 / Q4_1 are here too: their shapes that the TMA kernels cannot take -- K not a multiple of 128 -- use the same fallback.)
// Per-row range exponent of a quantized weight matrix (see ggb_internal.h: launch_weight_rowexp).  One warp per row; a lane visits
// every 32nd group and bounds the largest |value| the group can dequantize to from its header alone:
//   Q4_0 8|d|   Q4_1 max(|m|, |m + 15 d|)   Q4_2 8 max(|d0|, |d1|)   Q5_0 16|d|   Q5_1 max(|m|, |m + 31 d|)   Q8_0 128|d|
// 2-byte loads throughout: rows of the expansion fallback are only 2-byte aligned.
template <int TYPE>
__device__ __forceinline__ float rowexp_bound(const uint8_t *__restrict__ row, int kb, int lane)
{
    constexpr int G = TYPE == GGML_TYPE_Q4_0 ? 20 : TYPE == GGML_TYPE_Q4_1 ? 24 : TYPE == GGML_TYPE_Q4_2 ? 20 : TYPE == GGML_TYPE_Q5_0 ? 22 : TYPE == GGML_TYPE_Q5_1 ? 24 : 36;
    float s = 0.0f;
    for (int b = lane; b < kb; b += 32) {
        const unsigned short *g = reinterpret_cast<const unsigned short *>(row + (long long)b * G);
        float v;
        if (TYPE == GGML_TYPE_Q4_0 || TYPE == GGML_TYPE_Q8_0) {
            const float d = __uint_as_float((uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16));
            v = fabsf(d) * (TYPE == GGML_TYPE_Q4_0 ? 8.0f : 128.0f);
        } else if (TYPE == GGML_TYPE_Q4_1) {
            const float d = __uint_as_float((uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16));
            const float m = __uint_as_float((uint32_t)__ldg(g + 2) | ((uint32_t)__ldg(g + 3) << 16));
            v = fmaxf(fabsf(m), fabsf(fmaf(15.0f, d, m)));
        } else if (TYPE == GGML_TYPE_Q4_2) {
            v = 8.0f * fmaxf(fabsf(h_val(__ldg(g))), fabsf(h_val(__ldg(g + 5))));
        } else if (TYPE == GGML_TYPE_Q5_0) {
            v = 16.0f * fabsf(h_val(__ldg(g)));
        } else {
            const float d = h_val(__ldg(g)), m = h_val(__ldg(g + 1));
            v = fmaxf(fabsf(m), fabsf(fmaf(31.0f, d, m)));
        }
        if (v <= 3.4028234e38f) s = fmaxf(s, v);                    // NaN / infinite headers do not decide the scale of the finite ones
    }
    return s;
}
// every node of a batch in one launch: warp -> (node, row)
__global__ void __launch_bounds__(256) k_weight_rowexp(const __grid_constant__ RowExpBatch b)
{
    const int lane = threadIdx.x & 31;
    const long long grow = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (grow >= b.total_rows) return;
    int n = 0;
    { int hi = b.n_nodes - 1; while (n < hi) { const int mid = (n + hi + 1) >> 1; if (grow >= b.node[mid].row0) n = mid; else hi = mid - 1; } }   // last node with row0 <= grow
    const RowExpNode &nd = b.node[n];
    const long long r = grow - nd.row0;
    const uint8_t *row = nd.W + r * nd.nb01;
    float s;
    switch (nd.type) {
    case GGML_TYPE_Q4_0: s = rowexp_bound<GGML_TYPE_Q4_0>(row, nd.kb, lane); break;
    case GGML_TYPE_Q4_1: s = rowexp_bound<GGML_TYPE_Q4_1>(row, nd.kb, lane); break;
    case GGML_TYPE_Q4_2: s = rowexp_bound<GGML_TYPE_Q4_2>(row, nd.kb, lane); break;
    case GGML_TYPE_Q5_0: s = rowexp_bound<GGML_TYPE_Q5_0>(row, nd.kb, lane); break;
    case GGML_TYPE_Q5_1: s = rowexp_bound<GGML_TYPE_Q5_1>(row, nd.kb, lane); break;
    default:             s = rowexp_bound<GGML_TYPE_Q8_0>(row, nd.kb, lane); break;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, off));
    if (lane == 0) nd.ew[r] = range_exp(s);
}

template <int TYPE>
__global__ void __launch_bounds__(256) k_expand_f16(const uint8_t *__restrict__ W, long long nb01, __half *__restrict__ out, long long M, int K, const int *__restrict__ ew)
{
    constexpr int G = TYPE == GGML_TYPE_Q4_0 ? 20 : TYPE == GGML_TYPE_Q4_1 ? 24 : TYPE == GGML_TYPE_Q4_2 ? 20 : TYPE == GGML_TYPE_Q5_0 ? 22 : TYPE == GGML_TYPE_Q5_1 ? 24 : 36;
    const int oct_per_row = K >> 3;
    const long long total = M * oct_per_row;
    // no early launch_dependents: whatever follows (the activation kernel, or the GEMM itself when its activations were already
    // staged for another node) must not start before these weights are complete
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / oct_per_row;
        const int oc = (int)(t - row * oct_per_row), sub = oc & 3;              // elements 8*sub .. 8*sub+7 of group oc >> 2
        const unsigned short *g = reinterpret_cast<const unsigned short *>(W + row * nb01 + (long long)(oc >> 2) * G);
        const float rs = ew ? exp2i(-ew[row]) : 1.0f;                          // exact power-of-two pre-scale of this weight row
        float v[8];
        if (TYPE == GGML_TYPE_Q4_0 || TYPE == GGML_TYPE_Q4_1) {           // [f32 d][f32 m (Q4_1)][16 nibble bytes], Ggml.cs:884-911 / 961-987
            constexpr int Q0 = TYPE == GGML_TYPE_Q4_0 ? 2 : 4;
            const float d = __uint_as_float((uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16));
            const float m = TYPE == GGML_TYPE_Q4_1 ? __uint_as_float((uint32_t)__ldg(g + 2) | ((uint32_t)__ldg(g + 3) << 16)) : 0.0f;
            const uint32_t q = (uint32_t)__ldg(g + Q0 + 2 * sub) | ((uint32_t)__ldg(g + Q0 + 2 * sub + 1) << 16);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int n = (int)((q >> (4 * i)) & 15u);
                v[i] = TYPE == GGML_TYPE_Q4_0 ? __fmul_rn((float)(n - 8), d) : __fadd_rn(__fmul_rn((float)n, d), m);
            }
        } else if (TYPE == GGML_TYPE_Q4_2) {
            const unsigned short *b = g + 5 * (sub >> 1);
            const float d = h_val(__ldg(b));
            const uint32_t q = (uint32_t)__ldg(b + 1 + 2 * (sub & 1)) | ((uint32_t)__ldg(b + 2 + 2 * (sub & 1)) << 16);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn((float)((int)((q >> (4 * i)) & 15u) - 8), d);
        } else if (TYPE == GGML_TYPE_Q5_0 || TYPE == GGML_TYPE_Q5_1) {
            constexpr int Q0 = TYPE == GGML_TYPE_Q5_0 ? 3 : 4;
            const float d = h_val(__ldg(g));
            const float m = TYPE == GGML_TYPE_Q5_1 ? h_val(__ldg(g + 1)) : 0.0f;
            const uint32_t qh = ((uint32_t)__ldg(g + Q0 - 2) | ((uint32_t)__ldg(g + Q0 - 1) << 16)) >> (8 * sub);
            const uint32_t q = (uint32_t)__ldg(g + Q0 + 2 * sub) | ((uint32_t)__ldg(g + Q0 + 2 * sub + 1) << 16);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int n = (int)(((q >> (4 * i)) & 15u) | (((qh >> i) & 1u) << 4));
                v[i] = TYPE == GGML_TYPE_Q5_0 ? __fmul_rn((float)(n - 16), d) : __fadd_rn(__fmul_rn((float)n, d), m);
            }
        } else {
            const float d = __uint_as_float((uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t q = __ldg(g + 2 + 4 * sub + i);
                v[2 * i] = __fmul_rn((float)(int)(int8_t)(q & 0xFFu), d);
                v[2 * i + 1] = __fmul_rn((float)(int)(int8_t)(q >> 8), d);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] *= rs;
        uint4 o;
        __half2 h;
        h = __floats2half2_rn(v[0], v[1]); o.x = *reinterpret_cast<const uint32_t *>(&h);
        h = __floats2half2_rn(v[2], v[3]); o.y = *reinterpret_cast<const uint32_t *>(&h);
        h = __floats2half2_rn(v[4], v[5]); o.z = *reinterpret_cast<const uint32_t *>(&h);
        h = __floats2half2_rn(v[6], v[7]); o.w = *reinterpret_cast<const uint32_t *>(&h);
        reinterpret_cast<uint4 *>(out + row * (long long)K)[oc] = o;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");           // completion stays transitive along a PDL chain (see k_act_f16_dequant)
}

#define GGB_SIB_SWITCH(type, ...) \
    switch (type) { \
    case GGML_TYPE_Q4_2: { constexpr int T = GGML_TYPE_Q4_2; __VA_ARGS__; } break; \
    case GGML_TYPE_Q5_0: { constexpr int T = GGML_TYPE_Q5_0; __VA_ARGS__; } break; \
    case GGML_TYPE_Q5_1: { constexpr int T = GGML_TYPE_Q5_1; __VA_ARGS__; } break; \
    case GGML_TYPE_Q8_0: { constexpr int T = GGML_TYPE_Q8_0; __VA_ARGS__; } break; \
    default: return set_error(GGB_E_UNSUPPORTED, "type %d is not a sibling quantized format", type); }

} // namespace

int launch_quantize_rows_sib(int type, const float *src, int64_t ldx, void *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "quantize: k=%lld is not a multiple of %d (Ggml.cs:549, 611, 674)", (long long)k, GGB_QK);
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (ldx & 3)) return set_error(GGB_E_UNSUPPORTED, "quantize: source rows must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(dst) & 1) return set_error(GGB_E_UNSUPPORTED, "quantize: destination must be 2-byte aligned");
    const int kb = (int)(k / GGB_QK);
    const long long ngrp = nrows * kb;
    static const bool simple = getenv("GGB200_QUANT_SIMPLE") != nullptr;     // testing: the one-thread-per-group kernel
    if (!simple && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        const long long ntiles = (ngrp + 31) / 32;
        const unsigned gridt = (unsigned)std::min<long long>((ntiles + ST_WARPS - 1) / ST_WARPS, (long long)device_sm_count() * 3);
        GGB_SIB_SWITCH(type, {
            constexpr size_t smem = (size_t)ST_WARPS * (ST_STAGES * 32 * ST_ROW + (32 * Sib<T>::G + 15) / 16 * 16);
            static PerDeviceOnce attr_once;
            if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(k_quantize_sib_tiles<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
            k_quantize_sib_tiles<T><<<gridt, ST_WARPS * 32, smem, s>>>(src, ldx, (uint8_t *)dst, ngrp, kb);
        });
    } else {
        const unsigned grid = (unsigned)std::min<long long>((ngrp + 127) / 128, (long long)device_sm_count() * 16);
        GGB_SIB_SWITCH(type, (k_quantize_sib<T><<<grid, 128, 0, s>>>(src, ldx, (uint8_t *)dst, ngrp, kb)));
    }
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_dequantize_rows_sib(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "dequantize: k=%lld is not a multiple of %d", (long long)k, GGB_QK);
    if ((reinterpret_cast<uintptr_t>(src) & 1) || (reinterpret_cast<uintptr_t>(dst) & 15)) return set_error(GGB_E_UNSUPPORTED, "dequantize: unaligned operand");
    const long long ngrp = nrows * (k / GGB_QK);
    const unsigned grid = (unsigned)std::min<long long>((ngrp * 8 + 1023) / 1024, (long long)device_sm_count() * 8);
    GGB_SIB_SWITCH(type, (k_dequantize_sib<T><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, ngrp)));
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_add_q_f32_sib(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "add_q_f32: ne00=%lld %% 32 != 0 (Ggml.cs:4891)", (long long)k);
    if (reinterpret_cast<uintptr_t>(src1) & 15) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: src1 must be 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(dst)) & 1) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: unaligned quantized operand");
    const long long ngrp = nrows * (k / GGB_QK);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>((ngrp + 127) / 128, (long long)device_sm_count() * 16));
    cfg.blockDim = dim3(128); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGB_SIB_SWITCH(type, GGB_CUDA(cudaLaunchKernelEx(&cfg, k_add_q_sib<T>, (const uint8_t *)src0, src1, (uint8_t *)dst, ngrp)));
    count_launch();
    return GGB_OK;
}

int launch_weight_rowexp_batch(RowExpBatch &b, cudaStream_t s)
{
    long long rows = 0;
    for (int i = 0; i < b.n_nodes; i++) {
        RowExpNode &nd = b.node[i];
        if (!is_q_weight(nd.type)) return set_error(GGB_E_UNSUPPORTED, "row exponents: type %d is not a quantized weight type", nd.type);
        if ((reinterpret_cast<uintptr_t>(nd.W) | (uintptr_t)nd.nb01) & 1) return set_error(GGB_E_UNSUPPORTED, "mul_mat: weight rows must be 2-byte aligned");
        nd.row0 = rows; rows += nd.M;
    }
    b.total_rows = rows;
    if (rows <= 0) return GGB_OK;
    k_weight_rowexp<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(b);
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_weight_rowexp(int type, const void *W, int64_t nb01, int64_t M, int64_t K, int *ew, cudaStream_t s)
{
    if (M <= 0) return GGB_OK;
    if (K <= 0 || K % GGB_QK) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld %% 32 != 0 (Ggml.cs:6694)", (long long)K);
    static thread_local RowExpBatch b;
    b.n_nodes = 1;
    b.node[0] = RowExpNode{static_cast<const uint8_t *>(W), (long long)nb01, ew, 0, (int)M, (int)(K / GGB_QK), type, 0};
    return launch_weight_rowexp_batch(b, s);
}

int launch_expand_f16(int type, const void *W, int64_t nb01, __half *out, int64_t M, int64_t K, const int *ew, cudaStream_t s, bool pdl)
{
    if (M <= 0 || K <= 0) return GGB_OK;
    if (K % GGB_QK) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld %% 32 != 0 (Ggml.cs:6694)", (long long)K);
    if ((reinterpret_cast<uintptr_t>(W) | (uintptr_t)nb01) & 1) return set_error(GGB_E_UNSUPPORTED, "mul_mat: weight rows must be 2-byte aligned");
    const long long total = M * (K / 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>((total + 255) / 256, (long long)device_sm_count() * 16));
    cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    if (type == GGML_TYPE_Q4_0) GGB_CUDA(cudaLaunchKernelEx(&cfg, k_expand_f16<GGML_TYPE_Q4_0>, (const uint8_t *)W, (long long)nb01, out, (long long)M, (int)K, ew));
    else if (type == GGML_TYPE_Q4_1) GGB_CUDA(cudaLaunchKernelEx(&cfg, k_expand_f16<GGML_TYPE_Q4_1>, (const uint8_t *)W, (long long)nb01, out, (long long)M, (int)K, ew));
    else GGB_SIB_SWITCH(type, GGB_CUDA(cudaLaunchKernelEx(&cfg, k_expand_f16<T>, (const uint8_t *)W, (long long)nb01, out, (long long)M, (int)K, ew)));
    count_launch();
    return GGB_OK;
}

} // namespace ggb
