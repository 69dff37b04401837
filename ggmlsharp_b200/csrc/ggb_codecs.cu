// Row codecs on the device: the quantize_fns[] columns of the reference (Ggml.cs:219-290).
//
// Bit-exactness rules: every float operation is issued with an explicit
// round-to-nearest intrinsic so nvcc can neither contract a*b+c into an FMA nor substitute an
// approximate division; Math.Round(double) is ties-to-even = rintf(); no --use_fast_math, no -ftz.
//
// Work split: 8 lanes own one block of 32 elements (one float4 each), so a warp reads 512
// contiguous bytes per load instruction and the kernel is purely HBM-bound (roofline: 4 B read
// + 0.625 / 0.75 B written per element).
#include "ggb_internal.h"
#include "ggb_act_q8.cuh"

#include <vector>

#include <algorithm>
#include <cstdlib>

namespace ggb {

namespace {

// (byte)Math.Round(v) as .NET 8 / x64 evaluates it: ties-to-even, then cvttsd2si + truncation to 8 bits; NaN and
// values outside int32 (only reachable when 1/d overflowed to infinity on a subnormal scale) give byte 0.
__device__ __forceinline__ int cs_byte(float r) { return (r != r || fabsf(r) >= 2147483648.0f) ? 0 : (__float2int_rz(r) & 0xFF); }
__device__ __forceinline__ int rne_byte(float v) { return cs_byte(rintf(v)); }
// (byte)Math.Min(15, Math.Round(v) + 8): Math.Min propagates NaN
__device__ __forceinline__ int rne_q4_0(float v) { const float r = rintf(v) + 8.0f; return (r != r) ? 0 : (r >= 15.0f ? 15 : cs_byte(r)); }

// (amax, signed value) of the FIRST element with the largest magnitude: `if (amax < |v|)` (Ggml.cs:349)
struct FirstAbsMax { float amax, val; };
__device__ __forceinline__ FirstAbsMax first_absmax(const float4 &v, int sub /*0..7*/)
{
    FirstAbsMax r{0.0f, 0.0f};
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) if (r.amax < fabsf(e[i])) { r.amax = fabsf(e[i]); r.val = e[i]; }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, r.amax, off);
        const float ov = __shfl_xor_sync(0xffffffffu, r.val, off);
        const bool other_first = ((sub ^ off) < sub);          // the partner holds the lower element indices
        if (oa > r.amax || (oa == r.amax && other_first)) { r.amax = oa; r.val = ov; }
    }
    return r;
}

template <int TYPE>
__global__ void __launch_bounds__(256) k_quantize_rows(const float *__restrict__ x, long long ldx, uint8_t *__restrict__ y,
                                                       long long nblk, int kb)
{
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = blk < nblk;
    const long long row = live ? blk / kb : 0;
    const int col = live ? (int)(blk - row * kb) : 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = __ldg(reinterpret_cast<const float4 *>(x + row * ldx + (long long)col * GGB_QK) + sub);
    const float e[4] = {v.x, v.y, v.z, v.w};

    if (TYPE == GGML_TYPE_Q4_0) {
        // Ggml.cs:341-376
        const FirstAbsMax fm = first_absmax(v, sub);
        const float d = __fdiv_rn(fm.val, -8.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        int q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = rne_q4_0(__fmul_rn(e[i], id));
        uint32_t h = (uint32_t)(q[0] | (q[1] << 4)) & 0xFF;
        h |= ((uint32_t)(q[2] | (q[3] << 4)) & 0xFF) << 8;
        const uint32_t hn = __shfl_down_sync(0xffffffffu, h, 1);
        if (live) {
            uint32_t *out = reinterpret_cast<uint32_t *>(y + blk * 20);
            if ((sub & 1) == 0) out[1 + (sub >> 1)] = h | (hn << 16);
            if (sub == 1) out[0] = __float_as_uint(d);
        }
    } else if (TYPE == GGML_TYPE_Q4_1) {
        // Ggml.cs:494-527
        // `if (v < min) min = v; if (v > max) max = v;` in element order: on equal values (+0 / -0) the
        // earlier element stays, which decides the sign bit of the stored m.
        float mn = e[0], mx = e[0];
#pragma unroll
        for (int i = 1; i < 4; i++) { if (e[i] < mn) mn = e[i]; if (e[i] > mx) mx = e[i]; }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const float on = __shfl_xor_sync(0xffffffffu, mn, off);
            const float ox = __shfl_xor_sync(0xffffffffu, mx, off);
            const bool other_first = ((sub ^ off) < sub);
            if (on < mn || (on == mn && other_first)) mn = on;
            if (ox > mx || (ox == mx && other_first)) mx = ox;
        }
        const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        int q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = rne_byte(__fmul_rn(__fsub_rn(e[i], mn), id));   // (byte), no clamp
        uint32_t h = (uint32_t)(q[0] | (q[1] << 4)) & 0xFF;
        h |= ((uint32_t)(q[2] | (q[3] << 4)) & 0xFF) << 8;
        const uint32_t hn = __shfl_down_sync(0xffffffffu, h, 1);
        if (live) {
            uint32_t *out = reinterpret_cast<uint32_t *>(y + blk * 24);
            if ((sub & 1) == 0) out[2 + (sub >> 1)] = h | (hn << 16);
            if (sub == 1) out[0] = __float_as_uint(d);
            if (sub == 3) out[1] = __float_as_uint(mn);
        }
    } else {
        // Q8_0 / Q8_1: Ggml.cs:738-761 / 786-822 with all 32 signed quants written (defects D2-D4)
        float amax = fmaxf(fmaxf(fabsf(e[0]), fabsf(e[1])), fmaxf(fabsf(e[2]), fabsf(e[3])));
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        int q[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { q[i] = (int)(int8_t)rne_byte(__fmul_rn(e[i], id)); s += q[i]; }
        const uint32_t w = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
        if (TYPE == GGML_TYPE_Q8_0) {
            if (live) {
                uint32_t *out = reinterpret_cast<uint32_t *>(y + blk * 36);
                out[1 + sub] = w;
                if (sub == 0) out[0] = __float_as_uint(d);
            }
        } else {
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);      // lanes 0-3: sum of elements 0..15, lanes 4-7: 16..31
            if (live) {
                uint32_t *out = reinterpret_cast<uint32_t *>(y + blk * 44);
                out[3 + sub] = w;
                if (sub == 0) { out[0] = __float_as_uint(d); out[1] = __float_as_uint(__fmul_rn(d, (float)s)); }
                if (sub == 4) out[2] = __float_as_uint(__fmul_rn(d, (float)s));
            }
        }
    }
}

// Weight quantizers, one thread per block of 32 (128 B in, 20 / 24 B out): the first-max and min/max scans run in the
// reference's element order, entirely in registers; eight 128-bit loads per thread are in flight at once.
// This is the EXACT path: every corner of the reference's evaluation order (NaN skipping, +-0 ties, .NET cast semantics
// for overflowed 1/d).  The fast path below handles ordinary blocks and hands anything unusual to it.
template <int TYPE>
__device__ __noinline__ void quantize_block_q4(const float4 *__restrict__ p, uint32_t *out)
{
    float e[32];                                               // reloaded (L1-hot) so the caller keeps no stack frame
#pragma unroll
    for (int i = 0; i < 8; i++) { const float4 v = __ldg(p + i); e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w; }
    uint32_t w[4];
    if (TYPE == GGML_TYPE_Q4_0) {
        float amax = 0.0f, mx = 0.0f;                          // Ggml.cs:343-354: strict <, the first maximum wins
#pragma unroll
        for (int i = 0; i < 32; i++) { const float av = fabsf(e[i]); if (amax < av) { amax = av; mx = e[i]; } }
        const float d = __fdiv_rn(mx, -8.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int q0 = rne_q4_0(__fmul_rn(e[8 * j + 2 * b], id)), q1 = rne_q4_0(__fmul_rn(e[8 * j + 2 * b + 1], id));
                acc |= (uint32_t)((q0 | (q1 << 4)) & 0xFF) << (8 * b);
            }
            w[j] = acc;
        }
        out[0] = __float_as_uint(d);
        out[1] = w[0]; out[2] = w[1]; out[3] = w[2]; out[4] = w[3];
    } else {
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;  // Ggml.cs:494-504: FLT_MAX / -FLT_MAX, equal values keep the earlier element, NaN never wins
#pragma unroll
        for (int i = 0; i < 32; i++) { if (e[i] < mn) mn = e[i]; if (e[i] > mx) mx = e[i]; }
        const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int q0 = rne_byte(__fmul_rn(__fsub_rn(e[8 * j + 2 * b], mn), id)), q1 = rne_byte(__fmul_rn(__fsub_rn(e[8 * j + 2 * b + 1], mn), id));
                acc |= (uint32_t)((q0 | (q1 << 4)) & 0xFF) << (8 * b);
            }
            w[j] = acc;
        }
        out[0] = __float_as_uint(d); out[1] = __float_as_uint(mn);
        out[2] = w[0]; out[3] = w[1]; out[4] = w[2]; out[5] = w[3];
    }
}

// Fast path for ordinary blocks.  rintf + float->int conversion run on the quarter-rate XU pipe (two per element made the
// first version of this kernel conversion-bound at 3.4 TB/s), so rounding is done on the FP32 pipe instead: for |t| < 2^22,
// t + 1.5*2^23 (round-to-nearest-even add) IS Math.Round(t), and the integer sits in the low mantissa bits.  Nibbles are
// packed by shifted integer adds of the raw float bits; the eight 0x4B400000 biases of a word come off as one constant.
// Everything the trick cannot represent -- NaN / infinite inputs, an overflowed 1/d, a |max| tie between a positive and a
// negative element (the reference takes the first), all-zero blocks (d = -0.0f), a +-0 minimum or maximum (sign of m) -- is
// detected from the OR of the biased bit patterns or two compares and handed to the exact path: returns false.
constexpr float Q_MAGIC = 12582912.0f;                          // 1.5 * 2^23
constexpr uint32_t Q_MAGIC_BITS = 0x4B400000u;
constexpr uint32_t q_pack_bias(uint32_t per_nibble) { uint32_t s = 0; for (int i = 0; i < 8; i++) s += per_nibble << (4 * i); return s; }

template <int TYPE>
__device__ __forceinline__ bool quantize_block_q4_fast(const float (&e)[32], uint32_t *out)
{
    float smax = e[0], smin = e[0];                              // fmaxf / fminf skip NaN exactly like `amax < |v|` does
#pragma unroll
    for (int i = 1; i < 32; i++) { smax = fmaxf(smax, e[i]); smin = fminf(smin, e[i]); }
    uint32_t w[4], seen = 0;
    if (TYPE == GGML_TYPE_Q4_0) {
        const float amax = fmaxf(smax, -smin);
        const bool pos = smax == amax, neg = -smin == amax;
        if (pos == neg) return false;                            // +a and -a both present (order decides), all zero, or NaN block
        const float d = __fdiv_rn(pos ? amax : -amax, -8.0f);
        const float id = __fdiv_rn(1.0f, d);                     // d != 0 here
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const float u = __fadd_rn(__fmul_rn(e[8 * j + b], id), Q_MAGIC);        // Round(x*id) + bias
                seen |= __float_as_uint(u);
                acc += __float_as_uint(fminf(u, Q_MAGIC + 7.0f)) << (4 * b);            // Math.Min(15, . + 8)
            }
            w[j] = acc - q_pack_bias(Q_MAGIC_BITS - 8u);
        }
        if (seen & 0x30000000u) return false;                    // some |x*id| was not a small finite number
        out[0] = __float_as_uint(d);
        out[1] = w[0]; out[2] = w[1]; out[3] = w[2]; out[4] = w[3];
    } else {
        if (smin == 0.0f || smax == 0.0f || !(smin < smax)) return false;   // sign of a zero min/max depends on element order; d == 0; NaN
        const float d = __fdiv_rn(__fsub_rn(smax, smin), 15.0f);
        if (d == 0.0f) return false;
        const float id = __fdiv_rn(1.0f, d);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const float u = __fadd_rn(__fmul_rn(__fsub_rn(e[8 * j + b], smin), id), Q_MAGIC);
                seen |= __float_as_uint(u);
                acc += __float_as_uint(u) << (4 * b);
            }
            w[j] = acc - q_pack_bias(Q_MAGIC_BITS);
        }
        if ((seen & 0xFFFFFFF0u) != Q_MAGIC_BITS) return false;  // a quant outside 0..15 (or NaN / inf): exact path, (byte) cast semantics
        out[0] = __float_as_uint(d); out[1] = __float_as_uint(smin);
        out[2] = w[0]; out[3] = w[1]; out[4] = w[2]; out[5] = w[3];
    }
    return true;
}

// Data movement (profiles/README.md, round-1 codec captures): with one thread per block and direct 128-bit loads, a warp
// load instruction touches 32 different 128-byte lines and the kernel stalled on memory at ~57-67 % of the copy peak.
// So each WARP runs its own cp.async pipeline: a tile is 32 blocks = 4 KB of contiguous source, fetched by eight fully
// coalesced 16-byte-per-lane async copies into shared rows padded to 144 B; lane l then reads block l with eight LDS.128
// (row stride 36 words: conflict-free) while the next two tiles are already in flight.  No CTA-wide barrier anywhere.  The
// 20 / 24-byte outputs go through a small shared staging row so the warp stores 640 / 768 contiguous bytes.
constexpr int QT_WARPS = 4, QT_STAGES = 3, QT_ROW = 144;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int TYPE>
__global__ void __launch_bounds__(QT_WARPS * 32) k_quantize_q4_tiles(const float *__restrict__ x, long long ldx, uint8_t *__restrict__ y,
                                                                     long long nblk, int kb, int exact_only)
{
    constexpr int BS = TYPE == GGML_TYPE_Q4_0 ? 20 : 24, OW = BS / 4;
    extern __shared__ uint4 qt_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *wbase = reinterpret_cast<uint8_t *>(qt_smem) + (size_t)warp * (QT_STAGES * 32 * QT_ROW + 32 * BS);
    const uint32_t in0 = (uint32_t)__cvta_generic_to_shared(wbase);
    uint32_t *ostage = reinterpret_cast<uint32_t *>(wbase + QT_STAGES * 32 * QT_ROW);
    const long long ntiles = (nblk + 31) >> 5;
    const long long gw = (long long)blockIdx.x * QT_WARPS + warp, nw = (long long)gridDim.x * QT_WARPS;
    const bool dense = ldx == (long long)kb * GGB_QK;
    const int sub = lane & 7, g = lane >> 3;

    auto src_of = [&](long long blk) -> const float * {
        if (dense) return x + blk * GGB_QK;
        const long long row = blk / kb;
        return x + row * ldx + (blk - row * kb) * GGB_QK;
    };
    auto issue = [&](long long tile, int stage) {
        if (tile < ntiles) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const long long blk = tile * 32 + i * 4 + g;
                if (blk < nblk) cp_async16(in0 + (uint32_t)(stage * 32 * QT_ROW + (i * 4 + g) * QT_ROW + sub * 16), src_of(blk) + sub * 4);
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
        const long long blk = tile * 32 + lane;
        float e[32];
        const uint4 *row = reinterpret_cast<const uint4 *>(wbase + stage * 32 * QT_ROW + lane * QT_ROW);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = row[i];
            e[4 * i] = __uint_as_float(v.x); e[4 * i + 1] = __uint_as_float(v.y); e[4 * i + 2] = __uint_as_float(v.z); e[4 * i + 3] = __uint_as_float(v.w);
        }
        uint32_t *o = ostage + lane * OW;
        if (blk < nblk) {
            if (exact_only || !quantize_block_q4_fast<TYPE>(e, o)) quantize_block_q4<TYPE>(reinterpret_cast<const float4 *>(src_of(blk)), o);
        }
        __syncwarp();
        // 32 blocks x OW words, contiguous in y: lane-contiguous 4-byte stores
        const long long nvalid = nblk - tile * 32 < 32 ? nblk - tile * 32 : 32;
        uint32_t *yo = reinterpret_cast<uint32_t *>(y + tile * 32 * BS);
#pragma unroll
        for (int j = 0; j < OW; j++) { const int wi = j * 32 + lane; if (wi < nvalid * OW) yo[wi] = ostage[wi]; }
        __syncwarp();                                            // ostage and this input stage are free again
        stage = stage == QT_STAGES - 1 ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ggml_compute_forward_add_q_f32 (Ggml.cs:4797-4906): dst_row = quantize_row_q(dequantize_row_q(src0_row) + src1_row).  Both
// codecs work on 32-element blocks, so the whole op is block-local: one thread per block reads 20 / 24 B of src0 and 128 B of
// src1, rebuilds the 32 floats in registers (dequantize exactly as k_dequantize_rows, then one float add: ggml_vec_acc_f32,
// Ggml.cs:2591-2594) and requantizes them with the quantizer above.  dst may alias src0 (ggml_add_inplace).
template <int TYPE>
__device__ __forceinline__ void addq_block_values(const uint8_t *__restrict__ qb, const float4 *__restrict__ xp, float (&e)[32])
{
    constexpr int QOFF = TYPE == GGML_TYPE_Q4_0 ? 4 : 8;
    const float d = *reinterpret_cast<const float *>(qb);
    const float m = TYPE == GGML_TYPE_Q4_1 ? *reinterpret_cast<const float *>(qb + 4) : 0.0f;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = *reinterpret_cast<const uint32_t *>(qb + QOFF + 4 * i);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 xv = __ldg(xp + i);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int l = 4 * i + j;                                   // element l: byte l/2, low nibble if l even
            const int q = (int)((w[l >> 3] >> (4 * (l & 7))) & 15u);
            const float v = TYPE == GGML_TYPE_Q4_0 ? __fmul_rn((float)(q - 8), d) : __fadd_rn(__fmul_rn((float)q, d), m);
            e[l] = __fadd_rn(v, xs[j]);
        }
    }
}

template <int TYPE>
__device__ __noinline__ void addq_block_exact(const uint8_t *__restrict__ qb, const float4 *__restrict__ xp, uint32_t *out)
{
    float e[32];
    addq_block_values<TYPE>(qb, xp, e);
    uint32_t tmp[6];
    // the exact scan of quantize_block_q4 on register values (same code as its body, fed from e[])
    if (TYPE == GGML_TYPE_Q4_0) {
        float amax = 0.0f, mx = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i++) { const float av = fabsf(e[i]); if (amax < av) { amax = av; mx = e[i]; } }
        const float d = __fdiv_rn(mx, -8.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        tmp[0] = __float_as_uint(d);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int q0 = rne_q4_0(__fmul_rn(e[8 * j + 2 * b], id)), q1 = rne_q4_0(__fmul_rn(e[8 * j + 2 * b + 1], id));
                acc |= (uint32_t)((q0 | (q1 << 4)) & 0xFF) << (8 * b);
            }
            tmp[1 + j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 5; i++) out[i] = tmp[i];
    } else {
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
#pragma unroll
        for (int i = 0; i < 32; i++) { if (e[i] < mn) mn = e[i]; if (e[i] > mx) mx = e[i]; }
        const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        tmp[0] = __float_as_uint(d); tmp[1] = __float_as_uint(mn);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int q0 = rne_byte(__fmul_rn(__fsub_rn(e[8 * j + 2 * b], mn), id)), q1 = rne_byte(__fmul_rn(__fsub_rn(e[8 * j + 2 * b + 1], mn), id));
                acc |= (uint32_t)((q0 | (q1 << 4)) & 0xFF) << (8 * b);
            }
            tmp[2 + j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) out[i] = tmp[i];
    }
}

template <int TYPE>
__global__ void __launch_bounds__(256) k_add_q_f32(const uint8_t *__restrict__ q, const float *__restrict__ x, uint8_t *__restrict__ y, long long nblk)
{
    constexpr int BS = TYPE == GGML_TYPE_Q4_0 ? 20 : 24, OW = BS / 4;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += (long long)gridDim.x * blockDim.x) {
        const uint8_t *qb = q + blk * BS;
        const float4 *xp = reinterpret_cast<const float4 *>(x + blk * GGB_QK);
        float e[32];
        addq_block_values<TYPE>(qb, xp, e);
        uint32_t o[OW];
        uint32_t *dst = reinterpret_cast<uint32_t *>(y + blk * BS);
        if (quantize_block_q4_fast<TYPE>(e, o)) {
#pragma unroll
            for (int i = 0; i < OW; i++) dst[i] = o[i];
        } else {
            addq_block_exact<TYPE>(qb, xp, dst);                      // reads qb completely before it writes dst (same block when in place)
        }
    }
}

__global__ void __launch_bounds__(256) k_f32_to_f16_rows(const float *__restrict__ x, long long ldx, __half *__restrict__ y,
                                                         long long nrows, long long k)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * k) return;
    const long long r = i / k, c = i - r * k;
    y[i] = __float2half_rn(x[r * ldx + c]);                     // (Half)float, Ggml.cs:6370
}

// Ggml.cs:886-910 (Q4_0) and 962-987 (Q4_1).  One thread per 16 output bytes: lane sub = t & 7 of a block expands nibble
// bytes 2*sub, 2*sub+1 into one float4, so a warp store is 512 contiguous bytes; the (tiny) scale and 2-byte nibble loads of
// four independent slices are issued before any of them is used (the first version -- one dependent load pair per thread
// and 32 768 small CTAs -- sat at 41 % of the copy peak, 76 % of its stall samples on memory).
template <int TYPE>
__global__ void __launch_bounds__(256) k_dequantize_rows(const uint8_t *__restrict__ x, float *__restrict__ y, long long nblk)
{
    constexpr int BS = TYPE == GGML_TYPE_Q4_0 ? 20 : 24, QOFF = TYPE == GGML_TYPE_Q4_0 ? 4 : 8, U = 4;
    const long long total = nblk * 8;
    for (long long base = (long long)blockIdx.x * (256 * U) + threadIdx.x; base < total; base += (long long)gridDim.x * (256 * U)) {
        float d[U], m[U]; uint32_t q[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long t = base + u * 256;
            d[u] = 0.f; m[u] = 0.f; q[u] = 0;
            if (t < total) {
                const uint8_t *b = x + (t >> 3) * BS;
                d[u] = __ldg(reinterpret_cast<const float *>(b));
                if (TYPE == GGML_TYPE_Q4_1) m[u] = __ldg(reinterpret_cast<const float *>(b + 4));
                q[u] = __ldg(reinterpret_cast<const unsigned short *>(b + QOFF + 2 * (int)(t & 7)));
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long t = base + u * 256;
            if (t >= total) continue;
            const int n0 = q[u] & 15, n1 = (q[u] >> 4) & 15, n2 = (q[u] >> 8) & 15, n3 = q[u] >> 12;
            float4 o;
            if (TYPE == GGML_TYPE_Q4_0) {
                o.x = __fmul_rn((float)(n0 - 8), d[u]); o.y = __fmul_rn((float)(n1 - 8), d[u]);
                o.z = __fmul_rn((float)(n2 - 8), d[u]); o.w = __fmul_rn((float)(n3 - 8), d[u]);
            } else {                                             // product rounded, then sum rounded (two roundings)
                o.x = __fadd_rn(__fmul_rn((float)n0, d[u]), m[u]); o.y = __fadd_rn(__fmul_rn((float)n1, d[u]), m[u]);
                o.z = __fadd_rn(__fmul_rn((float)n2, d[u]), m[u]); o.w = __fadd_rn(__fmul_rn((float)n3, d[u]), m[u]);
            }
            reinterpret_cast<float4 *>(y)[t] = o;
        }
    }
}

// ---- activation staging for mul_mat (the reference's INIT phase) ----

// Q8P: quantize exactly as quantize_row_q8_0 / q8_1 do (same d, same 32 quants), stored even/odd-split so a
// 32-bit word of weight nibbles pairs with one 32-bit word of activations for dp4a, plus the block's integer
// sum (Q4_0 needs -8*sum, Q4_1 needs m*d*sum; the reference's s0+s1 is d*sum).
template <int CAP>
__global__ void __launch_bounds__(256) k_act_batch(const __grid_constant__ ActBatchT<CAP> b)
{
    // Released at once: the GEMV behind this kernel may be scheduled as soon as every CTA of this grid is running.  Its producer warps
    // stream weights (which do not depend on this kernel) while the activations are being staged; its consumer warps wait for this
    // grid to COMPLETE (griddepcontrol.wait) before they touch the workspace.  No deadlock: a dependent grid is only launched once
    // every CTA here has started, and started CTAs finish whatever the GEMV then occupies.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (b.wtype != GGML_TYPE_F16 && b.wtype != GGML_TYPE_F32) {     // every quantized weight type: src1 -> Q8 blocks (vec_dot_type)
        const int lane = threadIdx.x & 31, sub = lane & 7;
        const int blk = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
        const bool live = blk < b.total_blk;
        int n = 0;
        if (live) { int hi = b.n_nodes - 1; while (n < hi) { const int mid = (n + hi + 1) >> 1; if (blk >= b.node[mid].blk0) n = mid; else hi = mid - 1; } }   // last node with blk0 <= blk
        const ActNode &nd = b.node[n];
        const int local = live ? blk - nd.blk0 : 0;
        const int row = local / b.kb, col = local - row * b.kb;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!b.no_wait) asm volatile("griddepcontrol.wait;" ::: "memory");      // activations may be the previous kernel's output
        if (live) {
            const char *p = reinterpret_cast<const char *>(nd.x) + (long long)row * nd.ldx_bytes + (long long)col * 128 + sub * 16;
            if (b.vec16) v = *reinterpret_cast<const float4 *>(p);
            else { const float *f = reinterpret_cast<const float *>(p); v = make_float4(f[0], f[1], f[2], f[3]); }
        }
        uint32_t ev, od; float d; int s;
        q8_block_sub8(v, sub, b.wtype == GGML_TYPE_Q4_2, ev, od, d, s);
        if (live) q8_block_store(nd.out + (long long)row * b.row_bytes, b.kb, b.bps, col, sub, ev, od, d, s);
    } else {
        // F16 weights: src1 -> Half (Ggml.cs:6362-6379); F32 weights: dense copy.  One thread per 4 elements.
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int q4 = (b.K + 3) / 4;                           // float4 groups per row
        const long long total = (long long)b.total_blk * q4;    // total_blk = total rows here
        if (!b.no_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (t < total) {
            const int grow = (int)(t / q4), c4 = (int)(t - (long long)grow * q4);
            int n = 0;
            { int hi = b.n_nodes - 1; while (n < hi) { const int mid = (n + hi + 1) >> 1; if (grow >= b.node[mid].blk0) n = mid; else hi = mid - 1; } }
            const ActNode &nd = b.node[n];
            const int row = grow - nd.blk0;
            const float *src = reinterpret_cast<const float *>(reinterpret_cast<const char *>(nd.x) + (long long)row * nd.ldx_bytes) + c4 * 4;
            uint8_t *o = nd.out + (long long)row * b.row_bytes;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = c4 * 4 + i;
                if (c < b.K) {
                    if (b.wtype == GGML_TYPE_F16) reinterpret_cast<__half *>(o)[c] = __float2half_rn(src[i]);
                    else reinterpret_cast<float *>(o)[c] = src[i];
                }
            }
        }
    }
    if (b.no_wait) asm volatile("griddepcontrol.wait;" ::: "memory");      // never complete before everything ahead in the stream has
}

// Batched (tensor-core) path: activations as fp16 values the reference's dot effectively multiplies by:
// quantized weights -> d * round(x / d) of the Q8 block (same d, same quants as quantize_row_q8_x); F16 -> (Half)x.
// Rows n >= N of the padded buffer are zero.
//
// Range (quantized weights only; the reference keeps d in float32, Ggml.cs:1158, 1190-1196): every row is pre-scaled by the exact
// power of two 2^-ex[n], ex[n] = ilogb(max |x[n][:]|) - 13, before the fp16 rounding, and the GEMM epilogue multiplies it back --
// so x * 2^+-20, a 1e5 outlier row or values under fp16's 6e-5 normal limit keep their full fp16 mantissa instead of overflowing
// or going subnormal.  F16 weights are NOT rescaled: there the reference itself rounds src1 to Half (Ggml.cs:6362-6379).
//
// One CTA per row at a time (the row maximum needs the whole row), one thread per block of 32 elements (128 B in, 64 B out): the
// thread's block stays in registers between the reduction and the store, so the kernel is still a pure stream (12 B per element);
// the K permutation 0,4,1,5,2,6,3,7 that the GEMM's nibble unpack produces costs nothing.
__device__ __forceinline__ void act_load_block(const char *p, int vec16, float *e)
{
    if (vec16) {
#pragma unroll
        for (int i = 0; i < 8; i++) { const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i); e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w; }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) e[i] = reinterpret_cast<const float *>(p)[i];
    }
}
__device__ __forceinline__ float act_block_amax(const float *e)
{
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(e[i]));
    return amax;
}
// one block of 32 activations -> 64 bytes of the fp16 row; rs = 2^-ex of the row (1 for F16 weights)
__device__ __forceinline__ void act_emit_block(float *e, int wtype, int perm, float rs, __half *dst_h)
{
    if (wtype != GGML_TYPE_F16) {
        const float amax = act_block_amax(e);
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        if (id < 3.0e38f) {                                    // every sane block: |x*id| <= 127.0000x, plain ties-to-even
#pragma unroll
            for (int i = 0; i < 32; i++) e[i] = __fmul_rn(__fmul_rn(d, (float)__float2int_rn(__fmul_rn(e[i], id))), rs);
        } else {                                               // 1/d overflowed (subnormal scale): .NET cast semantics
#pragma unroll
            for (int i = 0; i < 32; i++) e[i] = __fmul_rn(__fmul_rn(d, (float)(int)(int8_t)rne_byte(__fmul_rn(e[i], id))), rs);
        }
    }
    uint32_t o[16];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        // group of 8 elements -> 4 half2; natural order (0,1)(2,3)(4,5)(6,7) or permuted (0,4)(1,5)(2,6)(3,7)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const float a = perm ? e[8 * g + t] : e[8 * g + 2 * t];
            const float b = perm ? e[8 * g + 4 + t] : e[8 * g + 2 * t + 1];
            const __half2 h = __floats2half2_rn(a, b);
            o[4 * g + t] = *reinterpret_cast<const uint32_t *>(&h);
        }
    }
    uint4 *dst = reinterpret_cast<uint4 *>(dst_h);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

__global__ void __launch_bounds__(256) k_act_f16_dequant(const __grid_constant__ ActGemmBatch b)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the GEMM's prologue and weight streaming may start now
    if (b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory"); // x may be the previous kernel's output (first launch of a batch only)
    const ActGemmNode &nd = b.node[blockIdx.y];
    const int wtype = b.wtype, perm = b.perm, N = nd.N, Npad = nd.Npad, K = nd.K, vec16 = nd.vec16;
    const float *__restrict__ x = nd.x; const long long ldx_bytes = nd.ldx_bytes; __half *__restrict__ out = nd.out;
    int *__restrict__ ex = nd.ex;
    const int kb = K / GGB_QK;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)(blockDim.x >> 5);
    const bool scaled = wtype != GGML_TYPE_F16;
    __shared__ float s_max[8];
    for (int row = blockIdx.x; row < Npad; row += gridDim.x) {
        const char *xrow = reinterpret_cast<const char *>(x) + (long long)row * ldx_bytes;
        __half *orow = out + (long long)row * K;
        float e[32];
        const bool mine = tid < kb;
        if (mine && row < N) act_load_block(xrow + (long long)tid * 128, vec16, e);
        else {
#pragma unroll
            for (int i = 0; i < 32; i++) e[i] = 0.0f;
        }
        float rs = 1.0f;
        if (scaled) {
            // ---- the row's largest magnitude -> its power-of-two exponent ----
            float am = act_block_amax(e);
            if (row < N)
                for (int col = tid + (int)blockDim.x; col < kb; col += (int)blockDim.x) {       // rows longer than 32 * blockDim elements
                    float t[32];
                    act_load_block(xrow + (long long)col * 128, vec16, t);
                    am = fmaxf(am, act_block_amax(t));
                }
#pragma unroll
            for (int off = 16; off; off >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
            __syncthreads();                                        // s_max of the previous row has been read by everyone
            if (lane == 0) s_max[warp] = am;
            __syncthreads();
            am = s_max[0];
            for (int w = 1; w < nwarps; w++) am = fmaxf(am, s_max[w]);
            const int er = range_exp(am);
            rs = exp2i(-er);
            if (tid == 0 && ex) ex[row] = er;
        }
        if (mine) act_emit_block(e, wtype, perm, rs, orow + (long long)tid * GGB_QK);
        for (int col = tid + (int)blockDim.x; col < kb; col += (int)blockDim.x) {
            if (row < N) act_load_block(xrow + (long long)col * 128, vec16, e);                  // else e is still all zero
            act_emit_block(e, wtype, perm, rs, orow + (long long)col * GGB_QK);
        }
    }
    // Completion must be transitive along the stream: kernels of one batch are launched programmatically dependent on their
    // predecessor only, so this kernel does not COMPLETE before everything ahead of it in the stream has (a consumer that
    // waits for the last kernel of the batch then sees every node's dst).  Costs nothing: the work above is already done.
    if (!b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// The same staging, software-pipelined (16-byte aligned rows; the row must fit twice in shared memory): a CTA loops over rows and
// copies row r + gridDim.x into the other of two shared-memory row buffers with cp.async while it converts row r.  In the kernel
// above a CTA's row is strictly phased -- load, reduce, convert, store -- and nothing of the NEXT row is in flight meanwhile: ncu r02
// measured 30 % of the DRAM rate (staging was 22 % of the 28-node prompt step) -- and its loads, one thread per 128-byte block, touch
// 32 lines per warp instruction.  Here the copies are coalesced 16-byte chunks, and the chunks of a block are rotated by the block
// index so that neither the asynchronous copies nor the LDS.128 of the owning thread hit one bank group.
__global__ void __launch_bounds__(256) k_act_f16_dequant_pipe(const __grid_constant__ ActGemmBatch b)
{
    extern __shared__ __align__(16) float act_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory");
    const ActGemmNode &nd = b.node[blockIdx.y];
    const int wtype = b.wtype, perm = b.perm, N = nd.N, Npad = nd.Npad, K = nd.K;
    const float *__restrict__ x = nd.x; const long long ldx_bytes = nd.ldx_bytes; __half *__restrict__ out = nd.out;
    int *__restrict__ ex = nd.ex;
    const int kb = K / GGB_QK;
    const int tid = threadIdx.x, nth = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const bool scaled = wtype != GGML_TYPE_F16;
    __shared__ float s_max[8];
    float *const buf[2] = {act_smem, act_smem + K};               // K of THIS node: both halves lie inside the launch's allocation (sized by the longest row)
    // COALESCED copies: consecutive threads take consecutive 16-byte chunks of the row (a warp instruction covers 512 contiguous bytes;
    // one thread per 128-byte block would touch 32 lines per instruction and is bound by L1 wavefronts, not DRAM), so a block's bytes
    // arrive through eight different threads and the row is complete only after a CTA barrier
    auto issue = [&](int row, float *dstrow) {
        if (row < N) {
            const char *src = reinterpret_cast<const char *>(x) + (long long)row * ldx_bytes;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dstrow);
            for (int c = tid; c < kb * 8; c += nth) {
                const int col = c >> 3, i = c & 7;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)(col * 128 + (((i + col) & 7) << 4))), "l"(src + (long long)c * 16) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_block = [&](const float *srow, int col, float *e) {
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(srow + col * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a0 + (uint32_t)(((i + col) & 7) << 4)));
            e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w;
        }
    };
    int row = blockIdx.x, cur = 0;
    issue(row, buf[0]);
    for (; row < Npad; row += gridDim.x, cur ^= 1) {
        __syncthreads();                                           // everybody has finished reading the other buffer (the previous row)
        issue(row + (int)gridDim.x, buf[cur ^ 1]);                 // the next row of this CTA, under the conversion of this one
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // this thread's copies of `row` have landed ...
        __syncthreads();                                           // ... and everybody else's
        const float *srow = buf[cur];
        __half *orow = out + (long long)row * K;
        float e[32];
        const bool mine = tid < kb;
        if (mine && row < N) load_block(srow, tid, e);
        else {
#pragma unroll
            for (int i = 0; i < 32; i++) e[i] = 0.0f;
        }
        float rs = 1.0f;
        if (scaled) {
            float am = act_block_amax(e);
            if (row < N)
                for (int col = tid + nth; col < kb; col += nth) { float t[32]; load_block(srow, col, t); am = fmaxf(am, act_block_amax(t)); }
#pragma unroll
            for (int off = 16; off; off >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
            __syncthreads();                                        // s_max of the previous row has been read by everyone
            if (lane == 0) s_max[warp] = am;
            __syncthreads();
            am = s_max[0];
            for (int w = 1; w < nwarps; w++) am = fmaxf(am, s_max[w]);
            const int er = range_exp(am);
            rs = exp2i(-er);
            if (tid == 0 && ex) ex[row] = er;
        }
        if (mine) act_emit_block(e, wtype, perm, rs, orow + (long long)tid * GGB_QK);
        for (int col = tid + nth; col < kb; col += nth) {
            if (row < N) load_block(srow, col, e);                  // else e is still all zero
            act_emit_block(e, wtype, perm, rs, orow + (long long)col * GGB_QK);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory");      // completion stays transitive along the PDL chain
}

// The same pipeline with TWO threads per 32-element block (16 elements each; the block maximum is one shuffle between neighbours):
// half the registers per thread, so twice the warps are resident to cover the latencies of the conversion phase
// (ncu r02 of the one-thread-per-block version: 79 registers, 24 warps per SM, issue slots 41 % busy at 36 % of the DRAM rate).
__global__ void __launch_bounds__(512) k_act_f16_dequant_pipe2(const __grid_constant__ ActGemmBatch b)
{
    extern __shared__ __align__(16) float act_smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory");
    const ActGemmNode &nd = b.node[blockIdx.y];
    const int wtype = b.wtype, perm = b.perm, N = nd.N, Npad = nd.Npad, K = nd.K;
    const float *__restrict__ x = nd.x; const long long ldx_bytes = nd.ldx_bytes; __half *__restrict__ out = nd.out;
    int *__restrict__ ex = nd.ex;
    const int kb = K / GGB_QK, nhalf = kb * 2;
    const int tid = threadIdx.x, nth = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const bool scaled = wtype != GGML_TYPE_F16;
    __shared__ float s_max[16];
    float *const buf[2] = {act_smem, act_smem + K};
    auto issue = [&](int row, float *dstrow) {
        if (row < N) {
            const char *src = reinterpret_cast<const char *>(x) + (long long)row * ldx_bytes;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dstrow);
            for (int c = tid; c < kb * 8; c += nth) {
                const int col = c >> 3, i = c & 7;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (uint32_t)(col * 128 + (((i + col) & 7) << 4))), "l"(src + (long long)c * 16) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // half h of block col: elements 16 h .. 16 h + 15 = chunks 4 h .. 4 h + 3
    auto load_half = [&](const float *srow, int hidx, float *e) {
        const int col = hidx >> 1, h = hidx & 1;
        const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(srow + col * 32);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a0 + (uint32_t)(((4 * h + i + col) & 7) << 4)));
            e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w;
        }
    };
    auto half_amax = [](const float *e) { float a = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) a = fmaxf(a, fabsf(e[i]));
        return a; };
    // convert and store one half block; bmax = the whole block's largest magnitude
    auto emit_half = [&](float *e, float bmax, float rs, __half *dst_h) {
        if (scaled) {
            const float d = __fdiv_rn(bmax, 127.0f);
            const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
            if (id < 3.0e38f) {
#pragma unroll
                for (int i = 0; i < 16; i++) e[i] = __fmul_rn(__fmul_rn(d, (float)__float2int_rn(__fmul_rn(e[i], id))), rs);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) e[i] = __fmul_rn(__fmul_rn(d, (float)(int)(int8_t)rne_byte(__fmul_rn(e[i], id))), rs);
            }
        }
        uint32_t o[8];
#pragma unroll
        for (int g = 0; g < 2; g++)
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float a = perm ? e[8 * g + t] : e[8 * g + 2 * t];
                const float bb = perm ? e[8 * g + 4 + t] : e[8 * g + 2 * t + 1];
                const __half2 hh = __floats2half2_rn(a, bb);
                o[4 * g + t] = *reinterpret_cast<const uint32_t *>(&hh);
            }
        uint4 *dst = reinterpret_cast<uint4 *>(dst_h);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    };
    int row = blockIdx.x, cur = 0;
    issue(row, buf[0]);
    for (; row < Npad; row += gridDim.x, cur ^= 1) {
        __syncthreads();
        issue(row + (int)gridDim.x, buf[cur ^ 1]);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const float *srow = buf[cur];
        __half *orow = out + (long long)row * K;
        float e[16];
        const bool mine = tid < nhalf;
        if (mine && row < N) load_half(srow, tid, e);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = 0.0f;
        }
        float hm = half_amax(e);
        float bmax = fmaxf(hm, __shfl_xor_sync(0xffffffffu, hm, 1));      // the neighbour owns the other half of the same block
        float rs = 1.0f;
        if (scaled) {
            float am = bmax;
            if (row < N)
                for (int hi = tid + nth; hi < nhalf; hi += nth) { float t[16]; load_half(srow, hi, t); am = fmaxf(am, half_amax(t)); }
#pragma unroll
            for (int off = 16; off; off >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
            if (lane == 0) s_max[warp] = am;
            __syncthreads();
            am = s_max[0];
            for (int w = 1; w < nwarps; w++) am = fmaxf(am, s_max[w]);
            const int er = range_exp(am);
            rs = exp2i(-er);
            if (tid == 0 && ex) ex[row] = er;
        }
        if (mine) emit_half(e, bmax, rs, orow + (long long)tid * 16);
        for (int base = nth; base < nhalf; base += nth) {             // rows longer than 16 * blockDim elements; the trip count is uniform (shuffle below)
            const int hi = base + tid;
            const bool in = hi < nhalf;
            if (in && row < N) load_half(srow, hi, e);
            hm = half_amax(e);
            bmax = fmaxf(hm, __shfl_xor_sync(0xffffffffu, hm, 1));
            if (in) emit_half(e, bmax, rs, orow + (long long)hi * 16);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!b.wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory");
}

} // namespace

size_t act_row_bytes(int wtype, int64_t K)
{
    switch (wtype) {
    case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_1: case GGML_TYPE_Q4_2: case GGML_TYPE_Q5_0: case GGML_TYPE_Q5_1: case GGML_TYPE_Q8_0:
        return align_up((size_t)(K / GGB_QK) * 40, 16);
    case GGML_TYPE_F16: return align_up((size_t)K * 2, 16);
    default: return align_up((size_t)K * 4, 16);
    }
}

int launch_quantize_rows(int type, const float *src, int64_t ldx, void *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (type == GGML_TYPE_F16) {
        const long long n = nrows * k;
        k_f32_to_f16_rows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, ldx, (__half *)dst, nrows, k);
        count_launch(); GGB_CUDA(cudaGetLastError()); return GGB_OK;
    }
    if (is_sibling_q(type)) return launch_quantize_rows_sib(type, src, ldx, dst, nrows, k, s);       // Q4_2, Q5_0, Q5_1, Q8_0
    if (reinterpret_cast<uintptr_t>(dst) & 3) return set_error(GGB_E_UNSUPPORTED, "quantize: destination must be 4-byte aligned");
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "quantize: k=%lld is not a multiple of %d (Ggml.cs:336)", (long long)k, GGB_QK);
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (ldx & 3))
        return set_error(GGB_E_UNSUPPORTED, "quantize: source rows must be 16-byte aligned");
    const int kb = (int)(k / GGB_QK);
    const long long nblk = nrows * kb;
    const unsigned grid = (unsigned)((nblk * 8 + 255) / 256);
    uint8_t *y = (uint8_t *)dst;
    static const int exact_only = getenv("GGB200_QUANT_EXACT") ? 1 : 0;      // testing: force the exact path for every block
    switch (type) {
    case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_1: {
        const int bs = type == GGML_TYPE_Q4_0 ? 20 : 24;
        const size_t smem = (size_t)QT_WARPS * (QT_STAGES * 32 * QT_ROW + 32 * bs);
        static PerDeviceOnce attr_once;
        if (attr_once.need()) {
            GGB_CUDA(cudaFuncSetAttribute(k_quantize_q4_tiles<GGML_TYPE_Q4_0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(QT_WARPS * (QT_STAGES * 32 * QT_ROW + 32 * 20))));
            GGB_CUDA(cudaFuncSetAttribute(k_quantize_q4_tiles<GGML_TYPE_Q4_1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(QT_WARPS * (QT_STAGES * 32 * QT_ROW + 32 * 24))));
        }
        const long long ntiles = (nblk + 31) / 32;
        const unsigned gridt = (unsigned)std::min<long long>((ntiles + QT_WARPS - 1) / QT_WARPS, (long long)device_sm_count() * 3);
        if (type == GGML_TYPE_Q4_0) k_quantize_q4_tiles<GGML_TYPE_Q4_0><<<gridt, QT_WARPS * 32, smem, s>>>(src, ldx, y, nblk, kb, exact_only);
        else k_quantize_q4_tiles<GGML_TYPE_Q4_1><<<gridt, QT_WARPS * 32, smem, s>>>(src, ldx, y, nblk, kb, exact_only);
        break;
    }
    case GGML_TYPE_Q8_1: k_quantize_rows<GGML_TYPE_Q8_1><<<grid, 256, 0, s>>>(src, ldx, y, nblk, kb); break;
    default: return set_error(GGB_E_UNSUPPORTED, "quantize: type %d has no quantize_row_q (Ggml.cs:219-290)", type);
    }
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_add_q_f32(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (is_sibling_q(type)) return launch_add_q_f32_sib(type, src0, src1, dst, nrows, k, s);
    if (type != GGML_TYPE_Q4_0 && type != GGML_TYPE_Q4_1) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: type %d has no codec pair (Ggml.cs:219-282)", type);
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "add_q_f32: ne00=%lld %% 32 != 0 (Ggml.cs:4891)", (long long)k);
    if (reinterpret_cast<uintptr_t>(src1) & 15) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: src1 must be 16-byte aligned");
    const long long nblk = nrows * (k / GGB_QK);
    const unsigned grid = (unsigned)std::min<long long>((nblk + 255) / 256, (long long)device_sm_count() * 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (type == GGML_TYPE_Q4_0) GGB_CUDA(cudaLaunchKernelEx(&cfg, k_add_q_f32<GGML_TYPE_Q4_0>, (const uint8_t *)src0, src1, (uint8_t *)dst, nblk));
    else GGB_CUDA(cudaLaunchKernelEx(&cfg, k_add_q_f32<GGML_TYPE_Q4_1>, (const uint8_t *)src0, src1, (uint8_t *)dst, nblk));
    count_launch();
    return GGB_OK;
}

int launch_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (is_sibling_q(type)) return launch_dequantize_rows_sib(type, src, dst, nrows, k, s);
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "dequantize: k=%lld is not a multiple of %d (Ggml.cs:839)", (long long)k, GGB_QK);
    const long long nblk = nrows * (k / GGB_QK);
    const unsigned grid = (unsigned)std::min<long long>((nblk * 8 + 1023) / 1024, (long long)device_sm_count() * 8);
    if (type == GGML_TYPE_Q4_0) k_dequantize_rows<GGML_TYPE_Q4_0><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, nblk);
    else if (type == GGML_TYPE_Q4_1) k_dequantize_rows<GGML_TYPE_Q4_1><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, nblk);
    else return set_error(GGB_E_UNSUPPORTED, "dequantize: type %d has no codec on this path", type);
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_act_batch(const ActBatch &b, cudaStream_t s, bool pdl)
{
    long long threads;
    if (is_q_weight(b.wtype)) threads = (long long)b.total_blk * 8;
    else threads = (long long)b.total_blk * ((b.K + 3) / 4);
    if (threads <= 0) return GGB_OK;
    cudaLaunchConfig_t cfg = {};
    // 128-thread CTAs at 25 registers: 4 K registers each, small enough to be scheduled beside a resident GEMV CTA (118 registers x
    // 512 threads leave 5 K of the SM's 64 K) -- the executor's second lane stages the next chunk while the current one multiplies
    cfg.gridDim = dim3((unsigned)((threads + 127) / 128));
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    {
        // ... and the same shared-memory carve-out as that GEMV: an SM is not shared by kernels whose carve-outs differ
        static PerDeviceOnce once;
        if (once.need()) {
            cudaFuncSetAttribute(k_act_batch<GGB_SMALL_BATCH_NODES>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_act_batch<GGB_MAX_BATCH_NODES>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaGetLastError();
        }
    }
    if (b.n_nodes <= GGB_SMALL_BATCH_NODES) {                     // small parameter block: see launch_gemv_batch
        static thread_local ActBatchT<GGB_SMALL_BATCH_NODES> sb;
        static_cast<ActHdr &>(sb) = static_cast<const ActHdr &>(b);
        for (int i = 0; i < b.n_nodes; i++) sb.node[i] = b.node[i];
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_batch<GGB_SMALL_BATCH_NODES>, sb));
    } else {
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_batch<GGB_MAX_BATCH_NODES>, b));
    }
    count_launch();
    return GGB_OK;
}

int launch_act_f16_dequant_batch(ActGemmBatch &b, cudaStream_t s, bool pdl)
{
    if (b.n_nodes <= 0) return GGB_OK;
    // nodes of two or three different row lengths (a Llama layer: 4096 and 11008): one launch per K, because the pipelined kernel
    // sizes its shared-memory row buffers -- and with them its occupancy -- by K.  More than that: one launch sized by the longest row.
    int n_distinct = 0;
    for (int i = 0; i < b.n_nodes; i++) {
        bool seen = false;
        for (int j = 0; j < i && !seen; j++) seen = b.node[j].K == b.node[i].K;
        if (!seen) n_distinct++;
    }
    for (int i = 1; i < b.n_nodes && n_distinct <= 3; i++)
        if (b.node[i].K != b.node[0].K) {
            static thread_local ActGemmBatch part;
            std::vector<char> done((size_t)b.n_nodes, 0);
            for (int i0 = 0; i0 < b.n_nodes; i0++) {
                if (done[(size_t)i0]) continue;
                part.n_nodes = 0; part.wtype = b.wtype; part.perm = b.perm; part.wait_prior = b.wait_prior;
                for (int j = i0; j < b.n_nodes; j++)
                    if (!done[(size_t)j] && b.node[j].K == b.node[i0].K) { part.node[part.n_nodes++] = b.node[j]; done[(size_t)j] = 1; }
                int rc = launch_act_f16_dequant_batch(part, s, pdl);
                if (rc) return rc;
            }
            return GGB_OK;
        }
    int max_rows = 0, max_kb = 0;
    for (int i = 0; i < b.n_nodes; i++) {
        ActGemmNode &nd = b.node[i];
        nd.vec16 = ((reinterpret_cast<uintptr_t>(nd.x) & 15) == 0 && (nd.ldx_bytes & 15) == 0) ? 1 : 0;
        max_rows = std::max(max_rows, nd.Npad); max_kb = std::max(max_kb, nd.K / GGB_QK);
    }
    if (max_rows <= 0 || max_kb <= 0) return GGB_OK;
    const int threads = max_kb <= 32 ? 32 : max_kb <= 64 ? 64 : max_kb <= 128 ? 128 : 256;     // one block of 32 activations per thread
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3((unsigned)threads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    // the software-pipelined kernel: every node's rows 16-byte aligned, all nodes of the launch share K (one shared-memory size),
    // two rows fit in shared memory
    bool pipe = true;
    for (int i = 0; i < b.n_nodes; i++) if (!b.node[i].vec16) pipe = false;
    const size_t smem = (size_t)2 * (size_t)max_kb * GGB_QK * 4;      // (a node with shorter rows uses the front of each half)
    static const bool no_pipe = getenv("GGB200_ACT_NO_PIPE") != nullptr;
    static const bool two_per_block = getenv("GGB200_ACT_PIPE1") == nullptr;
    if (pipe && !no_pipe && two_per_block && smem <= 200 * 1024) {
        static PerDeviceOnce once2;
        if (once2.need()) GGB_CUDA(cudaFuncSetAttribute(k_act_f16_dequant_pipe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int threads2 = max_kb <= 16 ? 32 : max_kb <= 32 ? 64 : max_kb <= 64 ? 128 : max_kb <= 128 ? 256 : 512;      // two threads per block of 32 activations
        const long long per_sm = std::max<long long>(1, std::min<long long>((long long)(220 * 1024) / (long long)(smem + 1024), 65536 / (40 * threads2)));
        const long long cap = std::max<long long>(1, (long long)device_sm_count() * per_sm / b.n_nodes);
        cfg.blockDim = dim3((unsigned)threads2);
        cfg.gridDim = dim3((unsigned)std::min<long long>(max_rows, cap), (unsigned)b.n_nodes);
        cfg.dynamicSmemBytes = smem;
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_f16_dequant_pipe2, b));
        count_launch();
        return GGB_OK;
    }
    if (pipe && !no_pipe && smem <= 200 * 1024) {
        static PerDeviceOnce once;
        if (once.need()) GGB_CUDA(cudaFuncSetAttribute(k_act_f16_dequant_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // as many CTAs as are resident at once (shared memory or registers decide), dealt to the nodes; each loops over its rows
        const long long per_sm = std::max<long long>(1, std::min<long long>((long long)(220 * 1024) / (long long)(smem + 1024), 65536 / (80 * threads)));
        const long long cap = std::max<long long>(1, (long long)device_sm_count() * per_sm / b.n_nodes);
        cfg.gridDim = dim3((unsigned)std::min<long long>(max_rows, cap), (unsigned)b.n_nodes);
        cfg.dynamicSmemBytes = smem;
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_f16_dequant_pipe, b));
        count_launch();
        return GGB_OK;
    }
    // one CTA per row; capped so that a group of many nodes still launches a bounded grid (CTAs loop over rows)
    const long long cap = std::max<long long>(1, (long long)device_sm_count() * (2048 / threads) / b.n_nodes);
    const long long grid = std::min<long long>(max_rows, cap);
    cfg.gridDim = dim3((unsigned)grid, (unsigned)b.n_nodes);
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_f16_dequant, b));
    count_launch();
    return GGB_OK;
}

int launch_act_f16_dequant(int wtype, int perm, const float *x, int64_t ldx_bytes, __half *out, int *ex, int64_t N, int64_t Npad, int64_t K, cudaStream_t s, bool wait_prior, bool pdl)
{
    static thread_local ActGemmBatch b;
    b.n_nodes = 1; b.wtype = wtype; b.perm = perm; b.wait_prior = wait_prior ? 1 : 0;
    b.node[0] = ActGemmNode{x, (long long)ldx_bytes, out, ex, (int)N, (int)Npad, (int)K, 0};
    return launch_act_f16_dequant_batch(b, s, pdl);
}

} // namespace ggb
