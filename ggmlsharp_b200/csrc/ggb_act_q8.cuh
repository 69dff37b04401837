// src1 -> Q8 blocks as the quantized vec_dot expects them (the INIT phase of ggml_compute_forward_mul_mat_q_f32,
// Ggml.cs:6610-6640: quantize_row_q8_0 at 1158-1196), in the staged layout the GEMV kernels read.
//
// ONE definition for both places that build it -- k_act_batch (ggb_codecs.cu: a separate staging launch) and the prologue of
// k_gemv_fast (ggb_gemv.cu: small decode levels quantize the row inside the GEMV) -- so that the two paths cannot differ by a bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ggb {

// (sbyte)(byte)Math.Round(v) as .NET 8 / x64 evaluates it (see cs_byte in ggb_codecs.cu): ties-to-even, cvttsd2si, truncation to
// 8 bits; NaN and values outside int32 (1/d overflowed on a subnormal scale) give 0
__device__ __forceinline__ int q8_round(float v)
{
    const float r = rintf(v);
    return (r != r || fabsf(r) >= 2147483648.0f) ? 0 : (int)(int8_t)(__float2int_rz(r) & 0xFF);
}

// Eight neighbouring lanes own one block of 32 activations; lane `sub` (0..7) holds elements 4 sub .. 4 sub + 3 in v.  Every lane
// of the warp must call this (shuffles).  Results: d = amax / 127 (all lanes); s = the block's integer sum (Q4_2: two int16 half
// sums); on EVEN sub, ev / od = the 32-bit words sub / 2 of the even-element and odd-element planes (word m = quants
// 8m, 8m+2, 8m+4, 8m+6 resp. 8m+1, ...: a word of weight nibbles pairs with one word of activations for dp4a).
__device__ __forceinline__ void q8_block_sub8(const float4 v, int sub, bool q4_2, uint32_t &ev, uint32_t &od, float &d, int &s)
{
    const float e[4] = {v.x, v.y, v.z, v.w};
    float amax = fmaxf(fmaxf(fabsf(e[0]), fabsf(e[1])), fmaxf(fabsf(e[2]), fabsf(e[3])));
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    d = __fdiv_rn(amax, 127.0f);
    const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
    int q[4];
    s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { q[i] = q8_round(__fmul_rn(e[i], id)); s += q[i]; }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);                 // lanes 0-3: sum of quants 0..15, lanes 4-7: 16..31
    {
        const int other = __shfl_xor_sync(0xffffffffu, s, 4);
        const int lo = sub < 4 ? s : other, hi = sub < 4 ? other : s;
        // Q4_2 weights: each 16-element weight block needs its own half sum -> two int16; everything else: the block sum
        s = q4_2 ? (int)(((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16)) : lo + hi;
    }
    ev = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[2] & 0xFF) << 8);
    od = (uint32_t)(q[1] & 0xFF) | ((uint32_t)(q[3] & 0xFF) << 8);
    ev |= __shfl_down_sync(0xffffffffu, ev, 1) << 16;
    od |= __shfl_down_sync(0xffffffffu, od, 1) << 16;
}

// The same block by TWO neighbouring lanes (h = lane & 1 holds elements 16 h .. 16 h + 15 in e[]): an eighth of the divisions and a third
// of the shuffles per block -- the decode program stages a row per dependency level and every microsecond of it is on the critical
// path.  Same d, same quants, same sum (a maximum and an integer sum do not depend on how they are split).  Returns this half's four
// 32-bit words: ev[0..1] / od[0..1] = words 2 h, 2 h + 1 of the two planes.  The half index h is the lane's parity.
__device__ __forceinline__ void q8_block_half16(const float (&e)[16], bool q4_2, uint32_t (&ev)[2], uint32_t (&od)[2], float &d, int &s)
{
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) amax = fmaxf(amax, fabsf(e[i]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    d = __fdiv_rn(amax, 127.0f);
    const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
    int q[16];
    s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { q[i] = q8_round(__fmul_rn(e[i], id)); s += q[i]; }
    {
        // lane h = 0 summed quants 0..15, lane h = 1 quants 16..31.  Q4_2 weights: two int16 half sums; everything else: the block sum
        const int other = __shfl_xor_sync(0xffffffffu, s, 1);
        const bool first = (threadIdx.x & 1) == 0;
        const int lo = first ? s : other, hi = first ? other : s;
        s = q4_2 ? (int)(((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16)) : lo + hi;
    }
#pragma unroll
    for (int m = 0; m < 2; m++) {
        ev[m] = (uint32_t)(q[8 * m] & 0xFF) | ((uint32_t)(q[8 * m + 2] & 0xFF) << 8) | ((uint32_t)(q[8 * m + 4] & 0xFF) << 16) | ((uint32_t)(q[8 * m + 6] & 0xFF) << 24);
        od[m] = (uint32_t)(q[8 * m + 1] & 0xFF) | ((uint32_t)(q[8 * m + 3] & 0xFF) << 8) | ((uint32_t)(q[8 * m + 5] & 0xFF) << 16) | ((uint32_t)(q[8 * m + 7] & 0xFF) << 24);
    }
}
__device__ __forceinline__ void q8_block_half16_store(uint8_t *row, int kb, int bps, int col, int h, const uint32_t (&ev)[2], const uint32_t (&od)[2], float d, int s)
{
    const int idx = bps > 1 ? (col % bps) * (kb / bps) + col / bps : col;
    *reinterpret_cast<uint2 *>(row + (long long)idx * 16 + h * 8) = make_uint2(ev[0], ev[1]);
    *reinterpret_cast<uint2 *>(row + (long long)kb * 16 + (long long)idx * 16 + h * 8) = make_uint2(od[0], od[1]);
    if (h == 0) *reinterpret_cast<int2 *>(row + (long long)kb * 32 + (long long)idx * 8) = make_int2(__float_as_int(d), s);
}

// store what q8_block_sub8 returned for block `col` of a row of kb blocks; bps = blocks per weight unit of the GEMV that reads it
// (the blocks of a unit are strided so that a lane's loads are conflict-free: staged index = (col % bps) * (kb / bps) + col / bps)
__device__ __forceinline__ void q8_block_store(uint8_t *row, int kb, int bps, int col, int sub, uint32_t ev, uint32_t od, float d, int s)
{
    const int idx = bps > 1 ? (col % bps) * (kb / bps) + col / bps : col;
    if ((sub & 1) == 0) {
        reinterpret_cast<uint32_t *>(row + (long long)idx * 16)[sub >> 1] = ev;
        reinterpret_cast<uint32_t *>(row + (long long)kb * 16 + (long long)idx * 16)[sub >> 1] = od;
    }
    if (sub == 1) *reinterpret_cast<int2 *>(row + (long long)kb * 32 + (long long)idx * 8) = make_int2(__float_as_int(d), s);
}

} // namespace ggb
