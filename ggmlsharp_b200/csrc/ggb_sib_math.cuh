// Register-level arithmetic of the sibling weight formats (Q4_2, Q5_0, Q5_1, Q8_0 -- SURVEY.md 8f-2), shared by the row codecs
// (ggb_codecs_sib.cu) and the single-token GEMV (ggb_gemv.cu).  Kept in one header, free of memory accesses, so that the
// exact same source can also be compiled for the HOST by tests/emul/ (GGB_HOST_EMUL: the handful of CUDA intrinsics used
// here are supplied as plain C++ by the test) and checked against the CPU oracle without a GPU.
//
// Reference functions restated here (Ggml.cs): quantize_row_q4_2/q5_0/q5_1/q8_0_reference_impl 547-590, 609-653, 672-714,
// 733-762; dequantize_row_q4_2/q5_0/q5_1/q8_0 992-1122; ggml_vec_dot_q4_2_q8_0 / q5_0_q8_0 / q5_1_q8_1 / q8_0_q8_0 1204-1380.
#pragma once
#ifndef GGB_HOST_EMUL
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <limits.h>
#include "../../include/ggb200.h"
#define GGB_DI __device__ __forceinline__
namespace ggb { namespace sib {
GGB_DI uint32_t h_bits(float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); }       // (Half)float, RNE, as a bit pattern
GGB_DI float h_val(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short)bits)); }
}}
#endif

namespace ggb { namespace sib {

// bytes of the weights that pair with one 32-element activation block ("group"): two Q4_2 blocks, one block of the others
template <int TYPE> struct Grp;
template <> struct Grp<GGML_TYPE_Q4_2> { static constexpr int G = 20; };
template <> struct Grp<GGML_TYPE_Q5_0> { static constexpr int G = 22; };
template <> struct Grp<GGML_TYPE_Q5_1> { static constexpr int G = 24; };
template <> struct Grp<GGML_TYPE_Q8_0> { static constexpr int G = 36; };

// ---- .NET 8 / x64 cast semantics (cvttsd2si / cvttss2si): NaN and out-of-range give the "integer indefinite" value ----
GGB_DI int cs_byte(float r) { return (r != r || fabsf(r) >= 2147483648.0f) ? 0 : (__float2int_rz(r) & 0xFF); }                              // (byte)double
// Math.Round(v) (ties to even) for |v| < 2^22 without the quarter-rate FRND / F2I conversions: v + 1.5*2^23 rounds to the nearest
// integer (RNE add) and leaves it in the low mantissa bits (the trick ggb_codecs.cu's Q4_0 fast path uses); anything larger,
// infinite or NaN takes the literal path
GGB_DI bool rne_small(float v, int &r)
{
    if (!(fabsf(v) < 4194304.0f)) return false;
    r = (int)(__float_as_uint(__fadd_rn(v, 12582912.0f)) - 0x4B400000u);
    return true;
}
GGB_DI int rne_q4(float v)                                                                                                                  // (byte)Math.Min(15, Math.Round(v) + 8)
{
    int r;
    if (rne_small(v, r)) { r += 8; return r >= 15 ? 15 : (r & 0xFF); }
    const float f = rintf(v) + 8.0f;
    return (f != f) ? 0 : (f >= 15.0f ? 15 : cs_byte(f));
}
GGB_DI int rne_byte(float v) { int r; return rne_small(v, r) ? (r & 0xFF) : cs_byte(rintf(v)); }                                              // (byte)Math.Round(v)
GGB_DI int cs_int_f(float t) { return (t != t || fabsf(t) >= 2147483648.0f) ? INT_MIN : __float2int_rz(t); }                                 // (int)float
GGB_DI uint32_t cs_uint_f(float t) { return (t != t || fabsf(t) >= 9223372036854775808.0f) ? 0u : (uint32_t)(unsigned long long)__float2ll_rz(t); }   // (uint)float

// `if (amax < |v|) { amax = |v|; max = v; }` over n elements in order (Ggml.cs:343-354, 557-568, 616-627): the signed value of the
// FIRST element of largest magnitude.  The order only matters when +a and -a are both present, so the common case is a
// max / min tree (fmaxf / fminf skip NaN exactly like the strict < does) and the sequential scan is the fallback.
template <int N_>
GGB_DI float first_signed_absmax(const float *e)
{
    float smax = e[0], smin = e[0];
#pragma unroll
    for (int l = 1; l < N_; l++) { smax = fmaxf(smax, e[l]); smin = fminf(smin, e[l]); }
    const float amax = fmaxf(smax, -smin);
    const bool pos = smax == amax, neg = -smin == amax;
    if (pos != neg) return pos ? amax : -amax;
    float am = 0.0f, mx = 0.0f;                                // tie between signs, all zero, or NaN everywhere
#pragma unroll
    for (int l = 0; l < N_; l++) { const float av = fabsf(e[l]); if (am < av) { am = av; mx = e[l]; } }
    return mx;
}

// ---- codecs: one group of 32 floats <-> Grp<TYPE>::G bytes held as 16-bit words o[] (groups are only 2-byte aligned) ----

template <int TYPE>
GGB_DI void quantize_group(const float (&e)[32], uint32_t (&o)[Grp<TYPE>::G / 2])
{
    if (TYPE == GGML_TYPE_Q4_2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float mx = first_signed_absmax<16>(e + 16 * h);     // Ggml.cs:557-568: strict <, the first maximum wins
            const float d = __fdiv_rn(mx, -8.0f);
            const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;    // from the unrounded d (Ggml.cs:575)
            o[5 * h] = h_bits(d);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t w = 0;
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const int q0 = rne_q4(__fmul_rn(e[16 * h + 4 * j + 2 * b], id)), q1 = rne_q4(__fmul_rn(e[16 * h + 4 * j + 2 * b + 1], id));
                    w |= (uint32_t)((q0 | (q1 << 4)) & 0xFF) << (8 * b);
                }
                o[5 * h + 1 + j] = w;
            }
        }
    } else if (TYPE == GGML_TYPE_Q5_0) {
        const float mx = first_signed_absmax<32>(e);           // Ggml.cs:616-627
        const float d = __fdiv_rn(mx, -16.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        uint32_t qh = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int l = 4 * j + b;
                const int t = cs_int_f(__fadd_rn(__fmul_rn(e[l], id), 16.5f));        // (int)(v + 16.5f): truncation, not Math.Round
                const uint32_t vi = (uint32_t)(t < 31 ? t : 31);                      // (uint)Math.Min(31, .)
                w |= (vi & 0x0Fu) << (4 * b);
                qh |= ((vi & 0x10u) >> 4) << l;
            }
            o[3 + j] = w;
        }
        o[0] = h_bits(d); o[1] = qh & 0xFFFFu; o[2] = qh >> 16;
    } else if (TYPE == GGML_TYPE_Q5_1) {
        // Ggml.cs:679-687: `if (v < min) min = v; if (v > max) max = v;` from +-FLT_MAX -- equal values keep the earlier element
        // (decides the sign of a zero min, which is stored), NaN never wins: a tree unless a zero or a NaN makes the order matter
        float mn = e[0], mx = e[0];
#pragma unroll
        for (int l = 1; l < 32; l++) { mn = fminf(mn, e[l]); mx = fmaxf(mx, e[l]); }
        if (!(mn < mx) || mn == 0.0f || mx == 0.0f) {
            mn = 3.402823466e+38f; mx = -3.402823466e+38f;
#pragma unroll
            for (int l = 0; l < 32; l++) { if (e[l] < mn) mn = e[l]; if (e[l] > mx) mx = e[l]; }
        }
        const float d = __fdiv_rn(__fsub_rn(mx, mn), 31.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        uint32_t qh = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int l = 4 * j + b;
                const uint32_t vi = cs_uint_f(__fadd_rn(__fmul_rn(__fsub_rn(e[l], mn), id), 0.5f));   // (uint)(v + 0.5f)
                w |= (vi & 0x0Fu) << (4 * b);
                qh |= ((vi & 0x10u) >> 4) << l;
            }
            o[4 + j] = w;
        }
        o[0] = h_bits(d); o[1] = h_bits(mn); o[2] = qh & 0xFFFFu; o[3] = qh >> 16;
    } else {
        float amax = 0.0f;                                     // Ggml.cs:738-761, all 32 signed quants (defects D2, D4)
#pragma unroll
        for (int l = 0; l < 32; l++) { const float av = fabsf(e[l]); if (amax < av) amax = av; }
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        const uint32_t db = __float_as_uint(d);
        o[0] = db & 0xFFFFu; o[1] = db >> 16;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int q0 = rne_byte(__fmul_rn(e[2 * j], id)), q1 = rne_byte(__fmul_rn(e[2 * j + 1], id));
            o[2 + j] = (uint32_t)q0 | ((uint32_t)q1 << 8);
        }
    }
}

template <int TYPE>
GGB_DI void dequantize_group(const uint32_t (&w)[Grp<TYPE>::G / 2], float (&e)[32])
{
    if (TYPE == GGML_TYPE_Q4_2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float d = h_val(w[5 * h]);
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int b = 0; b < 4; b++) e[16 * h + 4 * j + b] = __fmul_rn((float)((int)((w[5 * h + 1 + j] >> (4 * b)) & 15u) - 8), d);
        }
    } else if (TYPE == GGML_TYPE_Q5_0 || TYPE == GGML_TYPE_Q5_1) {
        constexpr int Q0 = TYPE == GGML_TYPE_Q5_0 ? 3 : 4;
        const float d = h_val(w[0]);
        const float m = TYPE == GGML_TYPE_Q5_1 ? h_val(w[1]) : 0.0f;
        const uint32_t qh = w[Q0 - 2] | (w[Q0 - 1] << 16);
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int l = 4 * j + b;
                const int q = (int)(((w[Q0 + j] >> (4 * b)) & 15u) | (((qh >> l) & 1u) << 4));
                e[l] = TYPE == GGML_TYPE_Q5_0 ? __fmul_rn((float)(q - 16), d) : __fadd_rn(__fmul_rn((float)q, d), m);
            }
    } else {
        const float d = __uint_as_float(w[0] | (w[1] << 16));
#pragma unroll
        for (int j = 0; j < 16; j++) {
            e[2 * j] = __fmul_rn((float)(int)(int8_t)(w[2 + j] & 0xFFu), d);
            e[2 * j + 1] = __fmul_rn((float)(int)(int8_t)(w[2 + j] >> 8), d);
        }
    }
}

// ---- quantized dots against one staged activation block ----
// XBlk is one Q8 activation block in the GEMV's "Q8P" form: ev = quants 0,2,4,..30, od = quants 1,3,..31 (so one 32-bit word
// of weight nibbles pairs with one word of each for dp4a), ds.x = bits of the block scale d, ds.y = sum of the 32 quants --
// or, for Q4_2 weights, the sums of quants 0..15 and 16..31 packed as two int16 (each 16-element weight block needs its own).
struct XBlk { int4 ev, od; int2 ds; };

GGB_DI int nib_isum2(const uint32_t q0, const uint32_t q1, const int ev0, const int od0, const int ev1, const int od1, int s)
{
    s = __dp4a((int)(q0 & 0x0F0F0F0Fu), ev0, s); s = __dp4a((int)((q0 >> 4) & 0x0F0F0F0Fu), od0, s);
    s = __dp4a((int)(q1 & 0x0F0F0F0Fu), ev1, s); s = __dp4a((int)((q1 >> 4) & 0x0F0F0F0Fu), od1, s);
    return s;
}

// Byte j of nibble word w holds elements 8w+2j (low nibble) and 8w+2j+1 (high nibble); their fifth bits are bits 8w+2j and
// 8w+2j+1 of qh.  hb = the 8 qh bits of this word.  (hb & 0x55) * 0x00410410 moves bit 2j to bit 8j+4 (shift 6j+4); the other
// partial products land on even positions that are never 8j+4, and the only double hits (bits 10, 16, 22) carry into odd,
// otherwise empty positions, so the mask 0x10101010 leaves exactly the four fifth bits.
GGB_DI int q5_word(const uint32_t q, const uint32_t hb, const int ev, const int od, int s)
{
    const uint32_t lo = (q & 0x0F0F0F0Fu) | (((hb & 0x55u) * 0x00410410u) & 0x10101010u);
    const uint32_t hi = ((q >> 4) & 0x0F0F0F0Fu) | ((((hb >> 1) & 0x55u) * 0x00410410u) & 0x10101010u);
    s = __dp4a((int)lo, ev, s);
    return __dp4a((int)hi, od, s);
}
GGB_DI int q5_isum(const uint32_t q0, const uint32_t q1, const uint32_t q2, const uint32_t q3, const uint32_t qh, const XBlk &x)
{
    int s = q5_word(q0, qh & 0xFFu, x.ev.x, x.od.x, 0);
    s = q5_word(q1, (qh >> 8) & 0xFFu, x.ev.y, x.od.y, s);
    s = q5_word(q2, (qh >> 16) & 0xFFu, x.ev.z, x.od.z, s);
    return q5_word(q3, qh >> 24, x.ev.w, x.od.w, s);
}

// ggml_vec_dot_q4_2_q8_0, one Q8 block = two Q4_2 blocks (Ggml.cs:1214-1251): sumf += (d0*yd)*sumi_0; sumf += (d1*yd)*sumi_1
GGB_DI float dot_q4_2(const uint32_t d0h, const uint32_t d1h, const uint32_t q0, const uint32_t q1, const uint32_t q2, const uint32_t q3,
                      const XBlk &x, float acc)
{
    const int sum_lo = (x.ds.y << 16) >> 16, sum_hi = x.ds.y >> 16;                  // sum of quants 0..15 / 16..31
    const int s0 = nib_isum2(q0, q1, x.ev.x, x.od.x, x.ev.y, x.od.y, sum_lo * -8);  // sum (q-8)*p = sum q*p - 8*sum p
    const int s1 = nib_isum2(q2, q3, x.ev.z, x.od.z, x.ev.w, x.od.w, sum_hi * -8);
    const float yd = __int_as_float(x.ds.x);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(h_val(d0h), yd), (float)s0));
    return __fadd_rn(acc, __fmul_rn(__fmul_rn(h_val(d1h), yd), (float)s1));
}
// ggml_vec_dot_q5_0_q8_0 (Ggml.cs:1270-1298): sumf += (d * sxy) * yd
GGB_DI float dot_q5_0(const uint32_t dh, const uint32_t qh, const uint32_t q0, const uint32_t q1, const uint32_t q2, const uint32_t q3,
                      const XBlk &x, float acc)
{
    const int sxy = q5_isum(q0, q1, q2, q3, qh, x) - 16 * x.ds.y;                    // sum (q5-16)*p
    return __fadd_rn(acc, __fmul_rn(__fmul_rn(h_val(dh), (float)sxy), __int_as_float(x.ds.x)));
}
// ggml_vec_dot_q5_1_q8_1 (Ggml.cs:1316-1345): sumf += (d * sxy) * yd + m * (s0 + s1); s0 + s1 = yd*sum_lo + yd*sum_hi is
// evaluated here as yd * (sum_lo + sum_hi) -- one rounding instead of three, the only deviation from the reference's order
GGB_DI float dot_q5_1(const uint32_t dmh, const uint32_t qh, const uint32_t q0, const uint32_t q1, const uint32_t q2, const uint32_t q3,
                      const XBlk &x, float acc)
{
    const int sxy = q5_isum(q0, q1, q2, q3, qh, x);
    const float yd = __int_as_float(x.ds.x);
    const float u = __fmul_rn(__fmul_rn(h_val(dmh & 0xFFFFu), (float)sxy), yd);
    const float w = __fmul_rn(h_val(dmh >> 16), __fmul_rn(yd, (float)x.ds.y));
    return __fadd_rn(acc, __fadd_rn(u, w));
}
// ggml_vec_dot_q8_0_q8_0 (Ggml.cs:1362-1377): a[] = the block's 32 int8 in memory order; even / odd bytes are separated with PRMT
GGB_DI float dot_q8_0(const float xd, const uint32_t (&a)[8], const XBlk &x, float acc)
{
    int s = 0;
    s = __dp4a((int)__byte_perm(a[0], a[1], 0x6420), x.ev.x, s); s = __dp4a((int)__byte_perm(a[0], a[1], 0x7531), x.od.x, s);
    s = __dp4a((int)__byte_perm(a[2], a[3], 0x6420), x.ev.y, s); s = __dp4a((int)__byte_perm(a[2], a[3], 0x7531), x.od.y, s);
    s = __dp4a((int)__byte_perm(a[4], a[5], 0x6420), x.ev.z, s); s = __dp4a((int)__byte_perm(a[4], a[5], 0x7531), x.od.z, s);
    s = __dp4a((int)__byte_perm(a[6], a[7], 0x6420), x.ev.w, s); s = __dp4a((int)__byte_perm(a[6], a[7], 0x7531), x.od.w, s);
    return __fadd_rn(acc, __fmul_rn(__fmul_rn(xd, __int_as_float(x.ds.x)), (float)s));
}

// One group whose first byte sits at bit 0 of w[0] (word-aligned start) -- or, for Q5_0 only, at bit 16 of w[0] (ODD: every
// second 22-byte block of a row starts in the upper half of a word).
template <int TYPE, bool ODD = false>
GGB_DI float dot_group_words(const uint32_t *w, const XBlk &x, float acc)
{
    if (TYPE == GGML_TYPE_Q4_2) {          // [d0 | qs0 0-1] [qs0 2-5] [qs0 6-7 | d1] [qs1 0-3] [qs1 4-7]
        return dot_q4_2(w[0] & 0xFFFFu, w[2] >> 16, __funnelshift_r(w[0], w[1], 16), __funnelshift_r(w[1], w[2], 16), w[3], w[4], x, acc);
    } else if (TYPE == GGML_TYPE_Q5_0) {
        if (ODD)                           // [.. | d] [qh] [qs 0-3] [qs 4-7] [qs 8-11] [qs 12-15]
            return dot_q5_0(w[0] >> 16, w[1], w[2], w[3], w[4], w[5], x, acc);
        // [d | qh lo] [qh hi | qs 0-1] [qs 2-5] [qs 6-9] [qs 10-13] [qs 14-15 | ..]
        return dot_q5_0(w[0] & 0xFFFFu, __funnelshift_r(w[0], w[1], 16), __funnelshift_r(w[1], w[2], 16), __funnelshift_r(w[2], w[3], 16),
                        __funnelshift_r(w[3], w[4], 16), __funnelshift_r(w[4], w[5], 16), x, acc);
    } else if (TYPE == GGML_TYPE_Q5_1) {   // [d | m] [qh] [qs x 4]
        return dot_q5_1(w[0], w[1], w[2], w[3], w[4], w[5], x, acc);
    } else {                               // [f32 d] [32 x int8]
        const uint32_t a[8] = {w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8]};
        return dot_q8_0(__uint_as_float(w[0]), a, x, acc);
    }
}

// One GEMV "unit" = the smallest run of whole groups a lane can fetch with aligned vector loads, as the words v[] it loaded:
// Q4_2 80 B = 4 groups, Q5_0 88 B = 4 groups (groups 1 and 3 start mid-word), Q5_1 48 B = 2 groups, Q8_0 144 B = 4 groups.
template <int TYPE> struct Unit;
template <> struct Unit<GGML_TYPE_Q4_2> { static constexpr int BYTES = 80, GROUPS = 4; };
template <> struct Unit<GGML_TYPE_Q5_0> { static constexpr int BYTES = 88, GROUPS = 4; };
template <> struct Unit<GGML_TYPE_Q5_1> { static constexpr int BYTES = 48, GROUPS = 2; };
template <> struct Unit<GGML_TYPE_Q8_0> { static constexpr int BYTES = 144, GROUPS = 4; };

template <int TYPE>
GGB_DI float dot_unit_words(const uint32_t *v, const XBlk (&x)[4], float acc)
{
    if (TYPE == GGML_TYPE_Q4_2) {
#pragma unroll
        for (int j = 0; j < 4; j++) acc = dot_group_words<GGML_TYPE_Q4_2>(v + 5 * j, x[j], acc);
        return acc;
    } else if (TYPE == GGML_TYPE_Q5_0) {   // groups at bytes 0, 22, 44, 66 = words 0, 5.5, 11, 16.5
        acc = dot_group_words<GGML_TYPE_Q5_0, false>(v, x[0], acc);
        acc = dot_group_words<GGML_TYPE_Q5_0, true>(v + 5, x[1], acc);
        acc = dot_group_words<GGML_TYPE_Q5_0, false>(v + 11, x[2], acc);
        return dot_group_words<GGML_TYPE_Q5_0, true>(v + 16, x[3], acc);
    } else if (TYPE == GGML_TYPE_Q5_1) {
        acc = dot_group_words<GGML_TYPE_Q5_1>(v, x[0], acc);
        return dot_group_words<GGML_TYPE_Q5_1>(v + 6, x[1], acc);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) acc = dot_group_words<GGML_TYPE_Q8_0>(v + 9 * j, x[j], acc);
        return acc;
    }
}

}} // namespace ggb::sib
