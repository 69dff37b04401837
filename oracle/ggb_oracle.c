/*
 * ggb_oracle.c -- CPU restatement of GGMLSharp's mul_mat hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The product (ggmlsharp_b200/csrc) never links, imports or calls it.
 *
 * Each function restates one function of /root/reference/GGMLSharp/Ggml.cs
 * ("Ggml.cs" below) in plain C99, with the reference's operation order, operand
 * widths and rounding:
 *   - C# float arithmetic on .NET 8 / x64 is strict IEEE binary32 per operation
 *     (SSE scalar), never contracted to FMA  -> build with -ffp-contract=off.
 *   - Math.Round(double) is MidpointRounding.ToEven         -> nearbyint().
 *   - (Half)float is IEEE round-to-nearest-even              -> f32_to_f16().
 *
 * Repairs.  The reference's quantized mul_mat cannot run as written (SURVEY.md
 * Appendix A, defects D1-D6).  The functions below implement the evident intent:
 *   D1  quantize_fns[] is looked up by TYPE (Q4_0 -> q8_0 / vec_dot_q4_0_q8_0,
 *       Q4_1 -> q8_1 / vec_dot_q4_1_q8_1), as the per-entry comments
 *       (Ggml.cs:221,230) say, not by array position.
 *   D2  quantize_row_q8_0 writes all 32 quants (Ggml.cs:756 steps by 2).
 *   D3  quantize_row_q8_1 pairs l with 16+l for l in 0..15 (Ggml.cs:808).
 *   D4  Q8 quants are int8 (TypeDefinitions.cs:281,289 declare byte).
 *   D5/D6  the scalar quantize_row_q4_0_reference_impl / scalar dequantize
 *       branches are the targets; the AVX branches are defective.
 * Sibling formats (Q4_2, Q5_0, Q5_1, Q8_0 as weights) add one more:
 *   D9  fp16 block scales declared `ushort` (block_q4_2.d, block_q5_1.d/.m,
 *       TypeDefinitions.cs:249-275) are stored with a NUMERIC cast
 *       `(ushort)(Half)d` (Ggml.cs:577, 678-679) and read back with numeric
 *       `(float)(Half)x.d` / `(float)x.d` (Ggml.cs:1003, 1068-1069, 1220-1221,
 *       1320-1321): a scale of 0.01 is stored as integer 0.  Upstream's
 *       GGML_FP32_TO_FP16 / GGML_FP16_TO_FP32 move the IEEE binary16 BIT PATTERN,
 *       which is what block_q5_0 (declared `Half d`, TypeDefinitions.cs:263) does
 *       in this very file; the oracle stores and loads bit patterns for all three.
 *       Q8_0 weights read their quants as int8 (D4).
 *
 * PARITY PINNING.  F32 mul_mat and the tensor stride rule are pinned by the
 * reference's own Test3 (generator + expected solution) and Test0 (strides);
 * see tests/test_oracle_pins.py.  Quantize / dequantize / Q4_0 / Q4_1 / F16
 * mul_mat are PARITY UNPINNED: the reference holds no golden vector, known-answer
 * test or fixture for them and no .NET toolchain exists here to run it.  The
 * known-answer vectors in tests/golden/ were derived by hand from the source.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

#define QK 32

/* TypeDefinitions.cs:236-248, 277-290 -- float32 scales, 16 nibble bytes. */
typedef struct { float d; uint8_t qs[QK / 2]; } block_q4_0;            /* 20 B */
typedef struct { float d; float m; uint8_t qs[QK / 2]; } block_q4_1;   /* 24 B */
typedef struct { float d; int8_t qs[QK]; } block_q8_0;                 /* 36 B, D4: int8 */
typedef struct { float d; float s0; float s1; int8_t qs[QK]; } block_q8_1; /* 44 B */

/* TypeDefinitions.cs:249-276 -- the sibling weight formats: fp16 scales, 16- or 32-element blocks, a 5th-bit plane qh. */
#define QK4_2 16
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[QK4_2 / 2]; } block_q4_2;                   /* 10 B */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[QK / 2]; } block_q5_0;       /* 22 B */
typedef struct { uint16_t d; uint16_t m; uint8_t qh[4]; uint8_t qs[QK / 2]; } block_q5_1; /* 24 B */
#pragma pack(pop)
_Static_assert(sizeof(block_q4_2) == 10, "block_q4_2");
_Static_assert(sizeof(block_q5_0) == 22, "block_q5_0");
_Static_assert(sizeof(block_q5_1) == 24, "block_q5_1");
_Static_assert(sizeof(block_q4_0) == 20, "block_q4_0");
_Static_assert(sizeof(block_q4_1) == 24, "block_q4_1");
_Static_assert(sizeof(block_q8_0) == 36, "block_q8_0");
_Static_assert(sizeof(block_q8_1) == 44, "block_q8_1");

/* TypeDefinitions.cs:153-169 */
enum { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q4_1 = 3, T_Q4_2 = 4, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8, T_Q8_1 = 9 };

/* ---- IEEE binary16 <-> binary32, what .NET's System.Half conversions do ---- */

uint16_t orc_f32_to_f16(float f)
{
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u)                       /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? 0x0200u | ((ax >> 13) & 0x3ffu) : 0));
    if (ax >= 0x477ff000u)                       /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    if (ax < 0x33000001u)                        /* < 2^-25 (or == 2^-25, tie to even 0) */
        return (uint16_t)sign;
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;   /* 24-bit significand */
    int shift;                                   /* bits to drop */
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }          /* subnormal half */
    else         { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;           /* ties to even; carry propagates into exponent */
    return (uint16_t)(sign | (base + q));
}

float orc_f16_to_f32(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int s = 0; while (!(m & 0x400u)) { m <<= 1; s++; }
               x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | ((m & 0x3ffu) << 13); }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}

/* ---- row quantizers ---- */

/* C# `(byte)someDouble` on .NET 8 / x64 is cvttsd2si + truncation: NaN and values outside int32 give the
 * "integer indefinite" 0x80000000, whose low byte is 0.  Only reachable when a block's scale is so small
 * (subnormal) that 1/d overflows to infinity; kept so that even those blocks are bit-defined. */
static inline uint8_t cs_byte(double t)
{
    if (t != t || t >= 2147483648.0 || t < -2147483648.0) return 0;
    return (uint8_t)(int32_t)t;
}
/* Math.Min(double, double) propagates NaN (C fmin does not). */
static inline double cs_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }

/* Ggml.cs:334-377 quantize_row_q4_0_reference_impl (the bit-exactness target). */
void orc_quantize_row_q4_0(const float *x, void *vy, int k)
{
    block_q4_0 *y = (block_q4_0 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }    /* strict <: first max wins */
        }
        const float d = max / -8.0f;                             /* all-zero block -> -0.0f */
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = d;
        for (int l = 0; l < QK; l += 2) {
            const float v0 = x[i * QK + l + 0] * id;
            const float v1 = x[i * QK + l + 1] * id;
            const uint8_t vi0 = cs_byte(cs_min(15.0, nearbyint((double)v0) + 8.0));
            const uint8_t vi1 = cs_byte(cs_min(15.0, nearbyint((double)v1) + 8.0));
            y[i].qs[l / 2] = (uint8_t)(vi0 | (vi1 << 4));
        }
    }
}

/* Ggml.cs:487-528 quantize_row_q4_1_reference_impl (530-539 just calls it). */
void orc_quantize_row_q4_1(const float *x, void *vy, int k)
{
    block_q4_1 *y = (block_q4_1 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float min = FLT_MAX, max = -FLT_MAX;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (v < min) min = v;
            if (v > max) max = v;
        }
        const float d = (max - min) / 15.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = d;
        y[i].m = min;
        for (int l = 0; l < QK; l += 2) {
            const float v0 = (x[i * QK + l + 0] - min) * id;
            const float v1 = (x[i * QK + l + 1] - min) * id;
            const uint8_t vi0 = cs_byte(nearbyint((double)v0));   /* no clamp */
            const uint8_t vi1 = cs_byte(nearbyint((double)v1));
            y[i].qs[l / 2] = (uint8_t)(vi0 | (vi1 << 4));
        }
    }
}

/* Ggml.cs:733-762 quantize_row_q8_0_reference_impl, D2 + D4 repaired. */
void orc_quantize_row_q8_0(const float *x, void *vy, int k)
{
    block_q8_0 *y = (block_q8_0 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (amax < fabsf(v)) amax = fabsf(v);
        }
        const float d = amax / 127.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = d;
        for (int l = 0; l < QK; l++) {
            const float v0 = x[i * QK + l] * id;
            y[i].qs[l] = (int8_t)cs_byte(nearbyint((double)v0));
        }
    }
}

/* Ggml.cs:781-823 quantize_row_q8_1_reference_impl, D3 + D4 repaired. */
void orc_quantize_row_q8_1(const float *x, void *vy, int k)
{
    block_q8_1 *y = (block_q8_1 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (amax < fabsf(v)) amax = fabsf(v);
        }
        const float d = amax / 127.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = d;
        int sum0 = 0, sum1 = 0;
        for (int l = 0; l < QK / 2; l++) {
            const float v0 = x[i * QK + l] * id;
            const float v1 = x[i * QK + QK / 2 + l] * id;
            y[i].qs[l] = (int8_t)cs_byte(nearbyint((double)v0));
            y[i].qs[QK / 2 + l] = (int8_t)cs_byte(nearbyint((double)v1));
            sum0 += y[i].qs[l];
            sum1 += y[i].qs[QK / 2 + l];
        }
        y[i].s0 = d * (float)sum0;
        y[i].s1 = d * (float)sum1;
    }
}

/* ---- row dequantizers ---- */

/* Ggml.cs:884-911, scalar branch (the AVX2 branch is defect D6). */
void orc_dequantize_row_q4_0(const void *vx, float *y, int k)
{
    const block_q4_0 *x = (const block_q4_0 *)vx;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        const float d = x[i].d;
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = x[i].qs[l / 2];
            const int vi0 = vi & 0x0F, vi1 = vi >> 4;
            y[i * QK + l + 0] = (float)(vi0 - 8) * d;
            y[i * QK + l + 1] = (float)(vi1 - 8) * d;
        }
    }
}

/* Ggml.cs:961-987 (the AVX2 branch 922-957 computes the same mul-then-add). */
void orc_dequantize_row_q4_1(const void *vx, float *y, int k)
{
    const block_q4_1 *x = (const block_q4_1 *)vx;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        const float d = x[i].d, m = x[i].m;
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = x[i].qs[l / 2];
            const float p0 = (float)(vi & 0x0F) * d;      /* rounded product ... */
            const float p1 = (float)(vi >> 4) * d;
            y[i * QK + l + 0] = p0 + m;                   /* ... then rounded sum */
            y[i * QK + l + 1] = p1 + m;
        }
    }
}

/* ---- dot products ---- */

/* Ggml.cs:2631-2640: float product, double accumulator, sequential. */
void orc_vec_dot_f32(int n, float *s, const float *x, const float *y)
{
    double sumf = 0.0;
    for (int i = 0; i < n; ++i) { const float p = x[i] * y[i]; sumf += (double)p; }
    *s = (float)sumf;
}

/* Ggml.cs:2642-2651: (float)Half * (float)Half in float, double accumulator. */
void orc_vec_dot_f16(int n, float *s, const uint16_t *x, const uint16_t *y)
{
    double sumf = 0.0;
    for (int i = 0; i < n; ++i) {
        const float p = orc_f16_to_f32(x[i]) * orc_f16_to_f32(y[i]);
        sumf += (double)p;
    }
    *s = (float)sumf;
}

/* Ggml.cs:1124-1162, q8 read as signed (D4). */
void orc_vec_dot_q4_0_q8_0(int n, float *s, const void *vx, const void *vy)
{
    const block_q4_0 *x = (const block_q4_0 *)vx;
    const block_q8_0 *y = (const block_q8_0 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        const float d0 = x[i].d, d1 = y[i].d;
        int sumi = 0;
        for (int j = 0; j < QK / 2; j++) {
            const uint8_t v0 = x[i].qs[j];
            const int i0 = (v0 & 0x0F) - 8, i1 = (v0 >> 4) - 8;
            const int i2 = y[i].qs[2 * j + 0], i3 = y[i].qs[2 * j + 1];
            sumi += i0 * i2 + i1 * i3;
        }
        const float dd = d0 * d1;              /* d0 * d1 * sumi == (d0*d1)*(float)sumi */
        sumf += dd * (float)sumi;
    }
    *s = sumf;
}

/* Ggml.cs:1164-1201, q8 read as signed (D4). */
void orc_vec_dot_q4_1_q8_1(int n, float *s, const void *vx, const void *vy)
{
    const block_q4_1 *x = (const block_q4_1 *)vx;
    const block_q8_1 *y = (const block_q8_1 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        const float d0 = x[i].d, m0 = x[i].m, d1 = y[i].d;
        for (int j = 0; j < QK / 2; j++) {
            const uint8_t v0 = x[i].qs[j];
            const float a0 = d0 * (float)(v0 & 0x0F);
            const float a1 = d0 * (float)(v0 >> 4);
            const float f0 = a0 + m0, f1 = a1 + m0;
            const float f2 = d1 * (float)y[i].qs[2 * j + 0];
            const float f3 = d1 * (float)y[i].qs[2 * j + 1];
            const float p0 = f0 * f2, p1 = f1 * f3;
            const float p = p0 + p1;
            sumf += p;
        }
    }
    *s = sumf;
}

/* ---- sibling weight formats: Q4_2, Q5_0, Q5_1, and Q8_0 as a weight type ---- */

/* C# `(int)someFloat` on .NET 8 / x64 is cvttss2si: NaN and out-of-range give 0x80000000. */
static inline int32_t cs_int_f(float t)
{
    if (t != t || t >= 2147483648.0f || t < -2147483648.0f) return INT32_MIN;
    return (int32_t)t;
}
/* C# `(uint)someFloat` on .NET 8 / x64 is cvttss2si r64 + truncation to 32 bits: NaN and |t| >= 2^63 give 0. */
static inline uint32_t cs_uint_f(float t)
{
    if (t != t || t >= 9223372036854775808.0f || t < -9223372036854775808.0f) return 0u;
    return (uint32_t)(uint64_t)(int64_t)t;
}

/* Ggml.cs:547-590 quantize_row_q4_2_reference_impl (593-601 just calls it); D9: d stored as fp16 bits. */
void orc_quantize_row_q4_2(const float *x, void *vy, int k)
{
    block_q4_2 *y = (block_q4_2 *)vy;
    const int nb = k / QK4_2;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int l = 0; l < QK4_2; l++) {
            const float v = x[i * QK4_2 + l];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -8.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;       /* from the UNROUNDED d (Ggml.cs:575) */
        y[i].d = orc_f32_to_f16(d);
        for (int l = 0; l < QK4_2; l += 2) {
            const float v0 = x[i * QK4_2 + l + 0] * id;
            const float v1 = x[i * QK4_2 + l + 1] * id;
            const uint8_t vi0 = cs_byte(cs_min(15.0, nearbyint((double)v0) + 8.0));
            const uint8_t vi1 = cs_byte(cs_min(15.0, nearbyint((double)v1) + 8.0));
            y[i].qs[l / 2] = (uint8_t)(vi0 | (vi1 << 4));
        }
    }
}

/* Ggml.cs:609-653 quantize_row_q5_0_reference_impl: d = max / -16, q = min(31, (int)(x*id + 16.5f)) -- truncation of a
 * float sum, NOT Math.Round; bit 4 of each quant goes to qh at the element's index. */
void orc_quantize_row_q5_0(const float *x, void *vy, int k)
{
    block_q5_0 *y = (block_q5_0 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -16.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = orc_f32_to_f16(d);
        uint32_t qh = 0;
        for (int l = 0; l < QK; l += 2) {
            const float v0 = x[i * QK + l + 0] * id;
            const float v1 = x[i * QK + l + 1] * id;
            const int32_t t0 = cs_int_f(v0 + 16.5f), t1 = cs_int_f(v1 + 16.5f);
            const uint32_t vi0 = (uint32_t)(t0 < 31 ? t0 : 31);     /* Math.Min(31, int), then (uint) */
            const uint32_t vi1 = (uint32_t)(t1 < 31 ? t1 : 31);
            y[i].qs[l / 2] = (uint8_t)((vi0 & 0x0F) | ((vi1 & 0x0F) << 4));
            qh |= ((vi0 & 0x10) >> 4) << (l + 0);
            qh |= ((vi1 & 0x10) >> 4) << (l + 1);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}

/* Ggml.cs:672-714 quantize_row_q5_1_reference_impl: d = (max-min)/31, q = (uint)((x-min)*id + 0.5f); D9 for d and m. */
void orc_quantize_row_q5_1(const float *x, void *vy, int k)
{
    block_q5_1 *y = (block_q5_1 *)vy;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        float min = FLT_MAX, max = -FLT_MAX;               /* float.MaxValue / float.MinValue */
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (v < min) min = v;
            if (v > max) max = v;
        }
        const float d = (max - min) / 31.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        y[i].d = orc_f32_to_f16(d);
        y[i].m = orc_f32_to_f16(min);
        uint32_t qh = 0;
        for (int l = 0; l < QK; l += 2) {
            const float v0 = (x[i * QK + l + 0] - min) * id;
            const float v1 = (x[i * QK + l + 1] - min) * id;
            const uint32_t vi0 = cs_uint_f(v0 + 0.5f), vi1 = cs_uint_f(v1 + 0.5f);
            y[i].qs[l / 2] = (uint8_t)((vi0 & 0x0F) | ((vi1 & 0x0F) << 4));
            qh |= ((vi0 & 0x10) >> 4) << (l + 0);
            qh |= ((vi1 & 0x10) >> 4) << (l + 1);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}

/* Ggml.cs:992-1022 dequantize_row_q4_2 (D9: d read as fp16 bits). */
void orc_dequantize_row_q4_2(const void *vx, float *y, int k)
{
    const block_q4_2 *x = (const block_q4_2 *)vx;
    const int nb = k / QK4_2;
    for (int i = 0; i < nb; i++) {
        const float d = orc_f16_to_f32(x[i].d);
        for (int l = 0; l < QK4_2; l += 2) {
            const uint8_t vi = x[i].qs[l / 2];
            y[i * QK4_2 + l + 0] = (float)((vi & 0x0F) - 8) * d;
            y[i * QK4_2 + l + 1] = (float)((vi >> 4) - 8) * d;
        }
    }
}

/* Ggml.cs:1025-1061 dequantize_row_q5_0. */
void orc_dequantize_row_q5_0(const void *vx, float *y, int k)
{
    const block_q5_0 *x = (const block_q5_0 *)vx;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        const float d = orc_f16_to_f32(x[i].d);
        uint32_t qh; memcpy(&qh, x[i].qh, 4);
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = x[i].qs[l / 2];
            const int vh0 = (int)((qh >> (l + 0)) & 1u) << 4, vh1 = (int)((qh >> (l + 1)) & 1u) << 4;
            const int vi0 = (vi & 0x0F) | vh0, vi1 = (vi >> 4) | vh1;
            y[i * QK + l + 0] = (float)(vi0 - 16) * d;
            y[i * QK + l + 1] = (float)(vi1 - 16) * d;
        }
    }
}

/* Ggml.cs:1064-1101 dequantize_row_q5_1 (D9): q*d rounded, then + m rounded. */
void orc_dequantize_row_q5_1(const void *vx, float *y, int k)
{
    const block_q5_1 *x = (const block_q5_1 *)vx;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++) {
        const float d = orc_f16_to_f32(x[i].d), m = orc_f16_to_f32(x[i].m);
        uint32_t qh; memcpy(&qh, x[i].qh, 4);
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = x[i].qs[l / 2];
            const int vh0 = (int)((qh >> (l + 0)) & 1u) << 4, vh1 = (int)((qh >> (l + 1)) & 1u) << 4;
            const float p0 = (float)((vi & 0x0F) | vh0) * d;
            const float p1 = (float)((vi >> 4) | vh1) * d;
            y[i * QK + l + 0] = p0 + m;
            y[i * QK + l + 1] = p1 + m;
        }
    }
}

/* Ggml.cs:1104-1122 dequantize_row_q8_0, quants signed (D4). */
void orc_dequantize_row_q8_0(const void *vx, float *y, int k)
{
    const block_q8_0 *x = (const block_q8_0 *)vx;
    const int nb = k / QK;
    for (int i = 0; i < nb; i++)
        for (int l = 0; l < QK; l++) y[i * QK + l] = (float)x[i].qs[l] * x[i].d;
}

/* Ggml.cs:1204-1254: two 16-element Q4_2 blocks against one Q8_0 block; q8 signed (D4), scales as fp16 bits (D9). */
void orc_vec_dot_q4_2_q8_0(int n, float *s, const void *vx, const void *vy)
{
    const block_q4_2 *x = (const block_q4_2 *)vx;
    const block_q8_0 *y = (const block_q8_0 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        const uint8_t *x0 = x[2 * i + 0].qs, *x1 = x[2 * i + 1].qs;
        const int8_t *y0 = y[i].qs;
        const float d0 = orc_f16_to_f32(x[2 * i + 0].d), d1 = orc_f16_to_f32(x[2 * i + 1].d);
        int sumi_0 = 0, sumi_1 = 0;
        for (int j = 0; j < QK / 4; j++) {
            const uint8_t v0 = x0[j], v1 = x1[j];
            const int i0_0 = (v0 & 0x0F) - 8, i1_0 = (v0 >> 4) - 8;
            const int i0_1 = (v1 & 0x0F) - 8, i1_1 = (v1 >> 4) - 8;
            const int i2_0 = y0[2 * j + 0], i3_0 = y0[2 * j + 1];
            const int i2_1 = y0[2 * (j + QK / 4) + 0], i3_1 = y0[2 * (j + QK / 4) + 1];
            sumi_0 += i0_0 * i2_0 + i1_0 * i3_0;
            sumi_1 += i0_1 * i2_1 + i1_1 * i3_1;
        }
        const float a = d0 * y[i].d, b = d1 * y[i].d;
        sumf += a * (float)sumi_0;
        sumf += b * (float)sumi_1;
    }
    *s = sumf;
}

/* Ggml.cs:1258-1301: sumf += (d * sxy) * y.d. */
void orc_vec_dot_q5_0_q8_0(int n, float *s, const void *vx, const void *vy)
{
    const block_q5_0 *x = (const block_q5_0 *)vx;
    const block_q8_0 *y = (const block_q8_0 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        uint32_t qh; memcpy(&qh, x[i].qh, 4);
        const float d = orc_f16_to_f32(x[i].d);
        int sxy = 0;
        for (int j = 0; j < QK / 2; j++) {
            const uint8_t v0 = x[i].qs[j];
            const int h0 = (int)((qh >> (2 * j + 0)) & 1u) << 4, h1 = (int)((qh >> (2 * j + 1)) & 1u) << 4;
            const int a0 = ((v0 & 0x0F) | h0) - 16, a1 = ((v0 >> 4) | h1) - 16;
            sxy += a0 * y[i].qs[2 * j + 0] + a1 * y[i].qs[2 * j + 1];
        }
        const float t = d * (float)sxy;
        sumf += t * y[i].d;
    }
    *s = sumf;
}

/* Ggml.cs:1304-1348: sumf += (d * sxy) * y.d + m * (y.s0 + y.s1)  (D9 for d and m). */
void orc_vec_dot_q5_1_q8_1(int n, float *s, const void *vx, const void *vy)
{
    const block_q5_1 *x = (const block_q5_1 *)vx;
    const block_q8_1 *y = (const block_q8_1 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        uint32_t qh; memcpy(&qh, x[i].qh, 4);
        const float d = orc_f16_to_f32(x[i].d), m = orc_f16_to_f32(x[i].m);
        int sxy = 0;
        for (int j = 0; j < QK / 2; j++) {
            const uint8_t v0 = x[i].qs[j];
            const int h0 = (int)((qh >> (2 * j + 0)) & 1u) << 4, h1 = (int)((qh >> (2 * j + 1)) & 1u) << 4;
            const int a0 = (v0 & 0x0F) | h0, a1 = (v0 >> 4) | h1;
            sxy += a0 * y[i].qs[2 * j + 0] + a1 * y[i].qs[2 * j + 1];
        }
        const float t = d * (float)sxy;
        const float u = t * y[i].d;
        const float ss = y[i].s0 + y[i].s1;
        const float w = m * ss;
        const float term = u + w;
        sumf += term;
    }
    *s = sumf;
}

/* Ggml.cs:1351-1380: both sides int8 (D4). */
void orc_vec_dot_q8_0_q8_0(int n, float *s, const void *vx, const void *vy)
{
    const block_q8_0 *x = (const block_q8_0 *)vx;
    const block_q8_0 *y = (const block_q8_0 *)vy;
    const int nb = n / QK;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        int sumi = 0;
        for (int j = 0; j < QK; j++) sumi += (int)x[i].qs[j] * (int)y[i].qs[j];
        const float dd = x[i].d * y[i].d;
        sumf += dd * (float)sumi;
    }
    *s = sumf;
}

/* quantize_fns[] looked up BY TYPE (D1): block bytes, block elements, and which Q8 flavour src1 is quantized to. */
static size_t q_type_size(int t)
{
    switch (t) { case T_Q4_0: return 20; case T_Q4_1: return 24; case T_Q4_2: return 10; case T_Q5_0: return 22;
                 case T_Q5_1: return 24; case T_Q8_0: return 36; case T_Q8_1: return 44; default: return 0; }
}
static int q_blck(int t) { return t == T_Q4_2 ? QK4_2 : QK; }
static int q_dot_type(int t)       /* vec_dot_type, Ggml.cs:226, 235, 244, 254, 263, 272 */
{
    switch (t) { case T_Q4_0: case T_Q4_2: case T_Q5_0: case T_Q8_0: return T_Q8_0;
                 case T_Q4_1: case T_Q5_1: return T_Q8_1; default: return -1; }
}
static void q_vec_dot(int t, int n, float *s, const void *x, const void *y)
{
    switch (t) {
    case T_Q4_0: orc_vec_dot_q4_0_q8_0(n, s, x, y); break;
    case T_Q4_1: orc_vec_dot_q4_1_q8_1(n, s, x, y); break;
    case T_Q4_2: orc_vec_dot_q4_2_q8_0(n, s, x, y); break;
    case T_Q5_0: orc_vec_dot_q5_0_q8_0(n, s, x, y); break;
    case T_Q5_1: orc_vec_dot_q5_1_q8_1(n, s, x, y); break;
    default:     orc_vec_dot_q8_0_q8_0(n, s, x, y); break;
    }
}

/* ---- mul_mat drivers ---- */

/* What ggml_compute_forward_mul_mat sees: shapes and byte strides of three
 * tensors (TypeDefinitions.cs:65-99) and the planner's work buffer. */
typedef struct {
    int type;                 /* src0 type */
    int64_t ne0[4]; uint64_t nb0[4]; const void *src0;
    int64_t ne1[4]; uint64_t nb1[4]; const void *src1;   /* F32 */
    int64_t ned[4]; uint64_t nbd[4]; void *dst;          /* F32 */
    void *wdata; size_t wsize;
    int nth;
} orc_mm;

/* Ggml.cs:3340-3384 -- planner's work-buffer bytes for a MUL_MAT node (D1 repaired). */
size_t orc_mul_mat_work_size(int type, int64_t nelements_src1)
{
    switch (type) {
    case T_F32: return 0;
    case T_F16: return (size_t)(2 * nelements_src1);
    default:
        if (q_dot_type(type) == T_Q8_0) return (size_t)(sizeof(block_q8_0) * (size_t)nelements_src1 / QK);
        if (q_dot_type(type) == T_Q8_1) return (size_t)(sizeof(block_q8_1) * (size_t)nelements_src1 / QK);
        return 0;
    }
}

/* INIT phase, main thread only (Ggml.cs:6362-6379 F16; 6641-6655 quantized). */
static void mm_init(const orc_mm *p)
{
    const int64_t ne10 = p->ne1[0], ne11 = p->ne1[1], ne12 = p->ne1[2], ne13 = p->ne1[3];
    if (p->type == T_F16) {
        uint16_t *w = (uint16_t *)p->wdata; size_t id = 0;
        for (int64_t i13 = 0; i13 < ne13; ++i13) for (int64_t i12 = 0; i12 < ne12; ++i12)
        for (int64_t i11 = 0; i11 < ne11; ++i11) for (int64_t i10 = 0; i10 < ne10; ++i10)
            w[id++] = orc_f32_to_f16(*(const float *)((const char *)p->src1 +
                        i13 * p->nb1[3] + i12 * p->nb1[2] + i11 * p->nb1[1] + i10 * p->nb1[0]));
    } else if (q_dot_type(p->type) >= 0) {
        const int dot0 = q_dot_type(p->type) == T_Q8_0;
        const size_t tsz = dot0 ? sizeof(block_q8_0) : sizeof(block_q8_1);
        const size_t row_size = (size_t)ne10 * tsz / QK;
        char *w = (char *)p->wdata;
        for (int64_t i13 = 0; i13 < ne13; ++i13) for (int64_t i12 = 0; i12 < ne12; ++i12)
        for (int64_t i11 = 0; i11 < ne11; ++i11) {
            const float *row = (const float *)((const char *)p->src1 + i13 * p->nb1[3] + i12 * p->nb1[2] + i11 * p->nb1[1]);
            if (dot0) orc_quantize_row_q8_0(row, w, (int)ne10);
            else      orc_quantize_row_q8_1(row, w, (int)ne10);
            w += row_size;
        }
    }
}

/* COMPUTE phase of thread ith (Ggml.cs:6127-6164 F32; 6390-6425 F16; 6662-6699 Q). */
static void mm_compute(const orc_mm *p, int ith)
{
    const int64_t ne00 = p->ne0[0], ne01 = p->ne0[1], ne02 = p->ne0[2], ne03 = p->ne0[3];
    const int64_t ne11 = p->ne1[1], ne12 = p->ne1[2];
    const uint64_t ne0 = (uint64_t)p->ned[0];
    const uint64_t nr = (uint64_t)(ne01 * ne02 * ne03);
    const uint64_t dr = (nr + (uint64_t)p->nth - 1) / (uint64_t)p->nth;
    const uint64_t ir0 = dr * (uint64_t)ith;
    const uint64_t ir1 = ir0 + dr < nr ? ir0 + dr : nr;
    size_t row_size = 0;
    if (q_dot_type(p->type) == T_Q8_0) row_size = (size_t)ne00 * sizeof(block_q8_0) / QK;
    if (q_dot_type(p->type) == T_Q8_1) row_size = (size_t)ne00 * sizeof(block_q8_1) / QK;

    for (uint64_t ir = ir0; ir < ir1; ++ir) {
        const uint64_t i03 = ir / (uint64_t)(ne02 * ne01);
        const uint64_t i02 = (ir - i03 * ne02 * ne01) / (uint64_t)ne01;
        const uint64_t i01 = ir - i03 * ne02 * ne01 - i02 * ne01;
        const char *src0_row = (const char *)p->src0 + i01 * p->nb0[1] + i02 * p->nb0[2] + i03 * p->nb0[3];
        if (p->type == T_F32) {
            for (int64_t ic = 0; ic < ne11; ++ic)
                orc_vec_dot_f32((int)ne00,
                    (float *)((char *)p->dst + i01 * p->nbd[0] + (uint64_t)ic * p->nbd[1] + i02 * p->nbd[2] + i03 * p->nbd[3]),
                    (const float *)src0_row,
                    (const float *)((const char *)p->src1 + (uint64_t)ic * p->nb1[1] + i02 * p->nb1[2] + i03 * p->nb1[3]));
        } else {
            /* dst_col[ic*ne0]: these two drivers index dst densely along dim 1 (Ggml.cs:6423, 6697) */
            float *dst_col = (float *)((char *)p->dst + i01 * p->nbd[0] + i02 * p->nbd[2] + i03 * p->nbd[3]);
            if (p->type == T_F16) {
                const uint16_t *col = (const uint16_t *)p->wdata + (i02 * (uint64_t)ne11 + i03 * (uint64_t)ne12 * (uint64_t)ne11) * (uint64_t)ne00;
                for (uint64_t ic = 0; ic < (uint64_t)ne11; ++ic)
                    orc_vec_dot_f16((int)ne00, &dst_col[ic * ne0], (const uint16_t *)src0_row, col + ic * (uint64_t)ne00);
            } else {
                const char *col = (const char *)p->wdata + (i02 * (uint64_t)ne11 + i03 * (uint64_t)ne12 * (uint64_t)ne11) * row_size;
                for (uint64_t ic = 0; ic < (uint64_t)ne11; ++ic)
                    q_vec_dot(p->type, (int)ne00, &dst_col[ic * ne0], src0_row, col + ic * row_size);
            }
        }
    }
}

typedef struct { const orc_mm *p; int ith; } mm_arg;
static void *mm_thread(void *a) { mm_compute(((mm_arg *)a)->p, ((mm_arg *)a)->ith); return NULL; }

/* ggml_compute_forward_mul_mat (Ggml.cs:6714-6744) driven the way
 * ggml_graph_compute does (Ggml.cs:3553-3628): INIT on the calling thread, then
 * COMPUTE on nth threads with the dr = ceil(nr/nth) row split.  Returns 0, or
 * -1 for a type the dispatch asserts on, -2 if wsize is too small. */
int orc_mul_mat(const orc_mm *p)
{
    if (p->type != T_F32 && p->type != T_F16 && q_dot_type(p->type) < 0) return -1;     /* Q4_3 / Q8_1 weights: null table entries */
    int64_t nel1 = p->ne1[0] * p->ne1[1] * p->ne1[2] * p->ne1[3];
    if (orc_mul_mat_work_size(p->type, nel1) > p->wsize) return -2;
    mm_init(p);
    const int nth = p->nth < 1 ? 1 : p->nth;
    orc_mm q = *p; q.nth = nth;
    if (nth == 1) { mm_compute(&q, 0); return 0; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nth);
    mm_arg *args = (mm_arg *)malloc(sizeof(mm_arg) * (size_t)nth);
    for (int i = 1; i < nth; i++) { args[i].p = &q; args[i].ith = i; pthread_create(&th[i], NULL, mm_thread, &args[i]); }
    mm_compute(&q, 0);
    for (int i = 1; i < nth; i++) pthread_join(th[i], NULL);
    free(th); free(args);
    return 0;
}

/* Convenience for contiguous 2-D operands: W[M][K] (type), X[N][K] f32 -> Y[N][M] f32. */
int orc_mul_mat_2d(int type, const void *W, int64_t M, int64_t K, const float *X, int64_t N, float *Y, int nth)
{
    size_t tsz = type == T_F32 ? 4 : type == T_F16 ? 2 : q_dot_type(type) >= 0 ? q_type_size(type) : 0;
    int64_t blck = (type == T_F32 || type == T_F16) ? 1 : q_blck(type);
    if (!tsz || K % blck || (blck > 1 && K % QK)) return -1;            /* the dot asserts n % QK8_0 == 0 (Ggml.cs:1209) */
    orc_mm p; memset(&p, 0, sizeof p);
    p.type = type; p.nth = nth;
    p.ne0[0] = K; p.ne0[1] = M; p.ne0[2] = p.ne0[3] = 1;
    p.nb0[0] = tsz; p.nb0[1] = tsz * (uint64_t)(K / blck); p.nb0[2] = p.nb0[1] * (uint64_t)M; p.nb0[3] = p.nb0[2];
    p.ne1[0] = K; p.ne1[1] = N; p.ne1[2] = p.ne1[3] = 1;
    p.nb1[0] = 4; p.nb1[1] = 4 * (uint64_t)K; p.nb1[2] = p.nb1[1] * (uint64_t)N; p.nb1[3] = p.nb1[2];
    p.ned[0] = M; p.ned[1] = N; p.ned[2] = p.ned[3] = 1;
    p.nbd[0] = 4; p.nbd[1] = 4 * (uint64_t)M; p.nbd[2] = p.nbd[1] * (uint64_t)N; p.nbd[3] = p.nbd[2];
    p.src0 = W; p.src1 = X; p.dst = Y;
    p.wsize = orc_mul_mat_work_size(type, K * N);
    p.wdata = p.wsize ? malloc(p.wsize) : NULL;
    int rc = orc_mul_mat(&p);
    free(p.wdata);
    return rc;
}

/* Row helpers over many rows (what ggml_compute_forward_dup_f32 does per row, Ggml.cs:4339-4363). */
int orc_quantize_rows(int type, const float *x, void *y, int64_t nrows, int64_t k)
{
    if (k % QK || !q_type_size(type)) return -1;
    size_t rs = (size_t)(k / q_blck(type)) * q_type_size(type);
    for (int64_t r = 0; r < nrows; r++) {
        const float *xr = x + r * k; char *yr = (char *)y + (size_t)r * rs;
        switch (type) {
        case T_Q4_0: orc_quantize_row_q4_0(xr, yr, (int)k); break;
        case T_Q4_1: orc_quantize_row_q4_1(xr, yr, (int)k); break;
        case T_Q4_2: orc_quantize_row_q4_2(xr, yr, (int)k); break;
        case T_Q5_0: orc_quantize_row_q5_0(xr, yr, (int)k); break;
        case T_Q5_1: orc_quantize_row_q5_1(xr, yr, (int)k); break;
        case T_Q8_0: orc_quantize_row_q8_0(xr, yr, (int)k); break;
        default:     orc_quantize_row_q8_1(xr, yr, (int)k); break;
        }
    }
    return 0;
}

int orc_dequantize_rows(int type, const void *x, float *y, int64_t nrows, int64_t k)
{
    if (k % QK || q_dot_type(type) < 0) return -1;                   /* Q8_1 has no dequantize_row_q (Ggml.cs:278) */
    size_t rs = (size_t)(k / q_blck(type)) * q_type_size(type);
    for (int64_t r = 0; r < nrows; r++) {
        const char *xr = (const char *)x + (size_t)r * rs; float *yr = y + r * k;
        switch (type) {
        case T_Q4_0: orc_dequantize_row_q4_0(xr, yr, (int)k); break;
        case T_Q4_1: orc_dequantize_row_q4_1(xr, yr, (int)k); break;
        case T_Q4_2: orc_dequantize_row_q4_2(xr, yr, (int)k); break;
        case T_Q5_0: orc_dequantize_row_q5_0(xr, yr, (int)k); break;
        case T_Q5_1: orc_dequantize_row_q5_1(xr, yr, (int)k); break;
        default:     orc_dequantize_row_q8_0(xr, yr, (int)k); break;
        }
    }
    return 0;
}

void orc_f32_to_f16_row(const float *x, uint16_t *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = orc_f32_to_f16(x[i]); }
void orc_f16_to_f32_row(const uint16_t *x, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = orc_f16_to_f32(x[i]); }

/* ==================================================================================================
 * Neighbours of mul_mat in a Llama layer: F32 element-wise ops, the quantized
 * accumulate add_q_f32, and the strided dup behind cont(transpose(x)).  Same rules as above: one C
 * function per reference function, reference operation order and rounding.
 * ================================================================================================== */

/* ggml_vec_add_f32 / ggml_vec_mul_f32 (Ggml.cs:2586-2589, 2621-2624) over the rows that
 * ggml_compute_forward_add_f32 / mul_f32 walk (Ggml.cs:4622-4685, 5007-5034): one float operation per element. */
void orc_add_f32(int64_t n, const float *x, const float *y, float *z) { for (int64_t i = 0; i < n; i++) z[i] = x[i] + y[i]; }
void orc_mul_f32(int64_t n, const float *x, const float *y, float *z) { for (int64_t i = 0; i < n; i++) z[i] = x[i] * y[i]; }

/* ggml_compute_forward_scale_f32 (Ggml.cs:6746-6780) -> ggml_vec_scale_f32 (Ggml.cs:1416-1443, scalar branch): the
 * result tensor is a view of src0 (ggml_scale_impl, Ggml.cs:8248-8271), so dst is scaled in place. */
void orc_scale_f32(int64_t n, float *y, float v) { for (int64_t i = 0; i < n; i++) y[i] *= v; }

/* ggml_vec_silu_f32 with GGML_SILU_FP16 defined (GGMLSharp.csproj:9; Ggml.cs:2736-2746): y = (float)table_silu_f16[(Half)x],
 * table_silu_f16[i] = (Half)ggml_silu_f32(f16_to_f32(i)) built in ggml_init (Ggml.cs:1461-1471), ggml_silu_f32(x) =
 * x / (1.0f + MathF.Exp(-x)) (Ggml.cs:2723-2726).  MathF.Exp is the platform C runtime's expf. */
static uint16_t g_silu_table[1 << 16];
static int g_silu_ready = 0;
static void silu_table_init(void)
{
    if (g_silu_ready) return;
    for (int i = 0; i < (1 << 16); i++) {
        const float f = orc_f16_to_f32((uint16_t)i);
        const float e = expf(-f);
        const float den = 1.0f + e;
        g_silu_table[i] = orc_f32_to_f16(f / den);
    }
    g_silu_ready = 1;
}
void orc_silu_table(uint16_t *out) { silu_table_init(); memcpy(out, g_silu_table, sizeof g_silu_table); }
void orc_silu_f32(int64_t n, const float *x, float *y)
{
    silu_table_init();
    for (int64_t i = 0; i < n; i++) y[i] = orc_f16_to_f32(g_silu_table[orc_f32_to_f16(x[i])]);
}

/* ggml_compute_forward_rms_norm_f32 (Ggml.cs:5858-5921): per row, sum of float products in a double, mean cast to
 * float, scale = 1.0f / MathF.Sqrt(mean + eps) with eps = 1e-6f, then y = x * scale. */
void orc_rms_norm_f32(int64_t nrows, int64_t ne00, const float *x, int64_t x_stride, float *y, int64_t y_stride)
{
    const float eps = 1e-6f;
    for (int64_t r = 0; r < nrows; r++) {
        const float *xr = x + r * x_stride;
        float *yr = y + r * y_stride;
        double sum = 0.0;
        for (int64_t i = 0; i < ne00; i++) { const float p = xr[i] * xr[i]; sum += p; }
        const float mean = (float)(sum / (double)(uint64_t)ne00);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < ne00; i++) yr[i] = xr[i] * scale;
    }
}

/* ggml_compute_forward_add_q_f32 (Ggml.cs:4797-4906) for contiguous rows: dequantize_row_q -> ggml_vec_acc_f32
 * (y[i] += x[i], Ggml.cs:2591-2594) -> quantize_row_q, with the codec table looked up by type (defect D1) and the scalar
 * quantizers (defect D5). */
int orc_add_q_f32(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k)
{
    if (q_dot_type(type) < 0) return -2;
    if (k % 32) return -1;
    const size_t rb = (size_t)(k / q_blck(type)) * q_type_size(type);
    float *w = (float *)malloc((size_t)k * sizeof(float));
    if (!w) return -4;
    for (int64_t r = 0; r < nrows; r++) {
        const uint8_t *s = (const uint8_t *)src0 + (size_t)r * rb;
        uint8_t *d = (uint8_t *)dst + (size_t)r * rb;
        orc_dequantize_rows(type, s, w, 1, k);
        for (int64_t i = 0; i < k; i++) w[i] += src1[r * k + i];
        orc_quantize_rows(type, w, d, 1, k);
    }
    free(w);
    return 0;
}

/* ggml_compute_forward_dup_f32, non-contiguous source -> contiguous F32 destination (Ggml.cs:4199-4398): elements are
 * visited in (i03, i02, i01, i00) order and written densely -- what ggml_cont(ggml_transpose(x)) executes
 * (MUL_MAT's backward, Ggml.cs:7453-7462). */
void orc_dup_f32_strided(const void *src, const int64_t ne[4], const uint64_t nb[4], float *dst)
{
    int64_t id = 0;
    for (int64_t i3 = 0; i3 < ne[3]; i3++)
        for (int64_t i2 = 0; i2 < ne[2]; i2++)
            for (int64_t i1 = 0; i1 < ne[1]; i1++)
                for (int64_t i0 = 0; i0 < ne[0]; i0++)
                    dst[id++] = *(const float *)((const char *)src + i0 * nb[0] + i1 * nb[1] + i2 * nb[2] + i3 * nb[3]);
}

/* ggml_compute_forward_repeat_f32 (Ggml.cs:5340-5383): 2-D tiling of src0 into dst, row copies in (i, j, k) order. */
void orc_repeat_f32(const float *src, int64_t nc0, int64_t nr0, float *dst, int64_t nc, int64_t nr)
{
    const int64_t ncr = nc / nc0, nrr = nr / nr0;
    for (int64_t i = 0; i < nrr; i++)
        for (int64_t j = 0; j < ncr; j++)
            for (int64_t k = 0; k < nr0; k++)
                memcpy(dst + (i * nr0 + k) * nc + j * nc0, src + k * nc0, (size_t)nc0 * sizeof(float));
}
