// Host build of ggmlsharp_b200/csrc/ggb_sib_math.cuh: the sibling-format register arithmetic (quantize / dequantize one group,
// the GEMV's per-unit quantized dots) compiled as plain C++ so the CPU test-suite can check it against the oracle bit for bit
// BEFORE a GPU is involved.  Only the CUDA intrinsics the header uses are supplied here; everything else is the product source.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC (tests/test_sib_emul.py does it).  TEST INFRASTRUCTURE.
#define GGB_HOST_EMUL 1
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/ggb200.h"

#define GGB_DI static inline
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __float2int_rz(float t) { return (int)t; }                     // callers exclude NaN / out-of-range first
static inline long long __float2ll_rz(float t) { return (long long)t; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __dp4a(int a, int b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)(int8_t)((uint32_t)a >> (8 * i)) * (int)(int8_t)((uint32_t)b >> (8 * i));
    return c;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
namespace ggb { namespace sib {
static inline uint32_t h_bits(float v) { _Float16 h = (_Float16)v; uint16_t u; memcpy(&u, &h, 2); return u; }   // gcc: RNE, like (Half)float
static inline float h_val(uint32_t bits) { uint16_t u = (uint16_t)bits; _Float16 h; memcpy(&h, &u, 2); return (float)h; }
}}
#include "../../ggmlsharp_b200/csrc/ggb_sib_math.cuh"

using namespace ggb::sib;

template <int T> static void quant_rows(const float *x, uint8_t *y, long n32)
{
    constexpr int G = Grp<T>::G;
    for (long g = 0; g < n32; g++) {
        float e[32]; memcpy(e, x + 32 * g, sizeof e);
        uint32_t o[G / 2];
        quantize_group<T>(e, o);
        for (int i = 0; i < G / 2; i++) { const uint16_t h = (uint16_t)o[i]; memcpy(y + g * G + 2 * i, &h, 2); }
    }
}
template <int T> static void dequant_rows(const uint8_t *x, float *y, long n32)
{
    constexpr int G = Grp<T>::G;
    for (long g = 0; g < n32; g++) {
        uint32_t w[G / 2];
        for (int i = 0; i < G / 2; i++) { uint16_t h; memcpy(&h, x + g * G + 2 * i, 2); w[i] = h; }
        float e[32];
        dequantize_group<T>(w, e);
        memcpy(y + 32 * g, e, sizeof e);
    }
}

// One activation row as k_act_batch stages it ("Q8P"): quantize_row_q8_0's d and quants, even / odd split, block sum (Q4_2: two
// packed half sums).  The quants come from the caller (the oracle's q8_0 row), so only the LAYOUT is restated here.
static XBlk make_xblk(const uint8_t *q8blk, bool q4_2)
{
    float d; memcpy(&d, q8blk, 4);
    const int8_t *q = (const int8_t *)(q8blk + 4);
    XBlk x;
    int ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0}, lo = 0, hi = 0;
    for (int i = 0; i < 32; i++) {
        const uint32_t b = (uint8_t)q[i];
        if (i & 1) od[i / 8] |= (int)(b << (8 * ((i % 8) / 2))); else ev[i / 8] |= (int)(b << (8 * ((i % 8) / 2)));
        if (i < 16) lo += q[i]; else hi += q[i];
    }
    x.ev = int4{ev[0], ev[1], ev[2], ev[3]}; x.od = int4{od[0], od[1], od[2], od[3]};
    memcpy(&x.ds.x, &d, 4);
    x.ds.y = q4_2 ? (int)(((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16)) : lo + hi;
    return x;
}

// The fast GEMV's view of one weight row: whole units, lane u%32 owns unit u, lanes summed at the end.
template <int T> static float row_dot_units(const uint8_t *row, const uint8_t *xq8, long K)
{
    constexpr int UB = Unit<T>::BYTES, NG = Unit<T>::GROUPS;
    const long units = K / 32 / NG;
    float lane_acc[32] = {0};
    for (long u = 0; u < units; u++) {
        uint32_t v[UB / 4];
        memcpy(v, row + u * UB, UB);
        XBlk x[4];
        for (int j = 0; j < NG; j++) x[j] = make_xblk(xq8 + (u * NG + j) * 36, T == GGML_TYPE_Q4_2);
        lane_acc[u % 32] = dot_unit_words<T>(v, x, lane_acc[u % 32]);
    }
    float s = 0; for (int l = 0; l < 32; l++) s += lane_acc[l];
    return s;
}
// The plain-staging GEMV's view: one group at a time, gathered as 16-bit words (ggb_gemv.cu: dot_sib).
template <int T> static float row_dot_groups(const uint8_t *row, const uint8_t *xq8, long K)
{
    constexpr int G = Grp<T>::G, NW = (G + 3) / 4;
    float lane_acc[32] = {0};
    for (long g = 0; g < K / 32; g++) {
        uint32_t wd[NW];
        for (int k = 0; k < NW; k++) {
            uint16_t a = 0, b = 0;
            memcpy(&a, row + g * G + 4 * k, 2);
            if (2 * k + 1 < G / 2) memcpy(&b, row + g * G + 4 * k + 2, 2);
            wd[k] = (uint32_t)a | ((uint32_t)b << 16);
        }
        lane_acc[g % 32] = dot_group_words<T>(wd, make_xblk(xq8 + g * 36, T == GGML_TYPE_Q4_2), lane_acc[g % 32]);
    }
    float s = 0; for (int l = 0; l < 32; l++) s += lane_acc[l];
    return s;
}

#define SW(type, EXPR) switch (type) { \
    case GGML_TYPE_Q4_2: { constexpr int T = GGML_TYPE_Q4_2; EXPR; } break; case GGML_TYPE_Q5_0: { constexpr int T = GGML_TYPE_Q5_0; EXPR; } break; \
    case GGML_TYPE_Q5_1: { constexpr int T = GGML_TYPE_Q5_1; EXPR; } break; case GGML_TYPE_Q8_0: { constexpr int T = GGML_TYPE_Q8_0; EXPR; } break; default: return -1; }

extern "C" {
int emul_quantize(int type, const float *x, uint8_t *y, long n32) { SW(type, quant_rows<T>(x, y, n32)); return 0; }
int emul_dequantize(int type, const uint8_t *x, float *y, long n32) { SW(type, dequant_rows<T>(x, y, n32)); return 0; }
// W: M rows of K elements of `type`; xq8: ONE activation row already quantized by quantize_row_q8_0 (36-byte blocks)
int emul_gemv(int type, int units, const uint8_t *W, long M, long K, const uint8_t *xq8, float *y)
{
    long rb = 0;
    SW(type, rb = K / 32 * Grp<T>::G);
    for (long m = 0; m < M; m++) { SW(type, y[m] = units ? row_dot_units<T>(W + m * rb, xq8, K) : row_dot_groups<T>(W + m * rb, xq8, K)); }
    return 0;
}
}
