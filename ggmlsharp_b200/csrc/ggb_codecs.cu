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

#include <algorithm>

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
template <int TYPE>
__device__ __forceinline__ void quantize_block_q4(const float (&e)[32], uint32_t *out)
{
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
        float mn = e[0], mx = e[0];                            // Ggml.cs:496-504 (equal values keep the earlier element)
#pragma unroll
        for (int i = 1; i < 32; i++) { if (e[i] < mn) mn = e[i]; if (e[i] > mx) mx = e[i]; }
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

template <int TYPE>
__global__ void __launch_bounds__(256) k_quantize_q4_rows(const float *__restrict__ x, long long ldx, uint8_t *__restrict__ y,
                                                          long long nblk, int kb)
{
    // (A variant that staged loads and stores through shared memory for fully coalesced 128-bit global accesses measured
    //  slower -- 2.9 vs 3.4 TB/s on 4096x4096 -- the kernel is latency-, not transaction-bound; profiles/README.md.)
    for (long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += (long long)gridDim.x * blockDim.x) {
        const long long row = blk / kb;
        const int col = (int)(blk - row * kb);
        const float4 *p = reinterpret_cast<const float4 *>(x + row * ldx + (long long)col * GGB_QK);
        float e[32];
#pragma unroll
        for (int i = 0; i < 8; i++) { const float4 v = __ldg(p + i); e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w; }
        quantize_block_q4<TYPE>(e, reinterpret_cast<uint32_t *>(y + blk * (TYPE == GGML_TYPE_Q4_0 ? 20 : 24)));
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

// Ggml.cs:886-910 (Q4_0) and 962-987 (Q4_1): one thread per nibble byte -> two floats
template <int TYPE>
__global__ void __launch_bounds__(256) k_dequantize_rows(const uint8_t *__restrict__ x, float *__restrict__ y, long long nblk)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long blk = t >> 4;
    const int j = (int)(t & 15);
    if (blk >= nblk) return;
    constexpr int BS = TYPE == GGML_TYPE_Q4_0 ? 20 : 24;
    const uint8_t *b = x + blk * BS;
    const float d = *reinterpret_cast<const float *>(b);
    float2 o;
    if (TYPE == GGML_TYPE_Q4_0) {
        const uint8_t vi = b[4 + j];
        o.x = __fmul_rn((float)((int)(vi & 0x0F) - 8), d);
        o.y = __fmul_rn((float)((int)(vi >> 4) - 8), d);
    } else {
        const float m = *reinterpret_cast<const float *>(b + 4);
        const uint8_t vi = b[8 + j];
        o.x = __fadd_rn(__fmul_rn((float)(vi & 0x0F), d), m);   // product rounded, then sum rounded
        o.y = __fadd_rn(__fmul_rn((float)(vi >> 4), d), m);
    }
    reinterpret_cast<float2 *>(y)[t] = o;
}

// ---- activation staging for mul_mat (the reference's INIT phase) ----

// Q8P: quantize exactly as quantize_row_q8_0 / q8_1 do (same d, same 32 quants), stored even/odd-split so a
// 32-bit word of weight nibbles pairs with one 32-bit word of activations for dp4a, plus the block's integer
// sum (Q4_0 needs -8*sum, Q4_1 needs m*d*sum; the reference's s0+s1 is d*sum).
__global__ void __launch_bounds__(256) k_act_batch(const __grid_constant__ ActBatch b)
{
    if (b.wtype == GGML_TYPE_Q4_0 || b.wtype == GGML_TYPE_Q4_1) {
        const int lane = threadIdx.x & 31, sub = lane & 7;
        const int blk = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
        const bool live = blk < b.total_blk;
        int n = 0;
        if (live) while (n + 1 < b.n_nodes && blk >= b.node[n + 1].blk0) n++;
        const ActNode &nd = b.node[n];
        const int local = live ? blk - nd.blk0 : 0;
        const int row = local / b.kb, col = local - row * b.kb;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("griddepcontrol.wait;" ::: "memory");      // activations may be the previous kernel's output
        if (live) {
            const char *p = reinterpret_cast<const char *>(nd.x) + (long long)row * nd.ldx_bytes + (long long)col * 128 + sub * 16;
            if (b.vec16) v = *reinterpret_cast<const float4 *>(p);
            else { const float *f = reinterpret_cast<const float *>(p); v = make_float4(f[0], f[1], f[2], f[3]); }
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
        float amax = fmaxf(fmaxf(fabsf(e[0]), fabsf(e[1])), fmaxf(fabsf(e[2]), fabsf(e[3])));
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
        int q[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { q[i] = (int)(int8_t)rne_byte(__fmul_rn(e[i], id)); s += q[i]; }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        uint32_t ev = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[2] & 0xFF) << 8);
        uint32_t od = (uint32_t)(q[1] & 0xFF) | ((uint32_t)(q[3] & 0xFF) << 8);
        ev |= __shfl_down_sync(0xffffffffu, ev, 1) << 16;
        od |= __shfl_down_sync(0xffffffffu, od, 1) << 16;
        if (live) {
            uint8_t *o = nd.out + (long long)row * b.row_bytes;
            const int idx = b.bps > 1 ? (col % b.bps) * (b.kb / b.bps) + col / b.bps : col;
            if ((sub & 1) == 0) {
                reinterpret_cast<uint32_t *>(o + (long long)idx * 16)[sub >> 1] = ev;
                reinterpret_cast<uint32_t *>(o + (long long)b.kb * 16 + (long long)idx * 16)[sub >> 1] = od;
            }
            if (sub == 1) *reinterpret_cast<int2 *>(o + (long long)b.kb * 32 + (long long)idx * 8) = make_int2(__float_as_int(d), s);
        }
    } else {
        // F16 weights: src1 -> Half (Ggml.cs:6362-6379); F32 weights: dense copy.  One thread per 4 elements.
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int q4 = (b.K + 3) / 4;                           // float4 groups per row
        const long long total = (long long)b.total_blk * q4;    // total_blk = total rows here
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (t < total) {
            const int grow = (int)(t / q4), c4 = (int)(t - (long long)grow * q4);
            int n = 0;
            while (n + 1 < b.n_nodes && grow >= b.node[n + 1].blk0) n++;
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
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Batched (tensor-core) path: activations as fp16 values the reference's dot effectively multiplies by:
// Q4_x weights -> d * round(x / d) of the Q8 block (same d, same quants as quantize_row_q8_x); F16 -> (Half)x.
// Rows n >= N of the padded buffer are zero.  One thread per block of 32 elements (128 B in, 64 B out): the kernel is
// a pure stream (12 B per element), so everything is kept in registers and the K permutation 0,4,1,5,2,6,3,7 that the
// GEMM's nibble unpack produces costs nothing.
__global__ void __launch_bounds__(256) k_act_f16_dequant(int wtype, int perm, const float *__restrict__ x, long long ldx_bytes,
                                                         __half *__restrict__ out, int N, int Npad, int K, int vec16, int wait_prior)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the GEMM's prologue and weight streaming may start now
    if (wait_prior) asm volatile("griddepcontrol.wait;" ::: "memory"); // x may be the previous kernel's output (first node of a batch only)
    const int kb = K / GGB_QK;
    const long long nblk = (long long)Npad * kb;
    for (long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(blk / kb), col = (int)(blk - (long long)row * kb);
        float e[32];
        if (row < N) {
            const char *p = reinterpret_cast<const char *>(x) + (long long)row * ldx_bytes + (long long)col * 128;
            if (vec16) {
#pragma unroll
                for (int i = 0; i < 8; i++) { const float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i); e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w; }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) e[i] = reinterpret_cast<const float *>(p)[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) e[i] = 0.0f;
        }
        if (wtype != GGML_TYPE_F16) {
            float amax = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(e[i]));
            const float d = __fdiv_rn(amax, 127.0f);
            const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
            if (id < 3.0e38f) {                                    // every sane block: |x*id| <= 127.0000x, plain ties-to-even
#pragma unroll
                for (int i = 0; i < 32; i++) e[i] = __fmul_rn(d, (float)__float2int_rn(__fmul_rn(e[i], id)));
            } else {                                               // 1/d overflowed (subnormal scale): .NET cast semantics
#pragma unroll
                for (int i = 0; i < 32; i++) e[i] = __fmul_rn(d, (float)(int)(int8_t)rne_byte(__fmul_rn(e[i], id)));
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
        uint4 *dst = reinterpret_cast<uint4 *>(out + (long long)row * K + (long long)col * GGB_QK);
#pragma unroll
        for (int i = 0; i < 4; i++) dst[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
}

} // namespace

size_t act_row_bytes(int wtype, int64_t K)
{
    switch (wtype) {
    case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_1: return align_up((size_t)(K / GGB_QK) * 40, 16);
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
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "quantize: k=%lld is not a multiple of %d (Ggml.cs:336)", (long long)k, GGB_QK);
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (ldx & 3))
        return set_error(GGB_E_UNSUPPORTED, "quantize: source rows must be 16-byte aligned");
    const int kb = (int)(k / GGB_QK);
    const long long nblk = nrows * kb;
    const unsigned grid = (unsigned)((nblk * 8 + 255) / 256);
    uint8_t *y = (uint8_t *)dst;
    const unsigned grid4 = (unsigned)std::min<long long>((nblk + 255) / 256, (long long)device_sm_count() * 16);
    switch (type) {
    case GGML_TYPE_Q4_0: k_quantize_q4_rows<GGML_TYPE_Q4_0><<<grid4, 256, 0, s>>>(src, ldx, y, nblk, kb); break;
    case GGML_TYPE_Q4_1: k_quantize_q4_rows<GGML_TYPE_Q4_1><<<grid4, 256, 0, s>>>(src, ldx, y, nblk, kb); break;
    case GGML_TYPE_Q8_0: k_quantize_rows<GGML_TYPE_Q8_0><<<grid, 256, 0, s>>>(src, ldx, y, nblk, kb); break;
    case GGML_TYPE_Q8_1: k_quantize_rows<GGML_TYPE_Q8_1><<<grid, 256, 0, s>>>(src, ldx, y, nblk, kb); break;
    default: return set_error(GGB_E_UNSUPPORTED, "quantize: type %d has no codec on this path", type);
    }
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "dequantize: k=%lld is not a multiple of %d (Ggml.cs:839)", (long long)k, GGB_QK);
    const long long nblk = nrows * (k / GGB_QK);
    const unsigned grid = (unsigned)((nblk * 16 + 255) / 256);
    if (type == GGML_TYPE_Q4_0) k_dequantize_rows<GGML_TYPE_Q4_0><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, nblk);
    else if (type == GGML_TYPE_Q4_1) k_dequantize_rows<GGML_TYPE_Q4_1><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, nblk);
    else return set_error(GGB_E_UNSUPPORTED, "dequantize: type %d has no codec on this path", type);
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_act_batch(const ActBatch &b, cudaStream_t s, bool pdl)
{
    long long threads;
    if (b.wtype == GGML_TYPE_Q4_0 || b.wtype == GGML_TYPE_Q4_1) threads = (long long)b.total_blk * 8;
    else threads = (long long)b.total_blk * ((b.K + 3) / 4);
    if (threads <= 0) return GGB_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((threads + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_batch, b));
    count_launch();
    return GGB_OK;
}

int launch_act_f16_dequant(int wtype, int perm, const float *x, int64_t ldx_bytes, __half *out, int64_t N, int64_t Npad, int64_t K, cudaStream_t s, bool wait_prior)
{
    const long long nblk = Npad * (K / GGB_QK);
    if (nblk <= 0) return GGB_OK;
    const int vec16 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx_bytes & 15) == 0) ? 1 : 0;
    long long grid = (nblk + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_act_f16_dequant, wtype, perm, x, (long long)ldx_bytes, out, (int)N, (int)Npad, (int)K, vec16, wait_prior ? 1 : 0));
    count_launch();
    return GGB_OK;
}

} // namespace ggb
