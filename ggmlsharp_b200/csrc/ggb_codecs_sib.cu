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

// One thread per 16 output bytes: lane sub = t & 7 of a group expands elements 4*sub .. 4*sub+3, so a warp store is 512 contiguous bytes.
template <int TYPE>
__global__ void __launch_bounds__(256) k_dequantize_sib(const uint8_t *__restrict__ x, float *__restrict__ y, long long ngrp)
{
    constexpr int G = Sib<TYPE>::G;
    const long long total = ngrp * 8;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int sub = (int)(t & 7);
        const unsigned short *g = reinterpret_cast<const unsigned short *>(x + (t >> 3) * G);
        float4 o;
        if (TYPE == GGML_TYPE_Q4_2) {
            const unsigned short *b = g + 5 * (sub >> 2);
            const float d = h_val(__ldg(b));
            const uint32_t q = __ldg(b + 1 + (sub & 3));
            o.x = __fmul_rn((float)((int)(q & 15u) - 8), d); o.y = __fmul_rn((float)((int)((q >> 4) & 15u) - 8), d);
            o.z = __fmul_rn((float)((int)((q >> 8) & 15u) - 8), d); o.w = __fmul_rn((float)((int)(q >> 12) - 8), d);
        } else if (TYPE == GGML_TYPE_Q5_0 || TYPE == GGML_TYPE_Q5_1) {
            constexpr int Q0 = TYPE == GGML_TYPE_Q5_0 ? 3 : 4;
            const float d = h_val(__ldg(g));
            const float m = TYPE == GGML_TYPE_Q5_1 ? h_val(__ldg(g + 1)) : 0.0f;
            const uint32_t qh = ((uint32_t)__ldg(g + Q0 - 2) | ((uint32_t)__ldg(g + Q0 - 1) << 16)) >> (4 * sub);
            const uint32_t q = __ldg(g + Q0 + sub);
            const int n0 = (int)((q & 15u) | ((qh & 1u) << 4)), n1 = (int)(((q >> 4) & 15u) | (((qh >> 1) & 1u) << 4));
            const int n2 = (int)(((q >> 8) & 15u) | (((qh >> 2) & 1u) << 4)), n3 = (int)((q >> 12) | (((qh >> 3) & 1u) << 4));
            if (TYPE == GGML_TYPE_Q5_0) {
                o.x = __fmul_rn((float)(n0 - 16), d); o.y = __fmul_rn((float)(n1 - 16), d);
                o.z = __fmul_rn((float)(n2 - 16), d); o.w = __fmul_rn((float)(n3 - 16), d);
            } else {                                             // product rounded, then sum rounded
                o.x = __fadd_rn(__fmul_rn((float)n0, d), m); o.y = __fadd_rn(__fmul_rn((float)n1, d), m);
                o.z = __fadd_rn(__fmul_rn((float)n2, d), m); o.w = __fadd_rn(__fmul_rn((float)n3, d), m);
            }
        } else {
            const float d = __uint_as_float((uint32_t)__ldg(g) | ((uint32_t)__ldg(g + 1) << 16));
            const uint32_t q0 = __ldg(g + 2 + 2 * sub), q1 = __ldg(g + 3 + 2 * sub);
            o.x = __fmul_rn((float)(int)(int8_t)(q0 & 0xFFu), d); o.y = __fmul_rn((float)(int)(int8_t)(q0 >> 8), d);
            o.z = __fmul_rn((float)(int)(int8_t)(q1 & 0xFFu), d); o.w = __fmul_rn((float)(int)(int8_t)(q1 >> 8), d);
        }
        reinterpret_cast<float4 *>(y)[t] = o;
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
template <int TYPE>
__global__ void __launch_bounds__(256) k_expand_f16(const uint8_t *__restrict__ W, long long nb01, __half *__restrict__ out, long long M, int K)
{
    constexpr int G = Sib<TYPE>::G;
    const int oct_per_row = K >> 3;
    const long long total = M * oct_per_row;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / oct_per_row;
        const int oc = (int)(t - row * oct_per_row), sub = oc & 3;              // elements 8*sub .. 8*sub+7 of group oc >> 2
        const unsigned short *g = reinterpret_cast<const unsigned short *>(W + row * nb01 + (long long)(oc >> 2) * G);
        float v[8];
        if (TYPE == GGML_TYPE_Q4_2) {
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

#define GGB_SIB_SWITCH(type, EXPR) \
    switch (type) { \
    case GGML_TYPE_Q4_2: { constexpr int T = GGML_TYPE_Q4_2; EXPR; } break; \
    case GGML_TYPE_Q5_0: { constexpr int T = GGML_TYPE_Q5_0; EXPR; } break; \
    case GGML_TYPE_Q5_1: { constexpr int T = GGML_TYPE_Q5_1; EXPR; } break; \
    case GGML_TYPE_Q8_0: { constexpr int T = GGML_TYPE_Q8_0; EXPR; } break; \
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
    const unsigned grid = (unsigned)std::min<long long>((ngrp + 127) / 128, (long long)device_sm_count() * 16);
    GGB_SIB_SWITCH(type, (k_quantize_sib<T><<<grid, 128, 0, s>>>(src, ldx, (uint8_t *)dst, ngrp, kb)));
    count_launch(); GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int launch_dequantize_rows_sib(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s)
{
    if (nrows <= 0 || k <= 0) return GGB_OK;
    if (k % GGB_QK) return set_error(GGB_E_INVALID, "dequantize: k=%lld is not a multiple of %d", (long long)k, GGB_QK);
    if ((reinterpret_cast<uintptr_t>(src) & 1) || (reinterpret_cast<uintptr_t>(dst) & 15)) return set_error(GGB_E_UNSUPPORTED, "dequantize: unaligned operand");
    const long long ngrp = nrows * (k / GGB_QK);
    const unsigned grid = (unsigned)std::min<long long>((ngrp * 8 + 255) / 256, (long long)device_sm_count() * 16);
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

int launch_expand_f16(int type, const void *W, int64_t nb01, __half *out, int64_t M, int64_t K, cudaStream_t s, bool pdl)
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
    GGB_SIB_SWITCH(type, GGB_CUDA(cudaLaunchKernelEx(&cfg, k_expand_f16<T>, (const uint8_t *)W, (long long)nb01, out, (long long)M, (int)K)));
    count_launch();
    return GGB_OK;
}

} // namespace ggb
