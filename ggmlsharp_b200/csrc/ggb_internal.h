// Internal declarations shared by the CUDA translation units of libggb200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ggb200.h"

#define GGB_QK 32
// nodes one activation / GEMV launch can carry (kernel-parameter descriptors: 256 x 40 B, inside the 32 KB parameter space)
#define GGB_MAX_BATCH_NODES 256
// nodes one grouped tensor-core launch can carry (two 128-byte tensor maps + 128 B of descriptor each)
#define GGB_GEMM_GROUP_NODES 64

namespace ggb {

// ---- error plumbing (ggb_shim.cu) ----
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);
#define GGB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return ggb::set_error(GGB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static inline size_t type_size(int t) {
    switch (t) { case GGML_TYPE_F32: return 4; case GGML_TYPE_F16: return 2; case GGML_TYPE_Q4_0: return 20;
                 case GGML_TYPE_Q4_1: return 24; case GGML_TYPE_Q8_0: return 36; case GGML_TYPE_Q8_1: return 44;
                 case GGML_TYPE_Q4_2: return 10; case GGML_TYPE_Q5_0: return 22; case GGML_TYPE_Q5_1: return 24;
                 case GGML_TYPE_I8: return 1; case GGML_TYPE_I16: return 2; case GGML_TYPE_I32: return 4; default: return 0; }
}
static inline int blck_size(int t) {
    switch (t) { case GGML_TYPE_Q4_0: case GGML_TYPE_Q4_1: case GGML_TYPE_Q5_0: case GGML_TYPE_Q5_1: case GGML_TYPE_Q8_0: case GGML_TYPE_Q8_1: return GGB_QK;
                 case GGML_TYPE_Q4_2: return 16; default: return 1; }
}
// weight types whose mul_mat goes through a quantized dot (the non-null vec_dot_q rows of quantize_fns[], Ggml.cs:219-282)
static inline bool is_q_weight(int t) {
    return t == GGML_TYPE_Q4_0 || t == GGML_TYPE_Q4_1 || t == GGML_TYPE_Q4_2 || t == GGML_TYPE_Q5_0 || t == GGML_TYPE_Q5_1 || t == GGML_TYPE_Q8_0;
}
// the sibling formats of SURVEY 8f-2 (everything quantized except the two headline Q4 types)
static inline bool is_sibling_q(int t) { return is_q_weight(t) && t != GGML_TYPE_Q4_0 && t != GGML_TYPE_Q4_1; }
static inline bool is_mm_weight(int t) { return t == GGML_TYPE_F32 || t == GGML_TYPE_F16 || is_q_weight(t); }
// bytes of the weights that pair with ONE 32-element activation block (Q4_2: two 16-element blocks)
static inline size_t q32_bytes(int t) { return t == GGML_TYPE_Q4_2 ? 20 : type_size(t); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- codecs (ggb_codecs.cu) ----
// reference-layout row codecs; rows of k elements, src row stride ldx elements, dst rows packed
int launch_quantize_rows(int type, const float *src, int64_t ldx, void *dst, int64_t nrows, int64_t k, cudaStream_t s);
int launch_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s);
// ggml_compute_forward_add_q_f32 on contiguous rows: dst = quantize(dequantize(src0) + src1); dst may alias src0
int launch_add_q_f32(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, cudaStream_t s);
// ggb_codecs_sib.cu: the same three for Q4_2 / Q5_0 / Q5_1 / Q8_0 (the launchers above forward to these), and the dense fp16
// expansion [M][K] of a sibling-format weight matrix that the batched (tensor-core) path multiplies
int launch_quantize_rows_sib(int type, const float *src, int64_t ldx, void *dst, int64_t nrows, int64_t k, cudaStream_t s);
int launch_dequantize_rows_sib(int type, const void *src, float *dst, int64_t nrows, int64_t k, cudaStream_t s);
int launch_add_q_f32_sib(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, cudaStream_t s);
// ew (may be null): per-row power-of-two exponents from launch_weight_rowexp; row m is expanded as value * 2^-ew[m]
int launch_expand_f16(int type, const void *W, int64_t nb01, __half *out, int64_t M, int64_t K, const int *ew, cudaStream_t s, bool pdl);
// Range handling of the tensor-core path.  The reference keeps block scales and dot products in float32 (Ggml.cs:1158, 1190-1196);
// the MMA operands are fp16, so both operands are pre-scaled by an exact power of two per row and the epilogue undoes it:
//   ew[m] = ilogb(max over the row's blocks of the largest |dequantized value| a block can hold) - 13   (weights, this kernel)
//   ex[n] = ilogb(max |x[n][:]|) - 13                                                                   (activations, k_act_f16_dequant)
// so every operand row peaks in [2^13, 2^14) whatever the magnitude of the data, and dst = acc * 2^(ew[m] + ex[n]).
int launch_weight_rowexp(int type, const void *W, int64_t nb01, int64_t M, int64_t K, int *ew, cudaStream_t s);
// the same for every node of a batch in one launch (row0 / total_rows are filled in by the launcher)
struct RowExpNode { const uint8_t *W; long long nb01; int *ew; long long row0; int M, kb, type, pad_; };
struct RowExpBatch { int n_nodes, pad_; long long total_rows; RowExpNode node[GGB_GEMM_GROUP_NODES]; };
int launch_weight_rowexp_batch(RowExpBatch &b, cudaStream_t s);

// ---- F32 neighbours of mul_mat (ggb_ops.cu) ----
int launch_binary_f32(int op, const float *a, const float *b, float *dst, int64_t n, cudaStream_t s);       // GGML_OP_ADD / GGML_OP_MUL
int launch_scale_f32(float *y, float v, int64_t n, cudaStream_t s);                                          // in place
int launch_silu_f32(const float *x, float *y, int64_t n, cudaStream_t s);
int launch_rms_norm_f32(const float *x, int64_t x_stride, float *y, int64_t y_stride, int64_t nrows, int64_t ne00, cudaStream_t s);   // strides in floats
int launch_repeat_f32(const float *src, int64_t src_stride, int64_t nc0, int64_t nr0, float *dst, int64_t dst_stride, int64_t nc, int64_t nr, cudaStream_t s);   // strides in floats
int launch_dup_f32_strided(const void *src, const int64_t ne[4], const uint64_t nb[4], float *dst, cudaStream_t s);                  // dst contiguous

// Activation staging for mul_mat (the reference's INIT phase, Ggml.cs:6362-6379 / 6641-6655), device-private layouts:
//   Q4_0/Q4_1 weights: "Q8P" rows  [kb x 16 B even quants][kb x 16 B odd quants][kb x {float d; int sum}]
//   F16 weights:       K halfs (RNE), F32 weights: K floats (copied so every row is 16-byte aligned and dense)
size_t act_row_bytes(int wtype, int64_t K);
struct ActNode { const float *x; long long ldx_bytes; uint8_t *out; int N; int blk0; };
// bps > 1: unit-major Q8P planes for the fast GEMV (block b = bps*u + j stored at plane index j*(kb/bps) + u); else linear
// no_wait: the caller guarantees that neither the activations nor the rows being written are touched by work still in flight on the
// stream, so the kernel does not wait for its predecessor before it starts (it still does before it COMPLETES, which keeps completion
// transitive along a chain of programmatically dependent launches)
struct ActHdr { int n_nodes; int K; int kb; int row_bytes; int wtype; int total_blk; int vec16; int bps; int no_wait; int pad_; };
// kernel parameters above 4 KB cost several microseconds per launch, so launches with few nodes pass the small variant
template <int CAP> struct ActBatchT : ActHdr { ActNode node[CAP]; };
using ActBatch = ActBatchT<GGB_MAX_BATCH_NODES>;
constexpr int GGB_SMALL_BATCH_NODES = 32;
int launch_act_batch(const ActBatch &b, cudaStream_t s, bool pdl);
// batched path: activations as dense fp16 [Npad][K] holding d * q (the value the reference's dot multiplies by)
int launch_act_f16_dequant(int wtype, int perm, const float *x, int64_t ldx_bytes, __half *out, int *ex, int64_t N, int64_t Npad, int64_t K, cudaStream_t s, bool wait_prior, bool pdl = true);
// the same for every node of a grouped GEMM launch, one kernel (grid.y = node)
struct ActGemmNode { const float *x; long long ldx_bytes; __half *out; int *ex; int N, Npad, K, vec16; };   // ex: per-row exponents out (null for F16 weights)
struct ActGemmBatch { int n_nodes, wtype, perm, wait_prior; ActGemmNode node[GGB_GEMM_GROUP_NODES]; };
// pdl = false: an ordinary stream-ordered launch (used when an earlier kernel of the batch produced the GEMM's weights)
int launch_act_f16_dequant_batch(ActGemmBatch &b, cudaStream_t s, bool pdl = true);

// ---- GEMV (ggb_gemv.cu) ----
struct GemvNode { const uint8_t *W; const uint8_t *xq; float *y; int M; int ldy; int g0; int ngroups; };
struct GemvHdr {
    int n_nodes, total_groups;
    int type, ncols;
    int K, row_bytes;          // row_bytes = bytes of K elements
    long long nb01;
    int rs, nchunk, chunk_bytes;
    int stage_bytes, depth;
    int xcol_bytes;
    int async;                 // 1: cp.async.bulk staging (16-byte aligned rows), 0: plain-load staging
    int n_peers;
    long long peer_delta[7];   // byte offset from a node's y to the same element of peer p's copy
    int fuse_x;                // 1: node.xq is the F32 activation row itself; every CTA quantizes it in its prologue (gemv_can_fuse_x)
    int wait_w;                // 1: some node's weights are written by work enqueued earlier on this stream (GGB_MM_W_IN_FLIGHT): the weight copies wait too
};
template <int CAP> struct GemvBatchT : GemvHdr { GemvNode node[CAP]; };
using GemvBatch = GemvBatchT<GGB_MAX_BATCH_NODES>;
int launch_gemv_batch(const GemvBatch &b, cudaStream_t s, bool pdl);
int gemv_plan(GemvHdr &b, int type, int64_t K, int64_t nb01, int ncols, const void *Wbase_probe);
int gemv_act_bps(const GemvHdr &b);
bool gemv_can_fuse_x(const GemvHdr &b);   // the planned kernel can quantize a single F32 activation row itself (no staging launch)
int gemv_group_rows(const GemvHdr &b);  // weight rows per work group (tile) of the planned kernel   // the ActBatch::bps the planned kernel expects
int gemv_num_ctas();
int64_t gemv_x_budget();             // bytes of staged activation columns one GEMV pass may keep in shared memory

// ---- decode program (ggb_gemv.cu: k_decode_program) ----
// A dependent chain of single-token nodes (one decode step of a Llama layer stack: 10 dependency levels per layer) as ONE persistent
// launch: one CTA per SM walks a list of STEPS.  A step is a short run of element-wise / row ops on a "running row" every CTA keeps
// in shared memory (evaluated redundantly by all CTAs; each stores its slice of every result tensor), then the single-token MUL_MATs
// that multiply that row (the GEMV tiles of the step dealt to the CTAs as in k_gemv_fast), then a grid-wide barrier.  The weight
// copies do not depend on the activations, so the producer warp of a CTA keeps streaming the NEXT steps' weights into the
// shared-memory ring while its consumer warps sit in the barrier: HBM stays busy across the dependency levels, which as separate
// launches cost 2-12 us each against 1.6-8 us of weight streaming (profiles/r02_chain_launches.csv).
enum { DP_LOAD = 0, DP_ADD, DP_MUL, DP_SILU, DP_RMS_NORM, DP_SCALE };
struct DpNode { const uint8_t *W; float *y; float *y2; long long nb01; int M, tile0, ntiles, pad_; };     // y2: SILU of the result (may be y itself: in place), or null
struct DpOp { const float *a, *b; float *dst; float scalar; int op, n, pad_; };      // a == null: the running row; b: second operand (global) or null
struct DpStep { int op0, nops, node0, nnodes, total_tiles, type, K, row_bytes, rs, nchunk, chunk_bytes, stage_bytes, copy0, ncopies, pad_[2]; };
// a result of an EARLIER step on its way to the host arena (the reference's tensors live in host memory): every CTA stores its slice at
// the start of the step, so the PCIe writes of a level run under the following levels instead of after the whole chain
struct DpCopy { const float4 *src; float4 *dst; int n4, pad_; };
constexpr int DP_MAX_STEPS = 96, DP_MAX_NODES = 168, DP_MAX_OPS = 192, DP_MAX_COPIES = 360;      // 30.6 KB of kernel parameters (the limit is 32 764 B): 24 Llama layers
struct DpProgram {
    unsigned *bar;                       // grid barrier counter, zero at launch
    const unsigned short *silu_table;    // table_silu_f16 (ggb_ops.cu)
    int n_steps, depth, slot_bytes, v_bytes, xs_bytes, pad_;
    long long *trace;                    // debugging (GGB200_PROGRAM_TRACE): globaltimer stamps [4 CTAs][64 steps][8 points], else null
    DpStep step[DP_MAX_STEPS];
    DpNode node[DP_MAX_NODES];
    DpOp op[DP_MAX_OPS];
    DpCopy copy[DP_MAX_COPIES];
};
static_assert(sizeof(DpProgram) <= 32764, "kernel parameter space");
int64_t decode_program_row_max();                                   // longest running row (elements)
// fills step.{type, K, row_bytes, rs, nchunk, chunk_bytes, stage_bytes} for single-token nodes of this shape; false: not a shape the program takes
bool decode_program_plan_step(DpStep &st, int type, int64_t K, int64_t nb01, const void *W);
int decode_program_tile_rows(const DpStep &st);                     // weight rows per tile of a planned step
bool decode_program_finish(DpProgram &p);                           // sizes the shared-memory regions; false: does not fit
bool decode_program_available();                                    // cooperative launch of one full CTA per SM is possible on the current device
int launch_decode_program(const DpProgram &p, cudaStream_t s);
int silu_table_device(const unsigned short **table);               // builds / returns the device copy of table_silu_f16

// ---- GEMM (ggb_gemm.cu): tcgen05 batched path ----
struct GemmArgs {
    int type; int64_t M, K, N; const void *W; int64_t nb01; const __half *Xh; int64_t Npad;
    float *Y; int64_t ldy; int n_peers; float *ypeer[7];
    void *trace;              // optional clock64 timeline buffer (128 x int64 per CTA), debugging only
    const int *ew, *ex;       // power-of-two row exponents of the weights [M] / staged activations [>= N rounded up to 256]; both null = unscaled (true F16 weights)
    int wait_w;               // the weights (or ew) were written earlier in this stream by a kernel of the same batch: the weight side waits too
};
// workspace slice of a batched node: [fp16 activations Npad x K][ex: N rounded up to 256 ints][ew: M ints]
static inline size_t gemm_ws_ex_offset(int64_t K, int64_t N) { const int64_t Npad = (N + 15) / 16 * 16; return align_up((size_t)Npad * (size_t)K * 2, 256); }
static inline size_t gemm_ws_ew_offset(int64_t K, int64_t N) { return gemm_ws_ex_offset(K, N) + align_up((size_t)N, 256) * 4; }
bool gemm_supported(int type, int64_t M, int64_t K, int64_t N, int64_t nb01, const void *W);
size_t gemm_workspace_bytes(int type, int64_t M, int64_t K, int64_t N);
int launch_gemm(const GemmArgs &a, void *ws, cudaStream_t s);
// ggb_gemm_grouped.cu: persistent CTA pairs over the tiles of up to GGB_GEMM_GROUP_NODES nodes of one weight type
bool gemm_grouped_supported(int type);
int launch_gemm_grouped(const GemmArgs *args, int count, cudaStream_t s);
int gemm_act_perm(int type);       // 1: the activation buffer must use the K order 0,4,1,5,2,6,3,7 per group of 8

int device_sm_count();
// cudaFuncSetAttribute (and anything else that is per device) must run once on EVERY device the row split drives:
// one flag per call site and device ordinal
struct PerDeviceOnce {
    bool done[16] = {};
    bool need() { int d = 0; cudaGetDevice(&d); d &= 15; if (done[d]) return false; done[d] = true; return true; }
};

#ifdef __CUDACC__
// ---- exact power-of-two range handling of the tensor-core path (see launch_weight_rowexp) ----
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((e + 127) << 23); }          // 2^e, e in [-126, 127]
__device__ __forceinline__ int range_exp(float amax)                                                 // peak lands in [2^13, 2^14)
{
    if (!(amax > 0.0f) || amax > 3.4028234e38f) return 0;                                            // zero, NaN, infinity: leave the row alone
    const int e = ilogbf(amax) - 13;
    return e < -126 ? -126 : e > 126 ? 126 : e;
}
// acc * 2^e for e in [-252, 252]: two half-exponent factors, so no intermediate over- or underflows unless the result itself does
__device__ __forceinline__ float scale2(float acc, int e) { const int h = e >> 1; return acc * exp2i(h) * exp2i(e - h); }
// acc * fa * fb for two power-of-two factors of possibly opposite sign of exponent: the smaller one first, so that a huge weight
// exponent against a tiny activation exponent (or the reverse) cannot overflow on the way to an in-range result
__device__ __forceinline__ float scale_pair(float acc, float fa, float fb) { return (acc * fminf(fa, fb)) * fmaxf(fa, fb); }
#endif

} // namespace ggb

// tensor-map encoding through the driver entry point (ggb_gemm.cu); CUtensorMap comes from <cuda.h>
#include <cuda.h>
namespace ggb {
int make_map_2d(CUtensorMap *map, CUtensorMapDataType dt, const void *base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                uint32_t box0, uint32_t box1, CUtensorMapSwizzle sw);

} // namespace ggb
