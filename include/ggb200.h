/*
 * ggb200.h -- C ABI of libggb200.so, the B200 (sm_100a) backend for GGMLSharp's
 * matrix-multiply path.
 *
 * The reference (kant2002/GGMLSharp, C#) has no FFI of its own: it is one static class
 * with no DllImport.  This header is therefore the boundary a maintainer would bind
 * with P/Invoke at the three seams upstream ggml used for its cuBLAS hook, which the
 * reference still carries as dead `#if GGML_USE_CUBLAS` text:
 *     seam A  pool allocation in ggml_init / ggml_free        Ggml.cs:1543-1545, 1584-1588
 *     seam B  ggml_graph_compute's per-node loop              Ggml.cs:3539-3704
 *     seam C  the MUL_MAT arm of ggml_compute_forward         Ggml.cs:8649-8653 -> 6714-6744
 *     (one-time init where ggml_init_cublas() sat             Ggml.cs:1499-1504)
 * INTEGRATION.md shows the C# stubs.  Every struct below is laid out exactly as the C#
 * struct it mirrors (TypeDefinitions.cs), so the C# side passes its own pointers.
 *
 * Threading: the library keeps one stream pair per device, so pool-level calls (pool_*, tensor_invalidate, mul_mat_node,
 * graph_compute_mul_mats, quantize/dequantize_rows) serialise on one lock, whatever the pool; different contexts may be driven from
 * different host threads.  ggb_dev_* calls run on the stream they are given without that lock; with stream == NULL they use the
 * library's own stream and take it.
 *
 * Conventions: every function returns 0 (GGB_OK) or a negative ggb_status and never
 * throws or aborts; ggb_last_error() returns a thread-local message.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with GGB_E_NODEVICE.
 */
#ifndef GGB200_H
#define GGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGB_ABI_VERSION 2

/* ---- mirrors of TypeDefinitions.cs ------------------------------------------------ */

#define GGML_MAX_DIMS   4
#define GGML_MAX_NODES  4096
#define GGML_MAX_OPT    4
#define GGML_MEM_ALIGN  16      /* Ggml.cs:16 */
#define GGML_OBJECT_SIZE 32     /* sizeof(ggml_object) */

/* TypeDefinitions.cs:153-169 */
typedef enum ggml_type {
    GGML_TYPE_F32 = 0, GGML_TYPE_F16 = 1, GGML_TYPE_Q4_0 = 2, GGML_TYPE_Q4_1 = 3,
    GGML_TYPE_Q4_2 = 4, GGML_TYPE_Q4_3 = 5, GGML_TYPE_Q5_0 = 6, GGML_TYPE_Q5_1 = 7,
    GGML_TYPE_Q8_0 = 8, GGML_TYPE_Q8_1 = 9, GGML_TYPE_I8 = 10, GGML_TYPE_I16 = 11,
    GGML_TYPE_I32 = 12, GGML_TYPE_COUNT = 13
} ggml_type;

/* TypeDefinitions.cs:172-220 (only the values this path dispatches on) */
enum { GGML_OP_NONE = 0, GGML_OP_DUP = 1, GGML_OP_ADD = 2, GGML_OP_MUL = 4, GGML_OP_REPEAT = 10, GGML_OP_SILU = 17, GGML_OP_RMS_NORM = 19,
       GGML_OP_MUL_MAT = 20, GGML_OP_SCALE = 21, GGML_OP_CPY = 22, GGML_OP_CONT = 23, GGML_OP_RESHAPE = 24, GGML_OP_VIEW = 25,
       GGML_OP_PERMUTE = 26, GGML_OP_TRANSPOSE = 27 };

/* TypeDefinitions.cs:65-99 -- 176 bytes, data at 160 */
typedef struct ggml_tensor {
    int32_t  type;
    int32_t  n_dims;
    int64_t  ne[GGML_MAX_DIMS];
    uint64_t nb[GGML_MAX_DIMS];
    int32_t  op;
    uint8_t  is_param;          /* C# bool, 1 byte */
    uint8_t  _pad0[3];
    struct ggml_tensor *grad;
    struct ggml_tensor *src0;
    struct ggml_tensor *src1;
    int64_t  opt[GGML_MAX_OPT];
    int32_t  n_tasks;
    int32_t  perf_runs;
    int64_t  perf_cycles;
    int64_t  perf_time_us;
    void    *data;
    uint8_t  padding[8];
} ggml_tensor;

/* TypeDefinitions.cs:102-152 -- 98 360 bytes, passed by pointer */
typedef struct ggml_cgraph {
    int32_t  n_nodes;
    int32_t  n_leafs;
    int32_t  n_threads;
    int32_t  _pad0;
    uint64_t work_size;
    ggml_tensor *work;
    ggml_tensor *nodes[GGML_MAX_NODES];
    ggml_tensor *grads[GGML_MAX_NODES];
    ggml_tensor *leafs[GGML_MAX_NODES];
    int32_t  perf_runs;
    int32_t  _pad1;
    int64_t  perf_cycles;
    int64_t  perf_time_us;
} ggml_cgraph;

/* TypeDefinitions.cs:236-248 -- weight blocks: float32 scale(s) + 16 bytes; byte j holds
 * element 2j in its low nibble and element 2j+1 in its high nibble (Ggml.cs:361-373). */
typedef struct { float d; uint8_t qs[16]; } block_q4_0;                    /* 20 B */
typedef struct { float d; float m; uint8_t qs[16]; } block_q4_1;           /* 24 B */
/* TypeDefinitions.cs:277-290 -- activation blocks (quants are int8: SURVEY.md defect D4) */
typedef struct { float d; int8_t qs[32]; } block_q8_0;                     /* 36 B */
typedef struct { float d; float s0; float s1; int8_t qs[32]; } block_q8_1; /* 44 B */
/* TypeDefinitions.cs:249-275 -- the sibling weight blocks.  d / m are IEEE binary16 BIT PATTERNS (the
 * reference declares them ushort and stores a numeric cast, defect D9 in oracle/ggb_oracle.c; block_q5_0.d is a Half there);
 * qh bit l is the fifth bit of element l; Q8_0 as a weight type is block_q8_0 above.  2-byte aligned, no padding. */
typedef struct { uint16_t d; uint8_t qs[8]; } block_q4_2;                               /* 10 B per 16 elements */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } block_q5_0;               /* 22 B */
typedef struct { uint16_t d; uint16_t m; uint8_t qh[4]; uint8_t qs[16]; } block_q5_1;   /* 24 B */
typedef char ggb_assert_block_q4_2[sizeof(block_q4_2) == 10 ? 1 : -1];
typedef char ggb_assert_block_q5_0[sizeof(block_q5_0) == 22 ? 1 : -1];
typedef char ggb_assert_block_q5_1[sizeof(block_q5_1) == 24 ? 1 : -1];

/* ---- status ------------------------------------------------------------------------ */

typedef enum ggb_status {
    GGB_OK            =  0,
    GGB_E_INVALID     = -1,   /* an assert of the reference path would have fired (shape/stride/type) */
    GGB_E_UNSUPPORTED = -2,   /* valid ggml, but not a type/op this backend implements */
    GGB_E_CUDA        = -3,   /* a CUDA call failed; message in ggb_last_error() */
    GGB_E_NOMEM       = -4,
    GGB_E_ABI         = -5,   /* struct layout mismatch between host language and this library */
    GGB_E_NODEVICE    = -6    /* no usable sm_100 device: there is deliberately no CPU fallback */
} ggb_status;

const char *ggb_last_error(void);
int  ggb_abi_version(void);

/* Called once from ggml_init's first-call block (where ggml_init_cublas() was, Ggml.cs:1499-1504)
 * with the host language's own sizeof/offsetof so a layout drift is refused, not corrupted. */
int  ggb_abi_check(int sizeof_tensor, int offsetof_data, int sizeof_cgraph, int offsetof_nodes,
                   int sizeof_block_q4_0, int sizeof_block_q4_1);

/* Idempotent.  Selects the device (env GGB200_DEVICE, default 0), creates the stream. */
int  ggb_init(void);
int  ggb_shutdown(void);
int  ggb_device_count(int *count);

/* ---- seam A: the memory pool behind a ggml_context --------------------------------- */

typedef struct ggb_pool ggb_pool;

/* Replaces NativeMemory.AlignedAlloc(mem_size, 16) (Ggml.cs:1545).  *host_base is pinned host
 * memory, >= 16-byte aligned; tensor->data pointers keep pointing into it, so user code that
 * reads and writes tensor->data directly (Test3/Program.cs:37-40) is unchanged. */
int  ggb_pool_alloc(size_t bytes, void **host_base, ggb_pool **pool);
/* For a caller-supplied ggml_init_params.mem_buffer (Ggml.cs:1543-1544): the memory stays the
 * caller's; the pool only tracks device mirrors of tensors inside it. */
int  ggb_pool_adopt(void *host_base, size_t bytes, ggb_pool **pool);
/* Replaces NativeMemory.AlignedFree (Ggml.cs:1587); frees device mirrors, and the host memory
 * only if ggb_pool_alloc made it (mirrors mem_buffer_owned, Ggml.cs:1546). */
int  ggb_pool_free(ggb_pool *pool);
/* Weight residency is OPT-IN.  The reference re-reads src0->data on every ggml_graph_compute (Ggml.cs:6139-6164, 6676-6699) and user
 * code writes tensor->data through raw pointers, so by default every leaf src0 is uploaded again on every compute.  With the cache
 * on (this call, or env GGB200_WEIGHT_CACHE=1 for pools created afterwards) a leaf src0 is uploaded on first use and its device
 * mirror kept; the caller then promises that such tensors change only through the API (CPY / in-place graph nodes and ggml_set_*
 * invalidate by byte range) or are followed by ggb_tensor_invalidate.  Never cached even then: parameters (is_param) and tensors
 * with a gradient, which ggml_opt rewrites in place between computes (Ggml.cs:1734-1760).  on = 0 also drops every mirror. */
int  ggb_pool_set_weight_cache(ggb_pool *pool, int on);
/* Row split across the GPUs of one box, inside ONE process and behind the unchanged ggml_graph_compute (north_star; the reference
 * splits the rows of src0 over the OS threads of ggml_graph_compute, Ggml.cs:3231-3252, 6665-6672 -- here each "thread" drives a GPU).
 * max_devices: 0 = off (default; env GGB200_ROW_SPLIT=<n>|all for pools created afterwards), < 0 = every sm_100 device with mutual
 * peer access (GGB200_DEVICES=<list> restricts the set), n = at most n.  A MUL_MAT is split only if its src0 is a 2-D tensor of at
 * least min_weight_bytes (0 = keep; default 4 MiB, env GGB200_ROW_SPLIT_MIN_BYTES): device g multiplies rows
 * [ceil(M/G) g, ceil(M/G) (g+1)) and its kernel stores the results into every device's copy of dst over NVLink (the all-gather,
 * fused into the epilogue); element-wise neighbours are replicated; smaller mul_mats run on device 0, which broadcasts them.  A
 * compute with prompt-sized nodes (N >= 16) uses at most 2 devices, because there the exchange of fp32 results outgrows the
 * math (see ggb_shim.cu: row_split_width).  Results are identical to the single-GPU ones bit for bit: a row's dot products do not
 * depend on which device computes them. */
int  ggb_pool_set_row_split(ggb_pool *pool, int max_devices, size_t min_weight_bytes);
/* The partition rule itself (no device needed): rows [*row0, *row0 + *rows) of an M-row src0 that device g of G multiplies --
 * dr = ceil(M / G), [dr g, min(dr g + dr, M)), the reference's thread split (Ggml.cs:6665-6672) with nth = G. */
int  ggb_row_split_rows(int64_t M, int g, int G, int64_t *row0, int64_t *rows);
/* Call after rewriting tensor->data of a cached weight behind the API's back: drops every mirror that shares a byte with the
 * tensor (a view of a cached leaf, or a leaf some cached view looks into); NULL drops every mirror of the pool. */
int  ggb_tensor_invalidate(ggb_pool *pool, const ggml_tensor *t);

/* ---- seam C / seam B: execute MUL_MAT nodes ------------------------------------------ */

/* Drop-in for ggml_compute_forward_mul_mat (Ggml.cs:6714-6744) on one node, synchronous:
 * validates what the reference asserts (Ggml.cs:6016-6034, 6221-6238, 6481-6504, 6694), uploads
 * src1 (and src0 unless cached), runs the kernels, copies dst->data back, fills perf_*. */
int  ggb_mul_mat_node(ggb_pool *pool, ggml_tensor *dst);

#define GGB_GRAPH_KEEP_ON_DEVICE 1   /* do not copy intermediate results back (only graph outputs) */
#define GGB_GRAPH_NO_WEIGHT_CACHE 2  /* re-upload every src0 (what the CPU path observes if weights change) */
#define GGB_GRAPH_MUL_MAT_ONLY 4     /* run MUL_MAT and F32 -> {F16, quantized} CPY nodes only */
#define GGB_GRAPH_SHARD 8            /* row-split this compute over the GPUs of the box even if the pool was not configured for it */

/* Called from ggml_graph_compute (Ggml.cs:3539) instead of walking MUL_MAT nodes one by one:
 * runs, in node order and on one stream with one final sync, every MUL_MAT node (and F32 ->
 * {F16, Q4_0, Q4_1, Q4_2, Q5_0, Q5_1, Q8_0} CPY node, the public route to quantize_row_q,
 * Ggml.cs:4339-4363) whose inputs
 * are leafs or nodes it runs itself.  done[i] (n_nodes bytes, may be NULL) is set to 1 for each
 * node it executed; the caller's loop runs the rest.  Returns the number executed or < 0. */
int  ggb_graph_compute_mul_mats(ggb_pool *pool, ggml_cgraph *graph, int flags, uint8_t *done);
/* Repeated computes: the second time the SAME cgraph arrives (same node headers, shapes, data pointers, flags; single-device computes
 * over pinned memory) the call records everything it enqueues -- uploads from the host arena, staging, kernels, result copies -- as a
 * CUDA graph, and replays it from then on: one launch instead of re-planning the node list (Ggml.cs:3260-3704 does that per compute).
 * The replay reads and writes the tensors' own bytes, so it follows whatever the user wrote into tensor->data in between exactly
 * as the first compute did.  kernel_launches and the byte counters of ggb_stats keep counting per compute; graph_replays counts
 * the replays.  Env GGB200_NO_GRAPH_CACHE=1 turns it off. */
/* The selection alone, without a device: done[i] = 1 for every node the call above would execute (or look through, for view ops).
 * The reference runs nodes strictly in order (Ggml.cs:3539-3704); seam B runs the selected nodes BEFORE the caller's loop runs the
 * rest, so a candidate is refused -- with everything downstream of it -- when that reordering could be observed through memory:
 * an operand that shares bytes with the output of an earlier node left to the CPU (ggml_*_inplace writes src0's bytes, CPY src1's),
 * an in-place / CPY destination that an earlier CPU node still reads or writes, or an operand that overlaps an earlier selected
 * node's result without lying inside it.  Returns the number of nodes it would execute or < 0. */
int  ggb_graph_plan(ggml_cgraph *graph, int flags, uint8_t *done);
/* Unless GGB_GRAPH_MUL_MAT_ONLY is set the same call also keeps the neighbours of mul_mat in a Llama layer on the device
 * (SURVEY.md 8f), so consecutive MUL_MATs need no host round trip: F32 ADD / MUL (Ggml.cs:4622-4685, 5007-5034), SILU
 * (5705-5747, fp16 table semantics of GGML_SILU_FP16), RMS_NORM (5858-5921), SCALE (6746-6780, in place on the view of src0),
 * ADD with a quantized src0 (add_q_f32, 4797-4906; Q4_0, Q4_1, Q4_2, Q5_0, Q5_1, Q8_0), 2-D REPEAT (5340-5383) and CONT / DUP of a transposed or permuted F32 tensor (4199-4398; what
 * MUL_MAT's backward issues, 7453-7462).  RESHAPE / VIEW / PERMUTE / TRANSPOSE nodes are no-ops in the reference
 * (8668-8687) and are marked done.  An element-wise node is taken only when at least one operand is produced on the device
 * by this call (otherwise uploading it would cost more than the C# loop); contiguous tensors only. */

/* ---- the codec column of quantize_fns[] (Ggml.cs:219-290) ----------------------------- */

/* quantize_row_q / dequantize_row_q over nrows rows of k elements; src and dst may each be a
 * host or a device pointer.  Bit-exact with quantize_row_q4_0_reference_impl (Ggml.cs:334-377),
 * quantize_row_q4_1_reference_impl (487-528), quantize_row_q8_0/q8_1 (733-823, defects D2-D4
 * repaired) and the scalar dequantize_row_q4_0/q4_1 (884-911, 961-987).  F16 = (Half) cast.
 * The sibling formats of the same table: quantize_row_q4_2/q5_0/q5_1_reference_impl
 * (547-590, 609-653, 672-714) and dequantize_row_q4_2/q5_0/q5_1/q8_0 (992-1122), with fp16 block
 * scales as IEEE bit patterns (the reference's numeric `(ushort)(Half)d` cast is a defect, see
 * oracle/ggb_oracle.c D9); k must be a multiple of 32 for every quantized type.  Q8_1 has no
 * dequantizer and Q4_3 no entry at all in the reference's table: GGB_E_UNSUPPORTED. */
int  ggb_quantize_rows(int type, const float *src, void *dst, int64_t nrows, int64_t k);
int  ggb_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k);

/* ---- device-resident entry points (what the executor itself calls) --------------------- */

/* One mul_mat on device pointers: dst[n][m] = sum_k W[m][k] * X[n][k]  (m < M, n < N).
 * W: M rows of K elements of `type`, row stride nb01 bytes.  X, Y: float32, row strides in
 * bytes.  A rank of a row-split passes its slice of W with Y offset to its column block.
 * type: F32, F16, Q4_0, Q4_1 (the north-star path) and Q4_2, Q5_0, Q5_1, Q8_0 (every other type
 * whose quantize_fns[] row has a vec_dot_q, Ggml.cs:219-282); Q4_3 / Q8_1 weights: GGB_E_UNSUPPORTED. */
typedef struct ggb_dev_mm {
    int32_t  type;
    int32_t  n_peers;              /* 0, or number of extra destinations in Y_peer (fused row-split epilogue) */
    int64_t  M, K, N;
    const void  *W;  int64_t nb01;
    const float *X;  int64_t ldx_bytes;
    float       *Y;  int64_t ldy_bytes;
    float       *Y_peer[7];        /* peer-mapped copies of Y on other GPUs, written by the same kernel */
    const int32_t *W_rowexp;       /* optional, N >= 16 with quantized W: M ints from ggb_dev_weight_rowexp (computed once while W is
                                      resident); NULL = computed inside every call (one extra pass over the block headers) */
    int32_t      flags;            /* GGB_MM_* */
    int32_t      _pad;
} ggb_dev_mm;
#define GGB_MM_W_IN_FLIGHT 1       /* W (or W_rowexp) is written by work enqueued earlier on the same stream: the weight loads, which otherwise
                                      start before the activation staging has finished, wait for it as well */
#define GGB_MM_X_HOST      2       /* X is pinned host memory the kernels read in place (UVA): it is staged by ONE pass over it; without the flag a
                                      small batch of single-token nodes lets every CTA of the GEMV read the row itself, which is right for device memory
                                      (16 KB from L2) and 148 trips over PCIe for host memory */

/* Range handling of the tensor-core path (N >= 16, quantized W).  The reference multiplies float32 block scales (Ggml.cs:1158,
 * 1190-1196); the MMA operands are fp16, so each weight row m is pre-scaled by the exact power of two 2^-rowexp[m] and each staged
 * activation row by its own, and the epilogue multiplies both back.  rowexp[m] = ilogb(largest |value| row m can dequantize to) - 13.
 * This computes it for M rows of K elements of `type` (row stride nb01 bytes) on `stream`; pass the result as ggb_dev_mm.W_rowexp. */
int  ggb_dev_weight_rowexp(int type, const void *W, int64_t nb01, int64_t M, int64_t K, int32_t *rowexp, void *stream);

/* Scratch the batch needs (quantized / converted activations, the reference's `wdata`). */
size_t ggb_dev_workspace_bytes(const ggb_dev_mm *mm, int count);
/* Enqueue `count` independent mul_mats on `stream` (a cudaStream_t; NULL = the library's own).
 * Single-token nodes that share (type, K) are fused into one persistent GEMV launch. */
int  ggb_dev_mul_mat_batch(const ggb_dev_mm *mm, int count, void *workspace, size_t workspace_bytes, void *stream);
/* The same call in two halves, for callers that pipeline a stream of single-token batches themselves (bench.py does): phase 1 only
 * stages the activations of the batch into `workspace` (the reference's INIT phase: src1 -> Q8 / Half rows, Ggml.cs:6362-6379, 6641-6655),
 * phase 2 only multiplies, reading what a phase-1 call with the same nodes and workspace left there; phase 0 = ggb_dev_mul_mat_batch.
 * With two workspaces the staging of batch i+1 (on one stream) runs under the GEMV of batch i (on another).  Phase 3 stages like
 * phase 1 but does not wait for the preceding kernel of its stream before it starts: for a caller who knows that neither the activations
 * nor this workspace are touched by work in flight (a third workspace) -- the staging kernel can then be issued IN the GEMV stream,
 * between two GEMVs, without breaking their back-to-back launch.  N < 16 nodes only. */
int  ggb_dev_mul_mat_batch_phase(const ggb_dev_mm *mm, int count, void *workspace, size_t workspace_bytes, void *stream, int phase);
/* Device-pointer codecs on a stream (no sync). */
int  ggb_dev_quantize_rows(int type, const float *src, void *dst, int64_t nrows, int64_t k, void *stream);
int  ggb_dev_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k, void *stream);

/* The same neighbours on device pointers (no sync).  op is GGML_OP_ADD or GGML_OP_MUL; n counts floats; rms_norm / repeat row
 * strides are in floats (repeat: dst[r][c] = src[r % nr0][c % nc0]); ggb_dev_cont copies the strided F32 view (ne, nb in bytes: the fields of a transposed / permuted ggml_tensor) into a
 * contiguous dst; ggb_dev_add_q is add_q_f32 over nrows contiguous rows of k elements (any quantized weight type), dst may alias src0. */
int  ggb_dev_binary(int op, const float *a, const float *b, float *dst, int64_t n, void *stream);
int  ggb_dev_scale(float *x, float v, int64_t n, void *stream);
int  ggb_dev_silu(const float *x, float *dst, int64_t n, void *stream);
int  ggb_dev_rms_norm(const float *x, int64_t x_stride, float *dst, int64_t dst_stride, int64_t nrows, int64_t ne00, void *stream);
int  ggb_dev_repeat(const float *src, int64_t src_stride, int64_t nc0, int64_t nr0, float *dst, int64_t dst_stride, int64_t nc, int64_t nr, void *stream);
int  ggb_dev_cont(const void *src, const int64_t ne[4], const uint64_t nb[4], float *dst, void *stream);
int  ggb_dev_add_q(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, void *stream);

/* Plain device-memory helpers so a host language without a CUDA binding can stage data. */
int  ggb_dev_alloc(size_t bytes, void **dptr);
int  ggb_dev_free(void *dptr);
int  ggb_dev_upload(void *dptr, const void *host, size_t bytes);
int  ggb_dev_download(void *host, const void *dptr, size_t bytes);
int  ggb_stream_sync(void *stream);

/* Same-node multi-process row-split: export a device allocation / map a peer's (CUDA IPC). */
int  ggb_ipc_export(void *dptr, uint8_t handle[64]);
int  ggb_ipc_open(const uint8_t handle[64], void **peer_dptr);
int  ggb_ipc_close(void *peer_dptr);
/* The exchange step of the fused row split.  Every rank owns `flags` (world x uint64, zero-initialised, inside an
 * IPC-exported allocation); peer_flags[r] is rank r's array mapped here (peer_flags[rank] = own).  Enqueues one tiny
 * kernel that publishes `epoch` into slot `rank` of every rank's array with system-scope release stores over NVLink
 * and then waits until all `world` local slots hold >= epoch: when it completes, every peer's Y_peer stores that were
 * enqueued before its own barrier call are visible in this rank's buffer.  epoch must increase by 1 per call. */
int  ggb_peer_barrier(uint64_t *const *peer_flags, int rank, int world, uint64_t epoch, void *stream);
/* The row-split exchange as ONE kernel over peer memory (no NCCL): copies this rank's n_seg segments
 * [seg_offset + s*seg_stride, +seg_bytes) of its symmetric buffer into the same place of every peer's buffer with
 * coalesced 16-byte stores over NVLink, then -- last CTA -- runs the flag barrier above.  peer_bases[r] is rank r's buffer
 * mapped here (own for r == rank), `counter` 16 zero-initialised bytes of local device memory (CTA counter; a uint64
 * running epoch at +8 that is used when `epoch` is passed as 0, which makes the call CUDA-graph replayable).  seg_offset, seg_bytes and
 * seg_stride must be multiples of 16.  Launched with programmatic dependent launch: it waits for the preceding kernel
 * (the GEMV that produced the segments) on the device, not on the host. */
int  ggb_peer_push_barrier(void *const *peer_bases, uint64_t *const *peer_flags, uint32_t *counter, int rank, int world,
                           size_t seg_offset, size_t seg_bytes, size_t seg_stride, int n_seg, uint64_t epoch, void *stream);

/* ---- counters -------------------------------------------------------------------------- */

typedef struct ggb_stats {
    uint64_t kernel_launches;      /* kernels of this library launched since init / last reset */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t weight_uploads, weight_cache_hits;
    uint64_t nodes_executed;
    double   last_graph_device_ms; /* CUDA-event time of the last ggb_graph_compute_mul_mats / ggb_mul_mat_node */
    double   timed_kernel_ms;      /* with ggb_set_kernel_timing(1): summed CUDA-event time of the GEMV / GEMM launches only */
    uint64_t timed_kernel_launches;
    uint64_t graph_replays;        /* computes served by replaying the CUDA graph recorded for the same cgraph (ggb_graph_compute_mul_mats) */
} ggb_stats;
/* Measurement aid for the roofline.  1: bracket every mul_mat kernel launch (not the activation staging) with CUDA events on
 * the launching stream -- the brackets defeat the programmatic-dependent-launch overlap, so leave it off when timing steps.
 * 2: additionally skip the activation staging of single-token nodes and reuse what the previous identical call left in the
 * workspace, so that a stream of calls is a stream of the GEMV kernel alone (time it with your own events). 0: off. */
int  ggb_set_kernel_timing(int on);
/* The decode program: a ggml_graph_compute whose device-runnable nodes are a dependent chain of single-token mul_mats (F32 / F16 /
 * quantized weights) and their ADD / MUL / SILU / RMS_NORM / SCALE neighbours on single rows -- one decode step of a layer stack -- is
 * enqueued as ONE persistent kernel that walks the dependency levels behind grid-wide barriers while its weight copies run ahead of
 * them (ggb_gemv.cu: k_decode_program).  Same results to the bit as the per-level launches it replaces; everything else (wide
 * levels, prompt-sized nodes, other ops, the row split) takes the per-level route as before.  1 (default): on, 0: per-level launches
 * only.  Process-wide; recorded graphs are dropped when it changes.  Environment: GGB200_NO_DECODE_PROGRAM=1 starts with it off. */
int  ggb_set_decode_program(int on);
int  ggb_get_stats(ggb_stats *out);
int  ggb_reset_stats(void);

#ifdef __cplusplus
}
#endif
#endif /* GGB200_H */
