// Single-token (N < 16) mul_mat: a persistent, bandwidth-bound GEMV over a LIST of graph nodes.
//
// Replaces the COMPUTE phase of ggml_compute_forward_mul_mat_{f32,f16_f32,q_f32} (Ggml.cs:6127-6164,
// 6390-6425, 6662-6699) and the dots ggml_vec_dot_{f32,f16,q4_0_q8_0,q4_1_q8_1} (Ggml.cs:2631-2651,
// 1124-1201).  The reference splits src0 rows over OS threads (dr = ceil(nr/nth)); here rows are split
// over every warp of a one-CTA-per-SM grid, and one launch walks all nodes of a batch.
//
// Data movement: weight rows are contiguous byte ranges (nb01 bytes each), so every warp streams its
// own rows HBM -> shared memory with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a private
// ring of `depth` stages guarded by mbarriers -- no CTA-wide barrier in the steady state, ~10-20 KB in
// flight per warp.  Lanes then read 20-byte Q4_0 / 24-byte Q4_1 blocks from shared memory with a lane
// stride of 5 / 6 words (bank-conflict free), so the raw reference block layout is kept in HBM and the
// unaligned-vector-load problem of 20-byte blocks never reaches the memory system.
// The activation vector(s) were quantized by k_act_batch exactly as quantize_row_q8_0/q8_1 do and sit
// in shared memory for the whole node.
//
// Arithmetic per Q4_0 block is the reference's: an exact int32 block sum (dp4a), then
// sumf += (d0*d1) * (float)sumi in float.  Only the cross-block summation order differs (lane-strided
// partial sums + a shuffle tree instead of one sequential chain).
//
// Roofline: HBM.  Algorithmic bytes per node = M*row_bytes (weights) + 4*K*N (x) + 4*M*N (y).
#include "ggb_internal.h"

namespace ggb {

namespace {

constexpr int NWARPS = 8;
constexpr int MAX_DEPTH = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ---- per-type partial dot of one staged chunk against NC staged activation columns ----

template <int NC>
__device__ __forceinline__ void dot_q4_0(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes, int kb,
                                         int lane, float (&acc)[NC])
{
    const int nblk = nbytes / 20, b0 = off_bytes / 20;
    for (int i = lane; i < nblk; i += 32) {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(w + i * 20);
        const float d0 = __uint_as_float(wp[0]);
        int lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const uint32_t q = wp[1 + j]; lo[j] = (int)(q & 0x0F0F0F0Fu); hi[j] = (int)((q >> 4) & 0x0F0F0F0Fu); }
        const int bi = b0 + i;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint8_t *xc = xs + c * xcol_bytes;
            const int4 ev = *reinterpret_cast<const int4 *>(xc + bi * 16);
            const int4 od = *reinterpret_cast<const int4 *>(xc + kb * 16 + bi * 16);
            const int2 ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + bi * 8);
            int s = ds.y * -8;                                   // sum (q-8)*p = sum q*p - 8*sum p
            s = __dp4a(lo[0], ev.x, s); s = __dp4a(hi[0], od.x, s);
            s = __dp4a(lo[1], ev.y, s); s = __dp4a(hi[1], od.y, s);
            s = __dp4a(lo[2], ev.z, s); s = __dp4a(hi[2], od.z, s);
            s = __dp4a(lo[3], ev.w, s); s = __dp4a(hi[3], od.w, s);
            acc[c] = __fadd_rn(acc[c], __fmul_rn(__fmul_rn(d0, __int_as_float(ds.x)), (float)s));   // Ggml.cs:1158
        }
    }
}

template <int NC>
__device__ __forceinline__ void dot_q4_1(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes, int kb,
                                         int lane, float (&acc)[NC])
{
    const int nblk = nbytes / 24, b0 = off_bytes / 24;
    for (int i = lane; i < nblk; i += 32) {
        const uint2 *wp = reinterpret_cast<const uint2 *>(w + i * 24);
        const uint2 dm = wp[0], qa = wp[1], qb = wp[2];
        const float d0 = __uint_as_float(dm.x), m0 = __uint_as_float(dm.y);
        const uint32_t q[4] = {qa.x, qa.y, qb.x, qb.y};
        int lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { lo[j] = (int)(q[j] & 0x0F0F0F0Fu); hi[j] = (int)((q[j] >> 4) & 0x0F0F0F0Fu); }
        const int bi = b0 + i;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint8_t *xc = xs + c * xcol_bytes;
            const int4 ev = *reinterpret_cast<const int4 *>(xc + bi * 16);
            const int4 od = *reinterpret_cast<const int4 *>(xc + kb * 16 + bi * 16);
            const int2 ds = *reinterpret_cast<const int2 *>(xc + kb * 32 + bi * 8);
            int s = 0;
            s = __dp4a(lo[0], ev.x, s); s = __dp4a(hi[0], od.x, s);
            s = __dp4a(lo[1], ev.y, s); s = __dp4a(hi[1], od.y, s);
            s = __dp4a(lo[2], ev.z, s); s = __dp4a(hi[2], od.z, s);
            s = __dp4a(lo[3], ev.w, s); s = __dp4a(hi[3], od.w, s);
            // sum_j (d0*q_j + m0) * (d1*p_j)  (Ggml.cs:1190-1196)  =  d0*d1*sum q_j p_j  +  m0*d1*sum p_j
            const float d1 = __int_as_float(ds.x);
            acc[c] += (d0 * d1) * (float)s + (m0 * d1) * (float)ds.y;
        }
    }
}

template <int NC>
__device__ __forceinline__ void dot_f16(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes,
                                        int lane, float (&acc)[NC])
{
    const int nvec = nbytes >> 4, v0 = off_bytes >> 4;
    for (int i = lane; i < nvec; i += 32) {
        const uint4 wv = *reinterpret_cast<const uint4 *>(w + i * 16);
        const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
        float2 wf[4];
#pragma unroll
        for (int j = 0; j < 4; j++) wf[j] = __half22float2(wh[j]);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint4 xv = *reinterpret_cast<const uint4 *>(xs + c * xcol_bytes + (v0 + i) * 16);
            const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 xf = __half22float2(xh[j]);
                acc[c] = fmaf(wf[j].x, xf.x, acc[c]);            // the f16*f16 product is exact in f32 (Ggml.cs:2647)
                acc[c] = fmaf(wf[j].y, xf.y, acc[c]);
            }
        }
    }
    const int tail = (nbytes & 15) >> 1;                         // K % 8 elements
    if (lane < tail) {
        const int e = (nvec << 3) + lane;
        const float wv = __half2float(reinterpret_cast<const __half *>(w)[e]);
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] = fmaf(wv, __half2float(reinterpret_cast<const __half *>(xs + c * xcol_bytes + off_bytes)[e]), acc[c]);
    }
}

template <int NC>
__device__ __forceinline__ void dot_f32(const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs, int xcol_bytes,
                                        int lane, float (&acc)[NC])
{
    const int nvec = nbytes >> 4, v0 = off_bytes >> 4;
    for (int i = lane; i < nvec; i += 32) {
        const float4 wv = *reinterpret_cast<const float4 *>(w + i * 16);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + c * xcol_bytes + (v0 + i) * 16);
            // float product, then accumulate (Ggml.cs:2636); the accumulator is f32 here, f64 in the reference
            acc[c] += __fmul_rn(wv.x, xv.x); acc[c] += __fmul_rn(wv.y, xv.y);
            acc[c] += __fmul_rn(wv.z, xv.z); acc[c] += __fmul_rn(wv.w, xv.w);
        }
    }
    const int tail = (nbytes & 15) >> 2;
    if (lane < tail) {
        const int e = (nvec << 2) + lane;
        const float wv = reinterpret_cast<const float *>(w)[e];
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] += __fmul_rn(wv, reinterpret_cast<const float *>(xs + c * xcol_bytes + off_bytes)[e]);
    }
}

template <int TYPE, int NC>
__device__ __forceinline__ void dot_chunk(const GemvBatch &b, const uint8_t *w, int off_bytes, int nbytes, const uint8_t *xs,
                                          int lane, float (&acc)[NC])
{
    if (TYPE == GGML_TYPE_Q4_0) dot_q4_0<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, b.K / GGB_QK, lane, acc);
    else if (TYPE == GGML_TYPE_Q4_1) dot_q4_1<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, b.K / GGB_QK, lane, acc);
    else if (TYPE == GGML_TYPE_F16) dot_f16<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, lane, acc);
    else dot_f32<NC>(w, off_bytes, nbytes, xs, b.xcol_bytes, lane, acc);
}

// The sequence of (group, chunk) items one warp streams: groups g_begin+warp, +NWARPS, ... below g_end.
struct Cursor {
    int g, c, n;
    __device__ __forceinline__ bool valid(int g_end) const { return g < g_end; }
    __device__ __forceinline__ void advance(const GemvBatch &b) { if (++c == b.nchunk) { c = 0; g += NWARPS; } }
    __device__ __forceinline__ void locate(const GemvBatch &b) { while (g >= b.node[n].g0 + b.node[n].ngroups) n++; }
};

template <int TYPE, int NC>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_gemv(const __grid_constant__ GemvBatch b)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xbytes = (NC * b.xcol_bytes + 127) & ~127;
    uint8_t *xs = smem;
    uint8_t *stages = smem + xbytes + (size_t)warp * b.depth * b.stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + xbytes + (size_t)NWARPS * b.depth * b.stage_bytes) + warp * MAX_DEPTH;

    const int g_begin = (int)((long long)b.total_groups * blockIdx.x / gridDim.x);
    const int g_end = (int)((long long)b.total_groups * (blockIdx.x + 1) / gridDim.x);

    if (b.async && lane == 0) {
        for (int s = 0; s < b.depth; s++) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    Cursor prod{g_begin + warp, 0, 0};
    int n_issued = 0;
    auto issue = [&]() {
        // all lanes keep the cursor; lane 0 arms the barrier and launches the copy
        prod.locate(b);
        const GemvNode &nd = b.node[prod.n];
        const int row0 = (prod.g - nd.g0) * b.rs;
        const int rows = min(b.rs, nd.M - row0);
        const uint32_t bytes = b.nchunk == 1 ? (uint32_t)(rows * b.row_bytes)
                                             : (uint32_t)min(b.chunk_bytes, b.row_bytes - prod.c * b.chunk_bytes);
        if (lane == 0) {
            const int st = n_issued % b.depth;
            const uint32_t bar = smem_u32(&bars[st]);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(stages + (size_t)st * b.stage_bytes), nd.W + (long long)row0 * b.nb01 + (long long)prod.c * b.chunk_bytes, bytes, bar);
        }
        n_issued++;
        prod.advance(b);
    };
    // Weights do not depend on the preceding kernel: start streaming before the PDL wait.
    if (b.async)
        for (int i = 0; i < b.depth && prod.valid(g_end); i++) issue();

    asm volatile("griddepcontrol.wait;" ::: "memory");           // activations come from k_act_batch
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (g_begin >= g_end) return;
    int it = 0;
    int n = 0;
    while (g_begin >= b.node[n].g0 + b.node[n].ngroups) n++;
    for (; n < b.n_nodes && b.node[n].g0 < g_end; n++) {
        const GemvNode &nd = b.node[n];
        __syncthreads();                                         // everyone is done with the previous node's x
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(nd.xq);
            uint4 *dst = reinterpret_cast<uint4 *>(xs);
            for (int i = threadIdx.x; i < (NC * b.xcol_bytes) >> 4; i += NWARPS * 32) dst[i] = src[i];
        }
        __syncthreads();
        const int seg_lo = max(g_begin, nd.g0), seg_hi = min(g_end, nd.g0 + nd.ngroups);
        int g = seg_lo + ((warp - (seg_lo - g_begin) % NWARPS + NWARPS) % NWARPS);
        for (; g < seg_hi; g += NWARPS) {
            const int row0 = (g - nd.g0) * b.rs;
            const int rows = min(b.rs, nd.M - row0);
            float acc[NC];
            for (int c = 0; c < b.nchunk; c++, it++) {
                const int st = b.async ? it % b.depth : 0;
                const uint8_t *stage = stages + (size_t)st * b.stage_bytes;
                const int nbytes = b.nchunk == 1 ? rows * b.row_bytes : min(b.chunk_bytes, b.row_bytes - c * b.chunk_bytes);
                if (b.async) {
                    mbar_wait(smem_u32(&bars[st]), (uint32_t)((it / b.depth) & 1));
                } else {
                    // unaligned rows: the warp stages its chunk itself with plain loads
                    const uint8_t *src = nd.W + (long long)row0 * b.nb01 + (long long)c * b.chunk_bytes;
                    __syncwarp();
                    if (b.nchunk == 1) {
                        for (int r = 0; r < rows; r++) {
                            const uint8_t *sr = src + (long long)r * b.nb01;
                            uint8_t *dr = const_cast<uint8_t *>(stage) + r * b.row_bytes;
                            if (((reinterpret_cast<uintptr_t>(sr) | (uintptr_t)b.row_bytes) & 3) == 0)
                                for (int i = lane; i < b.row_bytes >> 2; i += 32) reinterpret_cast<uint32_t *>(dr)[i] = reinterpret_cast<const uint32_t *>(sr)[i];
                            else
                                for (int i = lane; i < b.row_bytes >> 1; i += 32) reinterpret_cast<uint16_t *>(dr)[i] = reinterpret_cast<const uint16_t *>(sr)[i];
                        }
                    } else {
                        uint8_t *dr = const_cast<uint8_t *>(stage);
                        if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 3) == 0)
                            for (int i = lane; i < nbytes >> 2; i += 32) reinterpret_cast<uint32_t *>(dr)[i] = reinterpret_cast<const uint32_t *>(src)[i];
                        else
                            for (int i = lane; i < nbytes >> 1; i += 32) reinterpret_cast<uint16_t *>(dr)[i] = reinterpret_cast<const uint16_t *>(src)[i];
                    }
                    __syncwarp();
                }
                if (b.nchunk == 1) {
                    for (int r = 0; r < rows; r++) {
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
                        dot_chunk<TYPE, NC>(b, stage + r * b.row_bytes, 0, b.row_bytes, xs, lane, acc);
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) {
                            const float v = warp_sum(acc[cc]);
                            if (lane == 0) {
                                float *yp = nd.y + (long long)cc * nd.ldy + row0 + r;
                                *yp = v;
                                for (int p = 0; p < b.n_peers; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + b.peer_delta[p]) = v;
                            }
                        }
                    }
                } else {
                    if (c == 0) {
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) acc[cc] = 0.0f;
                    }
                    dot_chunk<TYPE, NC>(b, stage, c * b.chunk_bytes, nbytes, xs, lane, acc);
                    if (c == b.nchunk - 1) {
#pragma unroll
                        for (int cc = 0; cc < NC; cc++) {
                            const float v = warp_sum(acc[cc]);
                            if (lane == 0) {
                                float *yp = nd.y + (long long)cc * nd.ldy + row0;
                                *yp = v;
                                for (int p = 0; p < b.n_peers; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + b.peer_delta[p]) = v;
                            }
                        }
                    }
                }
                if (b.async) {
                    __syncwarp();                                // every lane has finished reading this stage
                    if (prod.valid(g_end)) issue();
                }
            }
        }
    }
}

template <int TYPE>
int launch_typed(const GemvBatch &b, size_t smem, int grid, cudaStream_t s, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
#define GGB_GEMV_CASE(NCV) case NCV: { \
        static bool attr_set = false; \
        if (!attr_set) { GGB_CUDA(cudaFuncSetAttribute(k_gemv<TYPE, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; } \
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_gemv<TYPE, NCV>, b)); } break;
    switch (b.ncols) {
        GGB_GEMV_CASE(1) GGB_GEMV_CASE(2) GGB_GEMV_CASE(4) GGB_GEMV_CASE(8)
    default: return set_error(GGB_E_INVALID, "gemv: ncols=%d", b.ncols);
    }
#undef GGB_GEMV_CASE
    count_launch();
    return GGB_OK;
}

constexpr int X_BUDGET = 64 * 1024;        // activation columns resident in shared memory
constexpr int W_BUDGET = 144 * 1024;       // weight stages of all warps
constexpr int STAGE_MAX = 4096;            // bytes of one bulk copy (one row, several short rows, or a K-chunk of a long row)

} // namespace

int gemv_num_ctas() { return device_sm_count(); }

// Fills the shape-dependent fields of b (everything but the node list).  ncols in {1,2,4,8}.
int gemv_plan(GemvBatch &b, int type, int64_t K, int64_t nb01, int ncols, const void *Wbase)
{
    const int blck = blck_size(type);
    if (K <= 0 || K % blck) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld is not a multiple of the block size %d (Ggml.cs:6694)", (long long)K, blck);
    const long long row_bytes = (long long)(K / blck) * (long long)type_size(type);
    const long long xcol = (long long)act_row_bytes(type, K);
    if (row_bytes > (1ll << 30) || xcol * ncols > X_BUDGET + 32 * 1024)
        return set_error(GGB_E_UNSUPPORTED, "mul_mat: K=%lld too large for the shared-memory resident activation vector", (long long)K);
    b.type = type; b.ncols = ncols; b.K = (int)K; b.row_bytes = (int)row_bytes; b.nb01 = nb01; b.xcol_bytes = (int)xcol;
    const int unit_async = type == GGML_TYPE_Q4_0 ? 80 : type == GGML_TYPE_Q4_1 ? 48 : 16;
    const int unit_sync = type == GGML_TYPE_Q4_0 ? 20 : type == GGML_TYPE_Q4_1 ? 24 : 16;
    b.async = ((reinterpret_cast<uintptr_t>(Wbase) & 15) == 0 && (nb01 & 15) == 0 && (row_bytes & 15) == 0 &&
               (row_bytes <= STAGE_MAX || row_bytes % unit_async == 0)) ? 1 : 0;
    const int unit = b.async ? unit_async : unit_sync;
    if (row_bytes <= STAGE_MAX) {
        b.nchunk = 1;
        b.chunk_bytes = (int)row_bytes;
        // several short rows per copy only when rows are densely packed
        b.rs = (nb01 == row_bytes && b.async) ? (int)(STAGE_MAX / row_bytes) : 1;
        if (b.rs > 64) b.rs = 64;
        b.stage_bytes = (int)align_up((size_t)b.rs * row_bytes, 16);
    } else {
        const long long units = (row_bytes + unit - 1) / unit;
        long long nchunk = (row_bytes + STAGE_MAX - 1) / STAGE_MAX;
        const long long cu = (units + nchunk - 1) / nchunk;
        b.chunk_bytes = (int)(cu * unit);
        b.nchunk = (int)((row_bytes + b.chunk_bytes - 1) / b.chunk_bytes);
        b.rs = 1;
        b.stage_bytes = (int)align_up((size_t)b.chunk_bytes, 16);
    }
    const long long xbytes = (xcol * ncols + 127) & ~127ll;
    long long wbudget = 227 * 1024 - xbytes - NWARPS * MAX_DEPTH * 8 - 1024;
    if (wbudget > W_BUDGET) wbudget = W_BUDGET;
    int depth = b.async ? (int)(wbudget / ((long long)NWARPS * b.stage_bytes)) : 1;
    if (depth > MAX_DEPTH) depth = MAX_DEPTH;
    if (depth < 1) return set_error(GGB_E_UNSUPPORTED, "mul_mat: shape needs more shared memory than one SM has");
    b.depth = depth;
    return GGB_OK;
}

int launch_gemv_batch(const GemvBatch &b, cudaStream_t s, bool pdl)
{
    if (b.n_nodes <= 0 || b.total_groups <= 0) return GGB_OK;
    const size_t xbytes = ((size_t)b.ncols * b.xcol_bytes + 127) & ~(size_t)127;
    const size_t smem = xbytes + (size_t)NWARPS * b.depth * b.stage_bytes + NWARPS * MAX_DEPTH * 8;
    int grid = gemv_num_ctas();
    const int want = (b.total_groups + NWARPS - 1) / NWARPS;
    if (grid > want) grid = want;
    switch (b.type) {
    case GGML_TYPE_Q4_0: return launch_typed<GGML_TYPE_Q4_0>(b, smem, grid, s, pdl);
    case GGML_TYPE_Q4_1: return launch_typed<GGML_TYPE_Q4_1>(b, smem, grid, s, pdl);
    case GGML_TYPE_F16: return launch_typed<GGML_TYPE_F16>(b, smem, grid, s, pdl);
    case GGML_TYPE_F32: return launch_typed<GGML_TYPE_F32>(b, smem, grid, s, pdl);
    default: return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d is not on this path (Ggml.cs:6739-6742)", b.type);
    }
}

} // namespace ggb
