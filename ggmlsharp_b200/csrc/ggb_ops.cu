// The F32 neighbours of mul_mat in a Llama layer, so that consecutive MUL_MAT nodes of a graph stay
// on the device with no host round trip between them: ADD, MUL, SCALE, SILU, RMS_NORM and the strided copy behind
// cont(transpose(x)).  All are HBM-bound streams (4-12 bytes per element); the only arithmetic care is rounding:
//
//   ADD / MUL / SCALE   one IEEE binary32 operation per element, explicit _rn intrinsics (no FMA contraction)  -> bit-exact
//                       ggml_vec_add_f32 / mul_f32 / scale_f32, Ggml.cs:2586-2589, 2621-2624, 1416-1443
//   SILU                the reference is built with GGML_SILU_FP16 (GGMLSharp.csproj:9): y = (float)table_silu_f16[(Half)x]
//                       (Ggml.cs:2736-2746); the 64 K-entry table is built on the HOST exactly as ggml_init builds it
//                       (Ggml.cs:1461-1471: (Half)(f / (1.0f + expf(-f)))) and uploaded once -> bit-exact lookups
//   RMS_NORM            Ggml.cs:5858-5921: the row's sum of squares is accumulated in double (float products), but by a
//                       block reduction, not sequentially: the double sum can differ in its last bits, which survives the
//                       cast to float with probability ~2^-29 per row -> tolerance 1e-6 relative, in practice identical
//   DUP / CONT          element copy (Ggml.cs:4199-4398) -> bit-exact
#include "ggb_internal.h"

#include <cmath>
#include <mutex>

namespace ggb {

namespace {

// Every kernel here is a programmatic dependent of whatever precedes it in the stream, and lets its own successor be scheduled while
// it runs: first wait until the predecessor is COMPLETE (so is, transitively, everything before that), then release.  The successor's
// CTAs become resident and sit in their own griddepcontrol.wait until this grid is complete and flushed -- its launch latency, about
// as long as one of these kernels on a single-token row, is hidden.  Released only AFTER the wait: a successor that reads something
// without waiting (the weight copies of a GEMV / GEMM, unless GGB_MM_W_IN_FLIGHT tells them to) can then only run beside THIS
// kernel, never beside an older one.
__device__ __forceinline__ void pdl_wait_then_release()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// (no __restrict__: the in-place ADD / MUL of ggml_add_inplace and friends passes z == a, and silu / scale run in place too)
__global__ void __launch_bounds__(256) k_binary_f32(int op, const float *a, const float *b, float *z, long long n)
{
    pdl_wait_then_release();
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4 *>(a)[i], y = reinterpret_cast<const float4 *>(b)[i];
        float4 r;
        if (op == GGML_OP_ADD) { r.x = __fadd_rn(x.x, y.x); r.y = __fadd_rn(x.y, y.y); r.z = __fadd_rn(x.z, y.z); r.w = __fadd_rn(x.w, y.w); }
        else { r.x = __fmul_rn(x.x, y.x); r.y = __fmul_rn(x.y, y.y); r.z = __fmul_rn(x.z, y.z); r.w = __fmul_rn(x.w, y.w); }
        reinterpret_cast<float4 *>(z)[i] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        z[i] = op == GGML_OP_ADD ? __fadd_rn(a[i], b[i]) : __fmul_rn(a[i], b[i]);
    }
}

__global__ void __launch_bounds__(256) k_binary_f32_scalar(int op, const float *a, const float *b, float *z, long long n)
{
    pdl_wait_then_release();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        z[i] = op == GGML_OP_ADD ? __fadd_rn(a[i], b[i]) : __fmul_rn(a[i], b[i]);
}

// VEC: 16-byte aligned operands, four elements per thread and step (the scalar form moved 3.4 -- SILU -- and 4.3 TB/s -- SCALE -- on a
// 64 MB tensor against ADD / MUL's 6.7: profiles/r02_ops_throughput.txt); the tail of up to three elements goes element by element
template <bool VEC>
__global__ void __launch_bounds__(256) k_scale_f32(float *__restrict__ y, float v, long long n)
{
    pdl_wait_then_release();
    const long long n4 = VEC ? n >> 2 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 r = reinterpret_cast<float4 *>(y)[i];
        r.x = __fmul_rn(r.x, v); r.y = __fmul_rn(r.y, v); r.z = __fmul_rn(r.z, v); r.w = __fmul_rn(r.w, v);
        reinterpret_cast<float4 *>(y)[i] = r;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = __fmul_rn(y[i], v);
}

__device__ __forceinline__ float silu_lookup(float x, const unsigned short *__restrict__ table)
{
    const unsigned short h = __half_as_ushort(__float2half_rn(x));                   // (Half)x, round to nearest even
    return __half2float(__ushort_as_half(__ldg(table + h)));                         // (float)table_silu_f16[t]
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_silu_f32(const float *x, float *y, long long n, const unsigned short *__restrict__ table)
{
    pdl_wait_then_release();
    const long long n4 = VEC ? n >> 2 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4 *>(x)[i];
        float4 r;
        r.x = silu_lookup(a.x, table); r.y = silu_lookup(a.y, table); r.z = silu_lookup(a.z, table); r.w = silu_lookup(a.w, table);
        reinterpret_cast<float4 *>(y)[i] = r;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = silu_lookup(x[i], table);
}

// one CTA per row
__global__ void __launch_bounds__(256) k_rms_norm_f32(const float *x, long long x_stride, float *y, long long y_stride, int ne00)
{
    pdl_wait_then_release();
    const float *xr = x + (long long)blockIdx.x * x_stride;
    float *yr = y + (long long)blockIdx.x * y_stride;
    double sum = 0.0;
    // rows of up to 4096 elements stay in registers between the two passes (the second read of the row was the other half of this
    // kernel's traffic and latency); the order of the additions is the one it always was: thread t adds elements t, t + 256, ...
    constexpr int KEEP = 16;
    float keep[KEEP];
    const bool kept = ne00 <= KEEP * (int)blockDim.x;
    if (kept) {
#pragma unroll
        for (int k = 0; k < KEEP; k++) { const int i = threadIdx.x + k * (int)blockDim.x; keep[k] = i < ne00 ? xr[i] : 0.0f; }
#pragma unroll
        for (int k = 0; k < KEEP; k++) if ((int)threadIdx.x + k * (int)blockDim.x < ne00) sum += (double)__fmul_rn(keep[k], keep[k]);
    } else {
        for (int i = threadIdx.x; i < ne00; i += blockDim.x) { const float v = xr[i]; sum += (double)__fmul_rn(v, v); }   // float product, double accumulate
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    __shared__ double part[8];
    __shared__ float s_scale;
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        const float mean = (float)(t / (double)ne00);
        s_scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, 1e-6f)));
    }
    __syncthreads();
    const float scale = s_scale;
    if (kept) {
#pragma unroll
        for (int k = 0; k < KEEP; k++) { const int i = threadIdx.x + k * (int)blockDim.x; if (i < ne00) yr[i] = __fmul_rn(keep[k], scale); }
    } else {
        for (int i = threadIdx.x; i < ne00; i += blockDim.x) yr[i] = __fmul_rn(xr[i], scale);
    }
}

// dst (contiguous, [ne3][ne2][ne1][ne0]) <- strided src.  TRANSPOSED: the source's fast dimension is i1 (nb[1] == 4), the
// ggml_transpose view of a contiguous matrix: 32 x 32 tiles through shared memory so that both sides are coalesced.
struct DupArgs { long long ne[4]; long long nb[4]; };
__global__ void __launch_bounds__(256) k_dup_f32_transposed(const uint8_t *__restrict__ src, float *__restrict__ dst, const __grid_constant__ DupArgs a)
{
    pdl_wait_then_release();
    __shared__ float tile[32][33];
    const long long plane = blockIdx.z;
    const long long i3 = plane / a.ne[2], i2 = plane - i3 * a.ne[2];
    const uint8_t *sp = src + i2 * a.nb[2] + i3 * a.nb[3];
    float *dp = dst + plane * a.ne[0] * a.ne[1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                         // 32 x 8
    const long long i0b = (long long)blockIdx.x * 32, i1b = (long long)blockIdx.y * 32;
#pragma unroll
    for (int r = 0; r < 32; r += 8) {                                                // read: consecutive threads walk i1 (the source's fast axis)
        const long long i0 = i0b + ty + r, i1 = i1b + tx;
        if (i0 < a.ne[0] && i1 < a.ne[1]) tile[ty + r][tx] = *reinterpret_cast<const float *>(sp + i0 * a.nb[0] + i1 * a.nb[1]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {                                                // write: consecutive threads walk i0 (the destination's fast axis)
        const long long i1 = i1b + ty + r, i0 = i0b + tx;
        if (i0 < a.ne[0] && i1 < a.ne[1]) dp[i1 * a.ne[0] + i0] = tile[tx][ty + r];
    }
}
__global__ void __launch_bounds__(256) k_dup_f32_generic(const uint8_t *__restrict__ src, float *__restrict__ dst, const __grid_constant__ DupArgs a, long long n)
{
    pdl_wait_then_release();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const long long i0 = r % a.ne[0]; r /= a.ne[0];
        const long long i1 = r % a.ne[1]; r /= a.ne[1];
        const long long i2 = r % a.ne[2]; const long long i3 = r / a.ne[2];
        dst[i] = *reinterpret_cast<const float *>(src + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    }
}

// ggml_compute_forward_repeat_f32 (Ggml.cs:5340-5383), 2-D: dst[r][c] = src[r % nr0][c % nc0]
__global__ void __launch_bounds__(256) k_repeat_f32(const float *__restrict__ src, long long src_stride, int nc0, int nr0,
                                                    float *__restrict__ dst, long long dst_stride, int nc, long long n)
{
    pdl_wait_then_release();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / nc; const int c = (int)(i - r * nc);
        dst[r * dst_stride + c] = src[(r % nr0) * src_stride + (c % nc0)];
    }
}

unsigned short *g_silu_table = nullptr;
std::once_flag g_silu_once;
int g_silu_rc = GGB_OK;

// ggml_init's table loop (Ggml.cs:1461-1471) for table_silu_f16, on the host: MathF.Exp is the C runtime's expf
void build_silu_table()
{
    static unsigned short host[1 << 16];
    for (int i = 0; i < (1 << 16); i++) {
        __half_raw r; r.x = (unsigned short)i;
        const float f = __half2float(__half(r));
        const float e = expf(-f);
        const float den = 1.0f + e;
        const __half_raw o = __half_raw(__float2half_rn(f / den));
        host[i] = o.x;
    }
    if (cudaMalloc(reinterpret_cast<void **>(&g_silu_table), sizeof host) != cudaSuccess ||
        cudaMemcpy(g_silu_table, host, sizeof host, cudaMemcpyHostToDevice) != cudaSuccess) {
        g_silu_rc = set_error(GGB_E_CUDA, "silu table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_silu_table = nullptr;
    }
}

template <typename K, typename... A>
int launch_pdl(K kern, dim3 grid, dim3 block, cudaStream_t s, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    count_launch();
    return GGB_OK;
}

unsigned stream_grid(long long work_items, int per_cta)
{
    const long long want = (work_items + per_cta - 1) / per_cta;
    return (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)device_sm_count() * 16));
}

} // namespace

int launch_binary_f32(int op, const float *a, const float *b, float *dst, int64_t n, cudaStream_t s)
{
    if (op != GGML_OP_ADD && op != GGML_OP_MUL) return set_error(GGB_E_UNSUPPORTED, "binary op %d", op);
    if (n <= 0) return GGB_OK;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    if (vec) return launch_pdl(k_binary_f32, dim3(stream_grid(n / 4 + 1, 256)), dim3(256), s, op, a, b, dst, (long long)n);
    return launch_pdl(k_binary_f32_scalar, dim3(stream_grid(n, 256)), dim3(256), s, op, a, b, dst, (long long)n);
}

int launch_scale_f32(float *y, float v, int64_t n, cudaStream_t s)
{
    if (n <= 0) return GGB_OK;
    if ((reinterpret_cast<uintptr_t>(y) & 15) == 0) return launch_pdl(k_scale_f32<true>, dim3(stream_grid(n / 4 + 1, 256)), dim3(256), s, y, v, (long long)n);
    return launch_pdl(k_scale_f32<false>, dim3(stream_grid(n, 256)), dim3(256), s, y, v, (long long)n);
}

int silu_table_device(const unsigned short **table)
{
    std::call_once(g_silu_once, build_silu_table);
    if (!g_silu_table) return g_silu_rc ? g_silu_rc : set_error(GGB_E_CUDA, "silu table unavailable");
    *table = g_silu_table;
    return GGB_OK;
}

int launch_silu_f32(const float *x, float *y, int64_t n, cudaStream_t s)
{
    if (n <= 0) return GGB_OK;
    std::call_once(g_silu_once, build_silu_table);
    if (!g_silu_table) return g_silu_rc ? g_silu_rc : set_error(GGB_E_CUDA, "silu table unavailable");
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
        return launch_pdl(k_silu_f32<true>, dim3(stream_grid(n / 4 + 1, 256)), dim3(256), s, x, y, (long long)n, (const unsigned short *)g_silu_table);
    return launch_pdl(k_silu_f32<false>, dim3(stream_grid(n, 256)), dim3(256), s, x, y, (long long)n, (const unsigned short *)g_silu_table);
}

int launch_rms_norm_f32(const float *x, int64_t x_stride, float *y, int64_t y_stride, int64_t nrows, int64_t ne00, cudaStream_t s)
{
    if (nrows <= 0 || ne00 <= 0) return GGB_OK;
    if (nrows > 0x7fffffffLL || ne00 > 0x7fffffffLL) return set_error(GGB_E_UNSUPPORTED, "rms_norm: shape too large");
    return launch_pdl(k_rms_norm_f32, dim3((unsigned)nrows), dim3(256), s, x, (long long)x_stride, y, (long long)y_stride, (int)ne00);
}

int launch_repeat_f32(const float *src, int64_t src_stride, int64_t nc0, int64_t nr0, float *dst, int64_t dst_stride, int64_t nc, int64_t nr, cudaStream_t s)
{
    if (nc <= 0 || nr <= 0) return GGB_OK;
    if (nc0 <= 0 || nr0 <= 0 || nc % nc0 || nr % nr0) return set_error(GGB_E_INVALID, "repeat: ggml_can_repeat fails (Ggml.cs:8398-8407)");
    if (nc > 0x7fffffffLL || nr0 > 0x7fffffffLL) return set_error(GGB_E_UNSUPPORTED, "repeat: shape too large");
    return launch_pdl(k_repeat_f32, dim3(stream_grid(nc * nr, 256)), dim3(256), s, src, (long long)src_stride, (int)nc0, (int)nr0, dst, (long long)dst_stride, (int)nc, (long long)(nc * nr));
}

int launch_dup_f32_strided(const void *src, const int64_t ne[4], const uint64_t nb[4], float *dst, cudaStream_t s)
{
    DupArgs a;
    long long n = 1;
    for (int i = 0; i < 4; i++) { a.ne[i] = ne[i]; a.nb[i] = (long long)nb[i]; n *= ne[i]; }
    if (n <= 0) return GGB_OK;
    const long long planes = ne[2] * ne[3];
    if (nb[1] == 4 && nb[0] >= 4 * (uint64_t)ne[1] && planes <= 65535 && (ne[1] + 31) / 32 <= 65535)
        return launch_pdl(k_dup_f32_transposed, dim3((unsigned)((ne[0] + 31) / 32), (unsigned)((ne[1] + 31) / 32), (unsigned)planes), dim3(256), s,
                          static_cast<const uint8_t *>(src), dst, a);
    return launch_pdl(k_dup_f32_generic, dim3(stream_grid(n, 256)), dim3(256), s, static_cast<const uint8_t *>(src), dst, a, n);
}

} // namespace ggb
