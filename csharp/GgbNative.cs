// P/Invoke surface of libggb200.so for GGMLSharp (include/ggb200.h).  Drop this file into the GGMLSharp
// project; it adds no public API.  NOT compiled in this repository: no .NET toolchain exists in the build
// image.  The layouts it relies on are verified at run time by ggb_abi_check().
using System;
using System.Runtime.InteropServices;

namespace GGMLSharp;

internal static unsafe partial class GgbNative
{
    const string Lib = "ggb200";   // libggb200.so next to the application

    [DllImport(Lib)] public static extern IntPtr ggb_last_error();
    [DllImport(Lib)] public static extern int ggb_abi_check(int sizeofTensor, int offsetofData, int sizeofCgraph, int offsetofNodes, int sizeofQ4_0, int sizeofQ4_1);
    [DllImport(Lib)] public static extern int ggb_init();
    [DllImport(Lib)] public static extern int ggb_shutdown();
    [DllImport(Lib)] public static extern int ggb_pool_alloc(nuint bytes, void** hostBase, IntPtr* pool);
    [DllImport(Lib)] public static extern int ggb_pool_adopt(void* hostBase, nuint bytes, IntPtr* pool);
    [DllImport(Lib)] public static extern int ggb_pool_free(IntPtr pool);
    [DllImport(Lib)] public static extern int ggb_tensor_invalidate(IntPtr pool, ggml_tensor* t);
    // weight residency is opt-in: the reference re-reads src0->data on every compute, so the default re-uploads every leaf src0
    [DllImport(Lib)] public static extern int ggb_pool_set_weight_cache(IntPtr pool, int on);
    // row split of ggml_graph_compute across the GPUs of the box (off by default): max_devices < 0 = all, min_weight_bytes 0 = default 4 MiB
    [DllImport(Lib)] public static extern int ggb_pool_set_row_split(IntPtr pool, int maxDevices, nuint minWeightBytes);
    [DllImport(Lib)] public static extern int ggb_row_split_rows(long M, int g, int G, long* row0, long* rows);
    [DllImport(Lib)] public static extern int ggb_mul_mat_node(IntPtr pool, ggml_tensor* dst);
    [DllImport(Lib)] public static extern int ggb_graph_compute_mul_mats(IntPtr pool, ggml_cgraph* graph, int flags, byte* done);
    [DllImport(Lib)] public static extern int ggb_graph_plan(ggml_cgraph* graph, int flags, byte* done);
    // type = (int)ggml_type.  Weights: F32, F16, Q4_0, Q4_1 and the sibling formats Q4_2, Q5_0, Q5_1, Q8_0; their fp16 block scales are
    // IEEE bit patterns in memory -- declare block_q4_2.d and block_q5_1.d / .m as Half (as block_q5_0.d already is) instead of
    // storing (ushort)(Half)d, which writes a rounded integer (Ggml.cs:577, 678-679).
    [DllImport(Lib)] public static extern int ggb_quantize_rows(int type, float* src, void* dst, long nrows, long k);
    [DllImport(Lib)] public static extern int ggb_dequantize_rows(int type, void* src, float* dst, long nrows, long k);

    // flags of ggb_graph_compute_mul_mats (include/ggb200.h)
    public const int GGB_GRAPH_KEEP_ON_DEVICE = 1, GGB_GRAPH_NO_WEIGHT_CACHE = 2, GGB_GRAPH_MUL_MAT_ONLY = 4, GGB_GRAPH_SHARD = 8;

    // device-pointer entry points (what the executor itself calls); stream = a cudaStream_t or IntPtr.Zero for the library's own
    [DllImport(Lib)] public static extern int ggb_dev_binary(int op, float* a, float* b, float* dst, long n, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_scale(float* x, float v, long n, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_silu(float* x, float* dst, long n, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_rms_norm(float* x, long xStride, float* dst, long dstStride, long nrows, long ne00, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_repeat(float* src, long srcStride, long nc0, long nr0, float* dst, long dstStride, long nc, long nr, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_cont(void* src, long* ne, ulong* nb, float* dst, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_add_q(int type, void* src0, float* src1, void* dst, long nrows, long k, IntPtr stream);

    // the rest of include/ggb200.h (used by benchmarks and by a caller that keeps tensors on the device itself)
    [StructLayout(LayoutKind.Sequential)]
    public struct ggb_dev_mm
    {
        public int type, n_peers; public long M, K, N;
        public void* W; public long nb01; public float* X; public long ldx_bytes; public float* Y; public long ldy_bytes;
        public float* Y_peer0, Y_peer1, Y_peer2, Y_peer3, Y_peer4, Y_peer5, Y_peer6;
        public int* W_rowexp; public int flags, _pad;       // GGB_MM_W_IN_FLIGHT = 1
    }
    public const int GGB_MM_W_IN_FLIGHT = 1;
    public const int GGB_MM_X_HOST = 2;
    [DllImport(Lib)] public static extern int ggb_dev_weight_rowexp(int type, void* W, long nb01, long M, long K, int* rowexp, IntPtr stream);
    [StructLayout(LayoutKind.Sequential)]
    public struct ggb_stats
    {
        public ulong kernel_launches, h2d_bytes, d2h_bytes, weight_uploads, weight_cache_hits, nodes_executed;
        public double last_graph_device_ms, timed_kernel_ms; public ulong timed_kernel_launches, graph_replays;
    }
    [DllImport(Lib)] public static extern int ggb_abi_version();
    [DllImport(Lib)] public static extern int ggb_device_count(int* count);
    [DllImport(Lib)] public static extern nuint ggb_dev_workspace_bytes(ggb_dev_mm* mm, int count);
    [DllImport(Lib)] public static extern int ggb_dev_mul_mat_batch(ggb_dev_mm* mm, int count, void* workspace, nuint workspaceBytes, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_mul_mat_batch_phase(ggb_dev_mm* mm, int count, void* workspace, nuint workspaceBytes, IntPtr stream, int phase);
    [DllImport(Lib)] public static extern int ggb_dev_quantize_rows(int type, float* src, void* dst, long nrows, long k, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_dequantize_rows(int type, void* src, float* dst, long nrows, long k, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_dev_alloc(nuint bytes, void** dptr);
    [DllImport(Lib)] public static extern int ggb_dev_free(void* dptr);
    [DllImport(Lib)] public static extern int ggb_dev_upload(void* dptr, void* host, nuint bytes);
    [DllImport(Lib)] public static extern int ggb_dev_download(void* host, void* dptr, nuint bytes);
    [DllImport(Lib)] public static extern int ggb_stream_sync(IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_set_kernel_timing(int on);
    [DllImport(Lib)] public static extern int ggb_set_decode_program(int on);
    [DllImport(Lib)] public static extern int ggb_get_stats(ggb_stats* stats);
    [DllImport(Lib)] public static extern int ggb_reset_stats();
    // row split across the GPUs of one box (one process per GPU): CUDA-IPC export / open of the symmetric dst buffer and the
    // exchange kernels; see ggmlsharp_b200/rowsplit.py for the call sequence
    [DllImport(Lib)] public static extern int ggb_ipc_export(void* dptr, byte* handle64);
    [DllImport(Lib)] public static extern int ggb_ipc_open(byte* handle64, void** dptr);
    [DllImport(Lib)] public static extern int ggb_ipc_close(void* dptr);
    [DllImport(Lib)] public static extern int ggb_peer_barrier(ulong** peerFlags, int rank, int world, ulong epoch, IntPtr stream);
    [DllImport(Lib)] public static extern int ggb_peer_push_barrier(void** peerBases, ulong** peerFlags, uint* counter, int rank, int world,
                                                                   nuint segOffset, nuint segBytes, nuint segStride, int nSeg, ulong epoch, IntPtr stream);

    public static string LastError() => Marshal.PtrToStringUTF8(ggb_last_error()) ?? "";

    // ggml_context* -> ggb_pool*.  ggml_context itself (TypeDefinitions.cs:32-46) is left untouched.
    static readonly System.Collections.Generic.Dictionary<IntPtr, IntPtr> Pools = new();
    public static void Bind(ggml_context* ctx, IntPtr pool) { lock (Pools) Pools[(IntPtr)ctx] = pool; }
    public static IntPtr PoolOf(ggml_context* ctx) { lock (Pools) return Pools.TryGetValue((IntPtr)ctx, out var p) ? p : IntPtr.Zero; }
    public static IntPtr Unbind(ggml_context* ctx) { lock (Pools) { Pools.Remove((IntPtr)ctx, out var p); return p; } }
}
