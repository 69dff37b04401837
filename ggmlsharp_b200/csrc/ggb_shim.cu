// libggb200.so -- the C ABI of include/ggb200.h: device state, the pool behind a ggml_context,
// weight residency, and the CUDA-stream executor that replaces the reference's CPU thread pool for
// MUL_MAT graph nodes (ggml_graph_compute, Ggml.cs:3209-3736).
//
// There is no CPU fallback anywhere in this file: every compute entry point needs an sm_100 device.
#include "ggb_internal.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace ggb {

static thread_local char g_err[512] = "";
static std::mutex g_mu;
static bool g_inited = false;
static int g_device = 0, g_sms = 148;
static cudaStream_t g_stream = nullptr;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static ggb_stats g_stats = {};
static bool g_decode_program = getenv("GGB200_NO_DECODE_PROGRAM") == nullptr;      // ggb_set_decode_program
static int g_timing = 0;            // 0 off, 1 bracket every mul_mat kernel, 2 also skip the activation staging (reuse the previous call's)
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_timed;   // pending event pairs, resolved in ggb_get_stats

int set_error(int code, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { __atomic_fetch_add(&g_stats.kernel_launches, (uint64_t)n, __ATOMIC_RELAXED); }   // the row split launches from one host thread per GPU
struct KernelTimer {            // RAII bracket around one mul_mat kernel launch
    cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
    explicit KernelTimer(cudaStream_t st) : s(st) {
        if (g_timing == 2) { g_stats.timed_kernel_launches++; return; }      // mode 2: the caller brackets the whole stream itself
        if (!g_timing || g_timed.size() >= 4096) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, s);
    }
    ~KernelTimer() { if (a) { cudaEventRecord(b, s); g_timed.emplace_back(a, b); } }
};
int device_sm_count() { return g_sms; }

struct DevCtx {
    int dev = -1; cudaStream_t stream = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // second lane of the executor (run_nodes: a wide level of single-token mul_mats is dealt to two streams, so that the PCIe-side
    // kernels of one chunk -- activation staging from the pinned arena, result copy back into it -- run beside the GEMV of the other)
    cudaStream_t stream2 = nullptr; cudaEvent_t fork = nullptr, join = nullptr;
    // one event per dependency level, created on demand by the device's own host thread and read by the others after a rendezvous:
    // a fixed table, so that growing it never moves what a peer is reading
    cudaEvent_t level_ev[GGML_MAX_NODES + 1] = {};
    std::vector<cudaEvent_t> level_time;      // timing events, one after each dependency level of an eager compute (per-node perf_time_us)
    unsigned *dp_bar = nullptr;               // grid-barrier counter of the decode program (run_nodes)
    bool dp_silu_ready = false;
};

static std::vector<DevCtx> g_devs;            // [0] = the library's own device (g_device / g_stream / g_ev0 / g_ev1); [1..] opened by ensure_multi
static int g_multi = 0;                       // devices opened with mutual peer access (0 = not probed yet)
static std::mutex g_init_mu;      // ggb_dev_* entry points reach ensure_init without g_mu
static int ensure_init()
{
    std::lock_guard<std::mutex> lk_init(g_init_mu);
    if (g_inited) { cudaSetDevice(g_device); return GGB_OK; }      // the calling host thread may be new: make the library's device current
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(GGB_E_NODEVICE, "no CUDA device (%s); libggb200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    int dev = 0;
    if (const char *s = getenv("GGB200_DEVICE")) dev = atoi(s);
    else if (const char *lr = getenv("LOCAL_RANK")) dev = atoi(lr) % n;      // one process per GPU under torchrun
    if (dev < 0 || dev >= n) return set_error(GGB_E_NODEVICE, "GGB200_DEVICE=%d but %d device(s) present", dev, n);
    GGB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp p;
    GGB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) return set_error(GGB_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
    g_device = dev; g_sms = p.multiProcessorCount;
    GGB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    GGB_CUDA(cudaEventCreate(&g_ev0));
    GGB_CUDA(cudaEventCreate(&g_ev1));
    g_devs.clear();
    g_devs.emplace_back();
    DevCtx &d0 = g_devs[0];
    d0.dev = dev; d0.stream = g_stream; d0.ev0 = g_ev0; d0.ev1 = g_ev1;
    // Lane 2 gets the HIGHEST priority: when its small PCIe-side kernels are issued, the CTAs of the next GEMV are often already pending
    // (programmatic launch), and the block scheduler serves pending CTAs of equal priority in order -- a 128-thread staging CTA would
    // wait a whole GEMV behind CTAs that cannot be placed yet (benchmarks/coresidency_probe.py: 50 us at equal priority, 19 us = launch
    // + run at high priority)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    GGB_CUDA(cudaStreamCreateWithPriority(&d0.stream2, cudaStreamNonBlocking, prio_hi));
    GGB_CUDA(cudaEventCreateWithFlags(&d0.fork, cudaEventDisableTiming));
    GGB_CUDA(cudaEventCreateWithFlags(&d0.join, cudaEventDisableTiming));
    g_multi = 0;
    g_inited = true;
    return GGB_OK;
}

// ------------------------------------------------------------------------------------------------
// The GPUs of one box behind ONE process (north_star: "large weight matrices are row-split across the GPUs of one 8xB200 box").
// The reference splits the rows of src0 over the OS threads of ggml_graph_compute (Ggml.cs:3231-3252, 6665-6672); here every
// "thread" of that split is a host thread that drives one GPU.  Device 0 of the set is the library's own device (ensure_init).
// ------------------------------------------------------------------------------------------------
struct Worker {
    std::thread th; std::mutex m; std::condition_variable cv;
    std::function<void()> job; bool has_job = false, done = false, quit = false;
    void loop() {
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<void()> j = std::move(job);
            has_job = false;
            lk.unlock();
            j();
            lk.lock();
            done = true;
            cv.notify_all();
        }
    }
};
// g_workers[g - 1] drives device g; the calling thread drives device 0.  Deliberately leaked: at process exit the threads sit in
// cv.wait, and destroying a condition variable with waiters (a static destructor would) blocks forever in pthread_cond_destroy.
// ggb_shutdown stops and joins them properly.
static std::vector<Worker *> &g_workers = *new std::vector<Worker *>();

// Opens up to `want` sm_100 devices (GGB200_DEVICES = comma list, default: the primary device, then the others in index order) and
// enables peer access between every pair.  Returns how many are usable together (>= 1).  Under torchrun (LOCAL_RANK set: one process
// per GPU) the other GPUs belong to other ranks, so nothing beyond the primary is opened unless GGB200_DEVICES says so.
static int ensure_multi(int want)
{
    if (g_multi) return std::min(g_multi, std::max(want, 1));
    std::vector<int> ids{g_device};
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 1; }
    if (const char *e = getenv("GGB200_DEVICES")) {
        ids.clear();
        for (const char *p = e; *p;) { char *end; long v = strtol(p, &end, 10); if (end == p) break; if (v >= 0 && v < n) ids.push_back((int)v); p = *end ? end + 1 : end; }
        if (ids.empty() || ids[0] != g_device) ids.insert(ids.begin(), g_device);
    } else if (!getenv("LOCAL_RANK")) {
        for (int d = 0; d < n; d++) if (d != g_device) ids.push_back(d);
    }
    if ((int)ids.size() > 8) ids.resize(8);
    g_devs.resize(1);                                        // [0] was set up by ensure_init
    for (size_t i = 1; i < ids.size(); i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, ids[i]) != cudaSuccess || p.major != 10) { cudaGetLastError(); continue; }
        bool ok = true;
        for (const DevCtx &o : g_devs) {                      // mutual peer access with every device already in the set
            int a = 0, b = 0;
            if (cudaDeviceCanAccessPeer(&a, ids[i], o.dev) != cudaSuccess || cudaDeviceCanAccessPeer(&b, o.dev, ids[i]) != cudaSuccess || !a || !b) { cudaGetLastError(); ok = false; break; }
        }
        if (!ok) continue;
        DevCtx dc; dc.dev = ids[i];
        if (cudaSetDevice(dc.dev) != cudaSuccess) { cudaGetLastError(); continue; }
        for (const DevCtx &o : g_devs) { cudaError_t e1 = cudaDeviceEnablePeerAccess(o.dev, 0); if (e1 != cudaSuccess && e1 != cudaErrorPeerAccessAlreadyEnabled) ok = false; cudaGetLastError(); }
        if (ok && (cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&dc.ev0) != cudaSuccess || cudaEventCreate(&dc.ev1) != cudaSuccess ||
                   cudaStreamCreateWithPriority(&dc.stream2, cudaStreamNonBlocking, -100 /* clamped to the highest */) != cudaSuccess || cudaEventCreateWithFlags(&dc.fork, cudaEventDisableTiming) != cudaSuccess ||
                   cudaEventCreateWithFlags(&dc.join, cudaEventDisableTiming) != cudaSuccess)) { cudaGetLastError(); ok = false; }
        if (ok) for (const DevCtx &o : g_devs) {
            cudaSetDevice(o.dev);
            cudaError_t e1 = cudaDeviceEnablePeerAccess(dc.dev, 0);
            if (e1 != cudaSuccess && e1 != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
        if (ok) g_devs.push_back(dc);
    }
    cudaSetDevice(g_device);
    g_multi = (int)g_devs.size();
    while ((int)g_workers.size() < g_multi - 1) {
        Worker *w = new Worker();
        g_workers.push_back(w);
        w->th = std::thread([w] { w->loop(); });
    }
    if (getenv("GGB200_VERBOSE")) fprintf(stderr, "[ggb200] row split: %d device(s) with mutual peer access\n", g_multi);
    return std::min(g_multi, std::max(want, 1));
}

static void sync_all_devices()
{
    if (!g_inited) return;
    for (const DevCtx &d : g_devs) { cudaSetDevice(d.dev); cudaStreamSynchronize(d.stream); cudaStreamSynchronize(d.stream2); }
    cudaSetDevice(g_device);
}
// fn(g) on G host threads, g = 0 on the caller's; returns when all are done
static void run_parallel(int G, const std::function<void(int)> &fn)
{
    for (int g = 1; g < G; g++) {
        Worker *w = g_workers[(size_t)g - 1];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [&fn, g] { fn(g); };
        w->has_job = true; w->done = false;
        w->cv.notify_all();
    }
    fn(0);
    for (int g = 1; g < G; g++) {
        Worker *w = g_workers[(size_t)g - 1];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
    }
}

// spinning rendezvous of the G host threads of one sharded call; fail() releases everybody
struct HostBarrier {
    int n = 1; std::atomic<int> count{0}, gen{0}, failed{0};
    bool arrive_and_wait()
    {
        const int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { count.store(0, std::memory_order_relaxed); gen.fetch_add(1, std::memory_order_release); }
        else while (gen.load(std::memory_order_acquire) == g && !failed.load(std::memory_order_acquire)) std::this_thread::yield();
        return !failed.load(std::memory_order_acquire);
    }
    void fail() { failed.store(1, std::memory_order_release); }
};
static const bool g_trace_shard = getenv("GGB200_TRACE_SHARD") != nullptr;
#define SHARD_TRACE(...) do { if (g_trace_shard) { fprintf(stderr, "[ggb200 shard] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)
struct ShardShared { int G = 1; HostBarrier bar; uint8_t *sym_base[8] = {}; int rc[8] = {}; char err[8][256] = {}; };
struct ShardCtx { int g, G; ShardShared *sh; };

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------------------------------------
// device-level batch: plan + launch
// ------------------------------------------------------------------------------------------------

// Sibling formats on the batched path.  Q4_2 and Q5_1 have the K-step footprint of Q4_0 / Q4_1 (80 / 96 bytes per 128 weights)
// and Q8_0's is 144 bytes, so the tcgen05 kernels dequantize them in flight like those; Q5_0's 88 bytes ride in a 96-byte box
// that advances by 88 (ggb_tc.cuh: RawRow).  Shapes the TMA path cannot take (K not a multiple of 128, unaligned rows) -- are expanded to dense fp16 [M][K] in the workspace
// (k_expand_f16) and the F16 kernel multiplies that, against activations staged as d*q like every quantized type.
static inline bool use_gemm_expanded(const ggb_dev_mm &m)
{
    return is_q_weight(m.type) && m.N >= 16 && m.M > 0 && m.K % GGB_QK == 0 && ((reinterpret_cast<uintptr_t>(m.W) | (uintptr_t)m.nb01) & 1) == 0 &&
           !gemm_supported(m.type, m.M, m.K, m.N, m.nb01, m.W);
}
static inline bool use_gemm(const ggb_dev_mm &m)
{
    return m.N >= 16 && (use_gemm_expanded(m) || gemm_supported(m.type, m.M, m.K, m.N, m.nb01, m.W));
}
static inline size_t expanded_bytes(int64_t M, int64_t K) { return align_up((size_t)M * (size_t)K * 2, 256); }
// a batched quantized node's slice: [fp16 activations][ex][ew] (gemm_workspace_bytes), then the expanded weights if the shape needs them
// Rows of any length (the reference loops over any ne00, Ggml.cs:6139-6164, 6676-6699).  The GEMV keeps one activation
// row in shared memory next to its weight stages, which bounds K (gemv_x_budget); a longer row is multiplied in K SEGMENTS of
// GGB_KSEG elements -- each an ordinary node over a column slice of W (same row stride) and of x (the Q8 blocks of a slice are the
// blocks of the whole row) writing its partial sums into the workspace -- and k_sum_segments adds the partials in segment order.
constexpr int64_t GGB_KSEG = 16384;           // 10 / 12 / 11 / 18 / 32 / 64 KB of a Q4_0 / Q4_1 / Q5_0 / Q8_0 / F16 / F32 row: whole units, 16-byte aligned
static inline bool needs_k_segments(const ggb_dev_mm &m)
{
    return m.M > 0 && m.N > 0 && !use_gemm(m) && (int64_t)act_row_bytes(m.type, m.K) > gemv_x_budget();
}
static inline int64_t k_segments(int64_t K) { return (K + GGB_KSEG - 1) / GGB_KSEG; }
static size_t mm_ws_bytes_plain(const ggb_dev_mm &m)
{
    if (use_gemm_expanded(m)) return align_up(gemm_workspace_bytes(m.type, m.M, m.K, m.N), 256) + expanded_bytes(m.M, m.K);
    if (use_gemm(m)) return align_up(gemm_workspace_bytes(m.type, m.M, m.K, m.N), 256);
    return align_up((size_t)m.N * act_row_bytes(m.type, m.K), 256);
}
static size_t mm_ws_bytes(const ggb_dev_mm &m)
{
    if (!needs_k_segments(m)) return mm_ws_bytes_plain(m);
    size_t t = align_up((size_t)k_segments(m.K) * (size_t)m.N * (size_t)m.M * 4, 256);      // the partial sums [segment][N][M]
    for (int64_t k0 = 0; k0 < m.K; k0 += GGB_KSEG) { ggb_dev_mm seg = m; seg.K = std::min(GGB_KSEG, m.K - k0); t += mm_ws_bytes_plain(seg); }
    return t;
}
// upper bound that does not depend on operand addresses (for sizing before buffers exist)
// may_expand: the weights might not satisfy the TMA path's alignment once staged (the caller knows nb01 and view offsets)
static size_t mm_ws_bytes_bound(int type, int64_t M, int64_t K, int64_t N, bool may_expand)
{
    const bool expand = is_q_weight(type) && N >= 16 && (may_expand || K % 128 != 0);
    const size_t tc = align_up(gemm_workspace_bytes(type, M, K, N), 256) + (expand ? expanded_bytes(M, K) : 0);      // incl. the exponent arrays
    size_t gemv = align_up((size_t)N * act_row_bytes(type, K), 256);
    if ((int64_t)act_row_bytes(type, K) > gemv_x_budget())      // K segments: partial sums + one staged slice per segment
        gemv = align_up((size_t)k_segments(K) * (size_t)N * (size_t)M * 4, 256) + (size_t)k_segments(K) * align_up((size_t)N * act_row_bytes(type, std::min(GGB_KSEG, K)), 256);
    return std::max(tc, gemv);
}

static int check_mm(const ggb_dev_mm &m)
{
    if (!is_mm_weight(m.type))
        return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d has no vec_dot (F32, F16, Q4_0, Q4_1, Q4_2, Q5_0, Q5_1, Q8_0; Ggml.cs:219-282)", m.type);
    if (m.M < 0 || m.N < 0 || m.K <= 0) return set_error(GGB_E_INVALID, "mul_mat: bad shape M=%lld N=%lld K=%lld", (long long)m.M, (long long)m.N, (long long)m.K);
    if (m.K % blck_size(m.type) || (is_q_weight(m.type) && m.K % GGB_QK)) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld %% 32 != 0 (Ggml.cs:6694)", (long long)m.K);
    const int64_t rb = m.K / blck_size(m.type) * (int64_t)type_size(m.type);
    if (m.nb01 < rb) return set_error(GGB_E_INVALID, "mul_mat: nb01=%lld < row bytes %lld", (long long)m.nb01, (long long)rb);
    if ((m.ldx_bytes & 3) || (m.ldy_bytes & 3) || m.ldx_bytes < 4 * m.K || (m.N > 1 && m.ldy_bytes < 4 * m.M))
        return set_error(GGB_E_INVALID, "mul_mat: bad activation/output row stride");
    if (m.n_peers < 0 || m.n_peers > 7) return set_error(GGB_E_INVALID, "mul_mat: n_peers=%d", m.n_peers);
    if (m.M > 0 && m.N > 0 && (!m.W || !m.X || !m.Y)) return set_error(GGB_E_INVALID, "mul_mat: null operand");
    return GGB_OK;
}

// ggb_dev_mul_mat_batch_phase: 0 = stage the activations and multiply (the default), 1 = stage only, 2 = multiply what an earlier
// staging call with the same arguments left in the workspace, 3 = stage only and do not wait for the preceding kernel of the stream
// (the caller vouches for the activations and the workspace).  Single-token nodes only.
static thread_local int tl_phase = 0;

static int dev_batch_inner(const ggb_dev_mm *mm, int count, void *ws, size_t ws_bytes, cudaStream_t s)
{
    if (count <= 0) return GGB_OK;
    std::vector<size_t> off(count + 1, 0);
    for (int i = 0; i < count; i++) {
        int rc = check_mm(mm[i]);
        if (rc) return rc;
        off[i + 1] = off[i] + mm_ws_bytes_plain(mm[i]);
    }
    if (off[count] > ws_bytes) return set_error(GGB_E_INVALID, "mul_mat: workspace %zu B < required %zu B", ws_bytes, off[count]);
    if (off[count] && (reinterpret_cast<uintptr_t>(ws) & 255)) return set_error(GGB_E_INVALID, "mul_mat: workspace must be 256-byte aligned");
    uint8_t *wsb = static_cast<uint8_t *>(ws);

    // ---- batched (tensor-core) nodes.  Every kernel is launched programmatically dependent on its predecessor; a GEMM starts its
    //      prologue and its weight streaming while the activation kernel in front of it still runs, and waits for it only before
    //      the first activation tile.  Activation kernels wait at their start for everything before them (see below) ----
    std::vector<int> gemv_idx;
    static const bool grouped = [] { const char *e = getenv("GGB200_GEMM_GROUPED"); return !e || atoi(e) != 0; }();
    constexpr int NSLOT = 8;                                      // one grouped launch sequence per kernel flavour
    static const int slot_type[NSLOT] = {GGML_TYPE_Q4_0, GGML_TYPE_Q4_1, GGML_TYPE_F16, GGML_TYPE_F16 /* expanded siblings */, GGML_TYPE_Q4_2, GGML_TYPE_Q5_1, GGML_TYPE_Q8_0, GGML_TYPE_Q5_0};
    std::vector<int> q_nodes[NSLOT];
    auto qslot = [&](const ggb_dev_mm &m) {
        if (use_gemm_expanded(m)) return 3;
        switch (m.type) { case GGML_TYPE_Q4_0: return 0; case GGML_TYPE_Q4_1: return 1; case GGML_TYPE_Q4_2: return 4; case GGML_TYPE_Q5_1: return 5; case GGML_TYPE_Q8_0: return 6; case GGML_TYPE_Q5_0: return 7; default: return 2; }
    };
    // a node whose shape the TMA kernels cannot take: expand the weights behind the activation buffer of its workspace slice and hand
    // the GEMM an F16 node.  The expansion is an ordinary (fully stream-ordered) launch that never releases its dependents early, so
    // nothing behind it in the stream -- in particular a GEMM, whose weight TMA does not wait for anything -- starts before it is complete.
    // power-of-two row exponents (ggb_internal.h: launch_weight_rowexp): the caller's, computed while the weights became resident,
    // or computed here -- ordinary launches ahead of everything else of the batch, so they are complete before any kernel below starts
    std::vector<const int *> ew_of(count, nullptr);
    std::vector<char> w_wait(count, 0);
    {
        static thread_local RowExpBatch rb;
        rb.n_nodes = 0;
        auto flush = [&]() -> int { if (!rb.n_nodes) return GGB_OK; int r = launch_weight_rowexp_batch(rb, s); rb.n_nodes = 0; return r; };
        for (int i = 0; i < count; i++) {
            const ggb_dev_mm &m = mm[i];
            if (m.M == 0 || m.N == 0 || !use_gemm(m) || !is_q_weight(m.type)) continue;
            w_wait[i] = (m.flags & GGB_MM_W_IN_FLIGHT) ? 1 : 0;
            if (m.W_rowexp) { ew_of[i] = m.W_rowexp; continue; }
            int *ew = reinterpret_cast<int *>(wsb + off[i] + gemm_ws_ew_offset(m.K, m.N));
            rb.node[rb.n_nodes++] = RowExpNode{static_cast<const uint8_t *>(m.W), (long long)m.nb01, ew, 0, (int)m.M, (int)(m.K / GGB_QK), m.type, 0};
            ew_of[i] = ew; w_wait[i] = 1;
            if (rb.n_nodes == GGB_GEMM_GROUP_NODES) { int rc = flush(); if (rc) return rc; }
        }
        int rc = flush();
        if (rc) return rc;
    }
    auto ex_of = [&](int i) { return is_q_weight(mm[i].type) ? reinterpret_cast<int *>(wsb + off[i] + gemm_ws_ex_offset(mm[i].K, mm[i].N)) : nullptr; };
    auto expand = [&](int i, const void *&Wout, int64_t &nb01_out) -> int {
        const ggb_dev_mm &m = mm[i];
        __half *wh = reinterpret_cast<__half *>(wsb + off[i] + align_up(gemm_workspace_bytes(m.type, m.M, m.K, m.N), 256));
        Wout = wh; nb01_out = 2 * m.K;
        return launch_expand_f16(m.type, m.W, m.nb01, wh, m.M, m.K, ew_of[i], s, false);
    };
    int n_tc = 0;
    for (int i = 0; i < count; i++) if (mm[i].M > 0 && mm[i].N > 0 && use_gemm(mm[i])) n_tc++;

    for (int i = 0; i < count; i++) {
        const ggb_dev_mm &m = mm[i];
        if (m.M == 0 || m.N == 0) continue;
        if (!use_gemm(m)) { gemv_idx.push_back(i); continue; }
        // a lone node keeps the per-node kernel (smaller launch); two or more go through the persistent grouped kernels
        const bool sibx = use_gemm_expanded(m);
        const int gtype = sibx ? GGML_TYPE_F16 : m.type;          // what the tensor-core kernel sees
        if (grouped && n_tc > 1 && !getenv("GGB200_GEMM_TRACE") && gemm_grouped_supported(gtype)) { q_nodes[qslot(m)].push_back(i); continue; }
        const int64_t Npad = (m.N + 15) / 16 * 16;
        __half *xh = reinterpret_cast<__half *>(wsb + off[i]);
        const void *Wg = m.W; int64_t nb01g = m.nb01;
        int rc = sibx ? expand(i, Wg, nb01g) : GGB_OK;
        if (rc) return rc;
        rc = launch_act_f16_dequant(m.type, gemm_act_perm(gtype), m.X, m.ldx_bytes, xh, ex_of(i), m.N, Npad, m.K, s, true);   // waits for whatever produced X
        if (rc) return rc;
        GemmArgs a = {};
        a.type = gtype; a.M = m.M; a.K = m.K; a.N = m.N; a.W = Wg; a.nb01 = nb01g; a.Xh = xh; a.Npad = Npad;
        a.ew = ew_of[i]; a.ex = ex_of(i); a.wait_w = w_wait[i];
        a.Y = m.Y; a.ldy = m.ldy_bytes / 4; a.n_peers = m.n_peers;
        for (int p = 0; p < m.n_peers; p++) a.ypeer[p] = m.Y_peer[p];
        if (const char *tr = getenv("GGB200_GEMM_TRACE")) a.trace = reinterpret_cast<void *>(strtoull(tr, nullptr, 0));   // debugging: device pointer
        { KernelTimer kt(s); rc = launch_gemm(a, nullptr, s); }
        if (rc) return rc;
    }
    // ---- two or more batched nodes: per weight type (kernel flavour) one activation launch + one persistent grouped GEMM launch
    //      per <= 64 nodes (ggb_gemm_grouped.cu) ----
    // Nodes that multiply the SAME activations (wq / wk / wv of a layer, w1 / w3 of its FFN) share one staged copy: the first
    // node's buffer is staged, the others point their B operand at it.  Keyed by what decides the staged bytes.
    struct StagedX { const float *X; int64_t ldx, N, K; int cls; __half *xh; int *ex; };
    std::vector<StagedX> staged;
    // Launch order: the activation kernels of ALL groups first, then the GEMMs.  Every activation kernel waits at its start
    // (griddepcontrol.wait): the first for whatever produced the activations, each later one for its predecessor -- so when the
    // last one is complete, all are.  (They used to be interleaved, act(g) GEMM(g) act(g+1) ..., with only the first one waiting:
    // a later one, launched programmatically dependent on a GEMM that triggers its dependents at start-up, could then read
    // activations an earlier graph level was still writing.  tests/test_gpu_graph_fuzz.py caught it as a rare wrong result.)
    struct PendingGroup { std::vector<GemmArgs> ga; };
    std::vector<PendingGroup> pending;
    for (int qi = 0; qi < NSLOT; qi++) {
        const std::vector<int> &qn = q_nodes[qi];
        for (size_t c0 = 0; c0 < qn.size(); c0 += GGB_GEMM_GROUP_NODES) {
            const int cnt = (int)std::min(qn.size() - c0, (size_t)GGB_GEMM_GROUP_NODES);
            static thread_local ActGemmBatch ab;
            pending.emplace_back();
            std::vector<GemmArgs> &ga = pending.back().ga;
            ga.resize((size_t)cnt);
            const int type = slot_type[qi];                                                             // slot 3: expanded siblings run the F16 kernel
            // slot 3 stages activations as d*q (any quantized wtype selects that), in the F16 kernel's natural K order
            ab.n_nodes = 0; ab.wtype = qi == 3 ? GGML_TYPE_Q8_0 : type; ab.perm = gemm_act_perm(type); ab.wait_prior = 1;
            const int cls = ab.wtype == GGML_TYPE_F16 ? 2 : ab.perm;      // (Half)x | d*q in natural K order | d*q in the nibble-unpack order
            for (int c = 0; c < cnt; c++) {
                const int i = qn[c0 + c];
                const ggb_dev_mm &m = mm[i];
                const int64_t Npad = (m.N + 15) / 16 * 16;
                __half *xh = nullptr; int *exs = nullptr;
                for (const StagedX &sx : staged)
                    if (sx.X == m.X && sx.ldx == m.ldx_bytes && sx.N == m.N && sx.K == m.K && sx.cls == cls) { xh = sx.xh; exs = sx.ex; break; }
                if (!xh) {
                    xh = reinterpret_cast<__half *>(wsb + off[i]); exs = ex_of(i);
                    ab.node[ab.n_nodes++] = ActGemmNode{m.X, (long long)m.ldx_bytes, xh, exs, (int)m.N, (int)Npad, (int)m.K, 0};
                    staged.push_back(StagedX{m.X, m.ldx_bytes, m.N, m.K, cls, xh, exs});
                }
                GemmArgs &a = ga[(size_t)c];
                a = GemmArgs{};
                a.type = type; a.M = m.M; a.K = m.K; a.N = m.N; a.W = m.W; a.nb01 = m.nb01; a.Xh = xh; a.Npad = Npad;
                a.Y = m.Y; a.ldy = m.ldy_bytes / 4; a.n_peers = m.n_peers;
                a.ew = ew_of[i]; a.ex = exs; a.wait_w = w_wait[i];
                for (int p = 0; p < m.n_peers; p++) a.ypeer[p] = m.Y_peer[p];
                if (qi == 3) { int rce = expand(i, a.W, a.nb01); if (rce) return rce; }      // an ordinary launch: a full barrier in the stream
            }
            int rc = launch_act_f16_dequant_batch(ab, s);               // no-op when every node of the group reuses staged activations
            if (rc) return rc;
        }
    }
    for (PendingGroup &pg : pending) {
        int rc;
        { KernelTimer kt(s); rc = launch_gemm_grouped(pg.ga.data(), (int)pg.ga.size(), s); }
        if (rc) return rc;
    }

    // ---- single-token nodes: group by (type, K), fuse each group into one act launch + GEMV launches ----
    std::vector<char> done(count, 0);
    std::vector<uint8_t *> act_base(count);
    for (int i = 0; i < count; i++) act_base[i] = wsb + off[i];
    for (size_t a0 = 0; a0 < gemv_idx.size(); a0++) {
        const int i0 = gemv_idx[a0];
        if (done[i0]) continue;
        // the activation layout depends on which GEMV kernel will read it, so that is part of the key
        auto act_bps = [&](const ggb_dev_mm &m, int &bps) -> int {
            GemvHdr probe = {};
            int r = gemv_plan(probe, m.type, m.K, m.nb01, 1, m.W);
            bps = r ? 0 : gemv_act_bps(probe);
            return r;
        };
        int bps0 = 1;
        { int r = act_bps(mm[i0], bps0); if (r) return r; }
        std::vector<int> grp;
        for (size_t a1 = a0; a1 < gemv_idx.size(); a1++) {
            const int i = gemv_idx[a1];
            if (done[i] || mm[i].type != mm[i0].type || mm[i].K != mm[i0].K) continue;
            int bps = 1;
            { int r = act_bps(mm[i], bps); if (r) return r; }
            if (bps == bps0) { grp.push_back(i); done[i] = 1; }
        }
        const int type = mm[i0].type; const int64_t K = mm[i0].K;
        const size_t arow = act_row_bytes(type, K);
        const bool quant = is_q_weight(type);
        // Small decode levels (a dependent chain runs one to three single-token nodes per level): the GEMV quantizes the activation
        // row in its own prologue and the staging launch in front of it -- half of such a level's launches -- goes away.  Larger
        // batches keep the staged copy: there one staging launch serves many nodes and every CTA would repeat the row's conversion.
        static const int fuse_max = [] { const char *e = getenv("GGB200_GEMV_FUSE_MAX"); return e ? atoi(e) : 8; }();
        bool fuse_x = quant && tl_phase == 0 && g_timing != 2 && (int)grp.size() <= fuse_max;
        for (size_t c = 0; c < grp.size() && fuse_x; c++) {
            const ggb_dev_mm &m = mm[grp[c]];
            GemvHdr probe = {};
            if (m.N != 1 || (m.flags & GGB_MM_X_HOST) || (reinterpret_cast<uintptr_t>(m.X) & 15) || gemv_plan(probe, type, K, m.nb01, 1, m.W) || !gemv_can_fuse_x(probe)) fuse_x = false;
        }
        // activation staging (INIT phase)
        for (size_t c0 = 0; c0 < grp.size(); c0 += GGB_MAX_BATCH_NODES) {
            static thread_local ActBatch ab;
            static_cast<ActHdr &>(ab) = ActHdr{};
            ab.K = (int)K; ab.kb = (int)(K / GGB_QK); ab.row_bytes = (int)arow; ab.wtype = type; ab.vec16 = 1; ab.bps = bps0; ab.no_wait = tl_phase == 3 ? 1 : 0;
            int tot = 0;
            for (size_t c = c0; c < std::min(grp.size(), c0 + (size_t)GGB_MAX_BATCH_NODES); c++) {
                const ggb_dev_mm &m = mm[grp[c]];
                // same activations as an earlier node of this (type, K, layout) group: reuse its staged rows
                bool shared = false;
                for (size_t e = 0; e < c && !shared; e++) {
                    const ggb_dev_mm &o = mm[grp[e]];
                    if (o.X == m.X && o.ldx_bytes == m.ldx_bytes && o.N == m.N) { act_base[grp[c]] = act_base[grp[e]]; shared = true; }
                }
                if (shared) continue;
                ActNode &an = ab.node[ab.n_nodes++];
                an.x = m.X; an.ldx_bytes = m.ldx_bytes; an.out = wsb + off[grp[c]]; an.N = (int)m.N; an.blk0 = tot;
                tot += quant ? (int)(m.N * ab.kb) : (int)m.N;
                if ((reinterpret_cast<uintptr_t>(m.X) & 15) || (m.ldx_bytes & 15)) ab.vec16 = 0;
            }
            ab.total_blk = tot;
            if (g_timing == 2 || tl_phase == 2 || !ab.n_nodes || fuse_x) continue;         // the workspace already holds these activations / the GEMV stages them itself
            int rc = launch_act_batch(ab, s, true);
            if (rc) return rc;
        }
        // column passes of 8/4/2/1, grouped by what must be uniform inside one launch
        struct Pass { int i, col0, nc; };
        std::vector<Pass> passes;
        // ... as many columns per pass as fit beside the weight stages in shared memory (K = 11008: 4 Q8P columns, 2 fp16, 2 fp32)
        int nc_cap = 8;
        while (nc_cap > 1 && (int64_t)arow * nc_cap > gemv_x_budget()) nc_cap >>= 1;
        for (int i : grp) {
            int64_t c = 0;
            while (c < mm[i].N) {
                int nc = mm[i].N - c >= 8 ? 8 : mm[i].N - c >= 4 ? 4 : mm[i].N - c >= 2 ? 2 : 1;
                if (nc > nc_cap) nc = nc_cap;
                passes.push_back({i, (int)c, nc}); c += nc;
            }
        }
        if (tl_phase == 1 || tl_phase == 3) continue;           // staging only
        std::vector<char> pdone(passes.size(), 0);
        bool first_launch = true;
        for (size_t p0 = 0; p0 < passes.size(); p0++) {
            if (pdone[p0]) continue;
            const ggb_dev_mm &m0 = mm[passes[p0].i];
            static thread_local GemvBatch gb;
            static_cast<GemvHdr &>(gb) = GemvHdr{};
            int rc = gemv_plan(gb, type, K, m0.nb01, passes[p0].nc, m0.W);
            if (rc) return rc;
            gb.n_peers = m0.n_peers;
            gb.fuse_x = fuse_x ? 1 : 0;
            for (int p = 0; p < m0.n_peers; p++) gb.peer_delta[p] = (long long)(reinterpret_cast<char *>(m0.Y_peer[p]) - reinterpret_cast<char *>(m0.Y));
            auto compatible = [&](const Pass &ps) {
                const ggb_dev_mm &m = mm[ps.i];
                if (ps.nc != passes[p0].nc || m.nb01 != m0.nb01 || m.n_peers != m0.n_peers) return false;
                if (((reinterpret_cast<uintptr_t>(m.W) & 15) == 0) != ((reinterpret_cast<uintptr_t>(m0.W) & 15) == 0)) return false;   // same staging mode
                for (int p = 0; p < m.n_peers; p++)
                    if ((long long)(reinterpret_cast<char *>(m.Y_peer[p]) - reinterpret_cast<char *>(m.Y)) != gb.peer_delta[p]) return false;
                return true;
            };
            auto flush = [&]() -> int {
                if (!gb.n_nodes) return GGB_OK;
                int r;
                { KernelTimer kt(s); r = launch_gemv_batch(gb, s, first_launch || g_timing == 2); }
                first_launch = false;
                gb.n_nodes = 0; gb.total_groups = 0;
                return r;
            };
            for (size_t p1 = p0; p1 < passes.size(); p1++) {
                if (pdone[p1] || !compatible(passes[p1])) continue;
                pdone[p1] = 1;
                const ggb_dev_mm &m = mm[passes[p1].i];
                if (m.flags & GGB_MM_W_IN_FLIGHT) gb.wait_w = 1;
                GemvNode &nd = gb.node[gb.n_nodes++];
                nd.W = static_cast<const uint8_t *>(m.W);
                nd.xq = fuse_x ? reinterpret_cast<const uint8_t *>(m.X) : act_base[passes[p1].i] + (size_t)passes[p1].col0 * arow;
                nd.y = m.Y + (size_t)passes[p1].col0 * (m.ldy_bytes / 4);
                nd.M = (int)m.M; nd.ldy = (int)(m.ldy_bytes / 4);
                const int64_t rows_per_group = gemv_group_rows(gb);
                nd.g0 = gb.total_groups; nd.ngroups = (int)((m.M + rows_per_group - 1) / rows_per_group);
                gb.total_groups += nd.ngroups;
                if (gb.n_nodes == GGB_MAX_BATCH_NODES) { rc = flush(); if (rc) return rc; }
            }
            rc = flush();
            if (rc) return rc;
        }
    }
    return GGB_OK;
}

// dst[n][m] = sum over segments of part[seg][n][m], in segment order (and into the peers' copies of a row split)
struct SegPeers { int n; long long delta[7]; };
__global__ void __launch_bounds__(256) k_sum_segments(const float *__restrict__ part, int nseg, long long M, long long N, float *__restrict__ y, long long ldy,
                                                      const __grid_constant__ SegPeers peers)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the segment GEMVs in front of this kernel
    const long long total = M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v = part[i];
        for (int sgi = 1; sgi < nseg; sgi++) v += part[(long long)sgi * total + i];
        const long long n = i / M, m = i - n * M;
        float *yp = y + n * ldy + m;
        *yp = v;
        for (int p = 0; p < peers.n; p++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + peers.delta[p]) = v;
    }
}

// One batch of independent mul_mats.  Nodes whose rows are too long for the GEMV's shared-memory resident activation row are split
// into K segments here (see needs_k_segments); everything else goes straight to dev_batch_inner.
static int dev_batch(const ggb_dev_mm *mm, int count, void *ws, size_t ws_bytes, cudaStream_t s)
{
    bool any = false;
    for (int i = 0; i < count; i++) if (needs_k_segments(mm[i])) any = true;
    if (!any) return dev_batch_inner(mm, count, ws, ws_bytes, s);
    std::vector<ggb_dev_mm> sub;
    struct Long { int node; size_t part_off; int nseg; };
    std::vector<Long> longs;
    size_t inner_bytes = 0;
    for (int i = 0; i < count; i++) {
        int rc = check_mm(mm[i]);
        if (rc) return rc;
        if (!needs_k_segments(mm[i])) { sub.push_back(mm[i]); inner_bytes += mm_ws_bytes_plain(mm[i]); continue; }
        const ggb_dev_mm &m = mm[i];
        if (m.K % GGB_QK && is_q_weight(m.type)) return set_error(GGB_E_INVALID, "mul_mat: ne00=%lld %% 32 != 0 (Ggml.cs:6694)", (long long)m.K);
        longs.push_back({i, 0, (int)k_segments(m.K)});
        const size_t esz = type_size(m.type);
        const int blck = blck_size(m.type);
        for (int64_t k0 = 0; k0 < m.K; k0 += GGB_KSEG) {
            ggb_dev_mm sg = m;
            sg.K = std::min(GGB_KSEG, m.K - k0);
            sg.W = static_cast<const uint8_t *>(m.W) + (size_t)(k0 / blck) * esz;      // same rows, columns k0 .. k0 + K
            sg.X = m.X + k0;
            sg.n_peers = 0;                                   // partial sums stay local; the reduction broadcasts
            sg.W_rowexp = nullptr;
            sg.Y = nullptr;                                   // filled in below, once the layout is known
            sg.ldy_bytes = m.M * 4;
            sub.push_back(sg);
            inner_bytes += mm_ws_bytes_plain(sg);
        }
    }
    size_t total = align_up(inner_bytes, 256);
    for (Long &l : longs) { l.part_off = total; total += align_up((size_t)l.nseg * (size_t)mm[l.node].N * (size_t)mm[l.node].M * 4, 256); }
    if (total > ws_bytes) return set_error(GGB_E_INVALID, "mul_mat: workspace %zu B < required %zu B", ws_bytes, total);
    uint8_t *wsb = static_cast<uint8_t *>(ws);
    {   // point every segment at its slab of partial sums
        size_t j = 0, li = 0;
        for (int i = 0; i < count; i++) {
            if (!needs_k_segments(mm[i])) { j++; continue; }
            const Long &l = longs[li++];
            for (int sgi = 0; sgi < l.nseg; sgi++, j++)
                sub[j].Y = reinterpret_cast<float *>(wsb + l.part_off) + (size_t)sgi * (size_t)mm[i].N * (size_t)mm[i].M;
        }
    }
    int rc = dev_batch_inner(sub.data(), (int)sub.size(), ws, inner_bytes, s);
    if (rc) return rc;
    for (const Long &l : longs) {
        const ggb_dev_mm &m = mm[l.node];
        SegPeers peers = {};
        peers.n = m.n_peers;
        for (int p = 0; p < m.n_peers; p++) peers.delta[p] = (long long)(reinterpret_cast<char *>(m.Y_peer[p]) - reinterpret_cast<char *>(m.Y));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)std::min<long long>((m.M * m.N + 255) / 256, (long long)device_sm_count() * 8));
        cfg.blockDim = dim3(256); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        GGB_CUDA(cudaLaunchKernelEx(&cfg, k_sum_segments, reinterpret_cast<const float *>(wsb + l.part_off), l.nseg, (long long)m.M, (long long)m.N, m.Y,
                                    (long long)(m.ldy_bytes / 4), peers));
        count_launch();
    }
    return GGB_OK;
}

// ------------------------------------------------------------------------------------------------
// pool + executor
// ------------------------------------------------------------------------------------------------

// Small results go back to the pinned host arena through ONE kernel (coalesced 16-byte stores over PCIe to
// UVA-mapped memory) instead of one cudaMemcpyAsync per node; small activations are read by the activation
// kernel straight from the arena the same way.  Large tensors still use the copy engine.
constexpr size_t ZC_MAX = 64 * 1024;
struct CopySeg { const uint8_t *src; uint8_t *dst; unsigned bytes; unsigned vec0; };     // vec0: first 16-byte vector index of this segment
struct CopyBatch { int n; unsigned total_vec; CopySeg seg[96]; };
__global__ void __launch_bounds__(256) k_copy_batch(const __grid_constant__ CopyBatch b)
{
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < b.total_vec; v += gridDim.x * blockDim.x) {
        int sgi = 0;
        while (sgi + 1 < b.n && v >= b.seg[sgi + 1].vec0) sgi++;
        const CopySeg &sg = b.seg[sgi];
        const unsigned off = (v - sg.vec0) * 16;
        if (off + 16 <= sg.bytes) *reinterpret_cast<uint4 *>(sg.dst + off) = *reinterpret_cast<const uint4 *>(sg.src + off);
        else for (unsigned i = off; i < sg.bytes; i += 4) *reinterpret_cast<uint32_t *>(sg.dst + i) = *reinterpret_cast<const uint32_t *>(sg.src + i);
    }
}
static int flush_copy_batch(CopyBatch &cb, cudaStream_t s)
{
    if (!cb.n) return GGB_OK;
    // same shared-memory carve-out as the GEMV it is meant to run beside: an SM is not shared by kernels of different carve-outs
    static PerDeviceOnce once;
    if (once.need()) { cudaFuncSetAttribute(k_copy_batch, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); cudaGetLastError(); }
    k_copy_batch<<<std::min(64u, (cb.total_vec + 255) / 256), 256, 0, s>>>(cb);
    count_launch();
    GGB_CUDA(cudaGetLastError());
    cb.n = 0; cb.total_vec = 0;
    return GGB_OK;
}

// Device copy of a leaf src0.  rowexp: the power-of-two row exponents of the tensor-core path (launch_weight_rowexp), computed the
// first time a view (byte offset, type, M, K, nb01) of the mirror meets a batched node and kept for as long as the mirror lives.
struct RowExpKey { size_t off; int type; int64_t M, K, nb01; bool operator<(const RowExpKey &o) const { return std::tie(off, type, M, K, nb01) < std::tie(o.off, o.type, o.M, o.K, o.nb01); } };
struct Mirror {
    void *dptr = nullptr; size_t bytes = 0;      // bytes: the HOST byte range [data, data + bytes) the mirror stands for
    int64_t row0 = 0, rows = 0;                  // which rows of the weight matrix this device holds (row split: a slice per GPU)
    std::map<RowExpKey, int *> rowexp;
};
// Bumped whenever device memory that an instantiated CUDA graph may point into goes away (a weight mirror is freed, an arena is
// reallocated): cached graphs of an older epoch are never launched again (ggb_graph_compute_mul_mats).
static uint64_t g_epoch = 1;
static void free_mirror(Mirror &m)
{
    g_epoch++;
    for (auto &kv : m.rowexp) cudaFree(kv.second);
    m.rowexp.clear();
    cudaFree(m.dptr);
    m.dptr = nullptr;
}

struct DevArena {           // grow-only device scratch, reset per compute
    uint8_t *base = nullptr; size_t cap = 0, used = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return GGB_OK;
        g_epoch++;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        const size_t want = align_up(bytes + bytes / 4, 1 << 20);
        GGB_CUDA(cudaMalloc(reinterpret_cast<void **>(&base), want));
        cap = want;
        return GGB_OK;
    }
    void *take(size_t bytes) { void *p = base + used; used += align_up(bytes, 256); return p; }
};

// per-device state of a pool.  mirrors: keyed by the host data pointer of a leaf src0
struct PoolDev { DevArena arena, sym; std::map<const void *, Mirror> mirrors; };

// One ggml_cgraph as the executor last saw it, and -- from the second identical compute on -- the CUDA graph of everything the
// executor enqueues for it (uploads from the host arena, staging, mul_mats, neighbours, result copies; both lanes).  The reference
// re-plans and re-walks the node list on every ggml_graph_compute (Ggml.cs:3260-3704); a decode loop computes the SAME graph over and
// over, so the replay turns ~25 us (32 nodes) to ~1 ms (224 nodes) of host work per compute into one cudaGraphLaunch.  The key covers
// every field the enqueue depends on (see graph_key); the bytes the graph reads and writes are the tensors' own, so results track
// whatever the user writes into tensor->data between computes exactly as the eager path does.
struct GraphEntry {
    uint64_t key = 0, epoch = 0; int seen = 0; bool bad = false;
    cudaGraphExec_t exec = nullptr;
    std::vector<uint8_t> done; int n_run = 0;
    std::vector<ggml_tensor *> run;
    uint64_t d_launches = 0, d_h2d = 0, d_d2h = 0, d_hits = 0, d_uploads = 0;      // what one replay adds to ggb_stats
    std::vector<float> share;                 // each node's share of the compute's device time, as measured level by level when it last ran eagerly
};

} // namespace ggb

struct ggb_pool {
    void *host_base = nullptr;
    size_t bytes = 0;
    bool owned = false, registered = false, register_tried = false;
    // Weight residency is OPT-IN (ggb_pool_set_weight_cache / GGB200_WEIGHT_CACHE=1): the reference re-reads src0->data on every
    // ggml_graph_compute, so by default every leaf src0 is uploaded again; a resident pool promises that leaf weights are rewritten
    // only through the API (CPY / in-place nodes, ggml_set_*), or followed by ggb_tensor_invalidate.
    bool weight_cache = false;
    // Row split across the GPUs of the box (ggb_pool_set_row_split / GGB200_ROW_SPLIT): 0 = off, n = up to n devices.  A MUL_MAT node is
    // split only if its src0 has at least shard_min_bytes (north_star: "only for matrices large enough to benefit").
    int shard_devices = 0;
    size_t shard_min_bytes = 4u << 20;
    std::vector<ggb::PoolDev> pd{1};                // per device: scratch arena, symmetric arena of node outputs, weight mirrors
    std::vector<ggb::GraphEntry> graphs;            // the few cgraphs this pool computes again and again (single-device computes only)
};

namespace ggb {

static size_t tensor_span(const ggml_tensor *t)
{
    // bytes from data to the end of the last element, honouring strides (views, padded rows, transposed / permuted tensors)
    const int blck = blck_size(t->type);
    size_t span = type_size(t->type);
    span += (size_t)(t->ne[0] / blck - 1) * t->nb[0];
    for (int i = 1; i < GGML_MAX_DIMS; i++) span += (size_t)(t->ne[i] - 1) * t->nb[i];
    return span;
}
static bool is_contiguous(const ggml_tensor *t)                  // ggml_is_contiguous, Ggml.cs:3823-3832
{
    const uint64_t ts = type_size(t->type);
    return t->nb[0] == ts && t->nb[1] == ts * (uint64_t)(t->ne[0] / blck_size(t->type)) && t->nb[2] == t->nb[1] * (uint64_t)t->ne[1] &&
           t->nb[3] == t->nb[2] * (uint64_t)t->ne[2];
}
static bool dst_contiguous(const ggml_tensor *t) { return t->type == GGML_TYPE_F32 && is_contiguous(t); }
static bool same_shape(const ggml_tensor *a, const ggml_tensor *b)
{
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}
static int64_t nelements(const ggml_tensor *t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }

// What ggml_compute_forward_mul_mat_* assert (Ggml.cs:6016-6034, 6221-6238, 6388, 6481-6504, 6694) and
// ggml_can_mul_mat (Ggml.cs:8345-8353), returned as a status instead of Debug.Assert.
static int validate_mul_mat(const ggml_tensor *dst)
{
    const ggml_tensor *a = dst->src0, *b = dst->src1;
    if (!a || !b) return set_error(GGB_E_INVALID, "MUL_MAT node without src0/src1");
    if (dst->op != GGML_OP_MUL_MAT) return set_error(GGB_E_INVALID, "node op %d is not GGML_OP_MUL_MAT", dst->op);
    if (b->type != GGML_TYPE_F32 || dst->type != GGML_TYPE_F32) return set_error(GGB_E_INVALID, "mul_mat: src1 and dst must be F32 (Ggml.cs:3379-3382)");
    if (!is_mm_weight(a->type))
        return set_error(GGB_E_UNSUPPORTED, "mul_mat: src0 type %d has no vec_dot in quantize_fns[] (Ggml.cs:219-282)", a->type);
    if (a->ne[0] != b->ne[0] || a->ne[2] != b->ne[2] || a->ne[3] != b->ne[3]) return set_error(GGB_E_INVALID, "mul_mat: ggml_can_mul_mat fails (Ggml.cs:8345-8353)");
    if (dst->ne[0] != a->ne[1] || dst->ne[1] != b->ne[1] || dst->ne[2] != a->ne[2] || dst->ne[3] != a->ne[3])
        return set_error(GGB_E_INVALID, "mul_mat: dst shape does not match (Ggml.cs:6031-6034)");
    if (a->nb[0] != type_size(a->type)) return set_error(GGB_E_INVALID, "mul_mat: permuted src0 (nb00 != type size, Ggml.cs:6022/6227/6492)");
    if (b->nb[0] != 4) return set_error(GGB_E_INVALID, "mul_mat: permuted src1 (nb10 != 4, Ggml.cs:6023/6388/6493)");
    if (dst->nb[0] != 4 || dst->nb[0] > dst->nb[1] || dst->nb[1] > dst->nb[2] || dst->nb[2] > dst->nb[3])
        return set_error(GGB_E_INVALID, "mul_mat: transposed or permuted dst (Ggml.cs:6026-6029)");
    if (a->ne[0] % blck_size(a->type) || (is_q_weight(a->type) && a->ne[0] % GGB_QK)) return set_error(GGB_E_INVALID, "mul_mat: ne00 %% 32 != 0 (Ggml.cs:6694, 1209)");
    if (!dst_contiguous(dst)) return set_error(GGB_E_UNSUPPORTED, "mul_mat: dst must be contiguous (as ggml_mul_mat always creates it, Ggml.cs:8237-8238)");
    if (!a->data || !b->data || !dst->data) return set_error(GGB_E_INVALID, "mul_mat: tensor without data (no_alloc context?)");
    return GGB_OK;
}

static int validate_cpy(const ggml_tensor *node)
{
    const ggml_tensor *a = node->src0, *b = node->src1;
    if (!a || !b || !a->data || !b->data) return set_error(GGB_E_INVALID, "CPY node without operands/data");
    if (a->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "cpy: only F32 sources are on this path");
    if (!is_q_weight(b->type) && b->type != GGML_TYPE_F16)
        return set_error(GGB_E_UNSUPPORTED, "cpy: destination type %d is not on this path (F16 and the quantized weight types)", b->type);
    if (nelements(a) != nelements(b)) return set_error(GGB_E_INVALID, "cpy: element counts differ (Ggml.cs:8281)");
    if (!is_contiguous(a)) return set_error(GGB_E_UNSUPPORTED, "cpy: non-contiguous source");
    if (!is_contiguous(b)) return set_error(GGB_E_UNSUPPORTED, "cpy: non-contiguous destination");
    if (a->ne[0] % blck_size(b->type) || (is_q_weight(b->type) && a->ne[0] % GGB_QK) || a->ne[0] != b->ne[0]) return set_error(GGB_E_UNSUPPORTED, "cpy: rows must map one to one (ne00 == ne0, %% 32)");
    return GGB_OK;
}

// The neighbours of mul_mat.  GGB_E_INVALID = the reference would assert; GGB_E_UNSUPPORTED = valid ggml that stays on the CPU loop.
static int validate_neighbour(const ggml_tensor *t)
{
    const ggml_tensor *a = t->src0, *b = t->src1;
    if (!a || !a->data || !t->data) return set_error(GGB_E_INVALID, "op %d: node without src0/data", t->op);
    switch (t->op) {
    case GGML_OP_ADD:
        if (!b || !b->data) return set_error(GGB_E_INVALID, "add: no src1");
        if (!same_shape(a, b) || !same_shape(a, t)) return set_error(GGB_E_INVALID, "add: shapes differ (Ggml.cs:4628, 4803)");
        if (is_q_weight(a->type)) {                                                    // add_q_f32
            if (b->type != GGML_TYPE_F32 || t->type != a->type) return set_error(GGB_E_INVALID, "add_q_f32: src1 must be F32 and dst of src0's type (Ggml.cs:4862-4864)");
            if (a->ne[0] % GGB_QK) return set_error(GGB_E_INVALID, "add_q_f32: ne00 %% 32 != 0 (Ggml.cs:4891)");
            if (!is_contiguous(a) || !is_contiguous(b) || !is_contiguous(t)) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: contiguous tensors only");
            if (reinterpret_cast<uintptr_t>(b->data) & 15) return set_error(GGB_E_UNSUPPORTED, "add_q_f32: unaligned src1");
            return GGB_OK;
        }
        // fallthrough: F32 + F32
    case GGML_OP_MUL:
        if (!b || !b->data) return set_error(GGB_E_INVALID, "op %d: no src1", t->op);
        if (!same_shape(a, b) || !same_shape(a, t)) return set_error(GGB_E_INVALID, "op %d: shapes differ (Ggml.cs:4628, 5014)", t->op);
        if (a->type != GGML_TYPE_F32 || b->type != GGML_TYPE_F32 || t->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "op %d: F32 operands only on this path", t->op);
        if (!is_contiguous(a) || !is_contiguous(b) || !is_contiguous(t)) return set_error(GGB_E_UNSUPPORTED, "op %d: contiguous tensors only", t->op);
        return GGB_OK;
    case GGML_OP_SILU: case GGML_OP_RMS_NORM:
        if (!same_shape(a, t)) return set_error(GGB_E_INVALID, "op %d: shapes differ (Ggml.cs:5712, 5863)", t->op);
        if (a->type != GGML_TYPE_F32 || t->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "op %d: F32 only (Ggml.cs:5757, 5927)", t->op);
        if (!is_contiguous(a) || !is_contiguous(t)) return set_error(GGB_E_UNSUPPORTED, "op %d: contiguous tensors only", t->op);
        return GGB_OK;
    case GGML_OP_SCALE:
        if (!b || !b->data) return set_error(GGB_E_INVALID, "scale: no src1");
        if (nelements(b) != 1) return set_error(GGB_E_INVALID, "scale: src1 is not a scalar (Ggml.cs:6755)");
        if (!same_shape(a, t)) return set_error(GGB_E_INVALID, "scale: shapes differ (Ggml.cs:6754)");
        if (a->type != GGML_TYPE_F32 || b->type != GGML_TYPE_F32 || t->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "scale: F32 only");
        if (!is_contiguous(a) || !is_contiguous(t)) return set_error(GGB_E_INVALID, "scale: non-contiguous (Ggml.cs:6752-6753)");
        if (b->op != GGML_OP_NONE) return set_error(GGB_E_UNSUPPORTED, "scale: the factor must be a leaf (it is read on the host when the node is enqueued)");
        return GGB_OK;
    case GGML_OP_REPEAT:
        if (a->type != GGML_TYPE_F32 || t->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "repeat: F32 only (Ggml.cs:5392)");
        if (a->ne[2] != 1 || a->ne[3] != 1 || t->ne[2] != 1 || t->ne[3] != 1) return set_error(GGB_E_INVALID, "repeat: rank > 2 (Ggml.cs:5354-5357)");
        if (a->ne[0] <= 0 || a->ne[1] <= 0 || t->ne[0] % a->ne[0] || t->ne[1] % a->ne[1]) return set_error(GGB_E_INVALID, "repeat: ggml_can_repeat fails (Ggml.cs:8398-8407)");
        if (a->nb[0] != 4 || t->nb[0] != 4 || (a->nb[1] & 3) || (t->nb[1] & 3)) return set_error(GGB_E_INVALID, "repeat: transposed operand (Ggml.cs:5367-5368)");
        return GGB_OK;
    case GGML_OP_CONT: case GGML_OP_DUP:
        if (a->type != GGML_TYPE_F32 || t->type != GGML_TYPE_F32) return set_error(GGB_E_UNSUPPORTED, "cont/dup: F32 -> F32 only on this path");
        if (nelements(a) != nelements(t)) return set_error(GGB_E_INVALID, "cont/dup: element counts differ (Ggml.cs:4206)");
        if (!same_shape(a, t) || !is_contiguous(t)) return set_error(GGB_E_UNSUPPORTED, "cont/dup: destination must be the contiguous tensor of src0's shape");
        for (int i = 0; i < GGML_MAX_DIMS; i++) if (a->nb[i] & 3) return set_error(GGB_E_UNSUPPORTED, "cont/dup: unaligned strides");
        return GGB_OK;
    default:
        return set_error(GGB_E_UNSUPPORTED, "op %d is not on this path", t->op);
    }
}
static bool is_view_op(int op) { return op == GGML_OP_RESHAPE || op == GGML_OP_VIEW || op == GGML_OP_PERMUTE || op == GGML_OP_TRANSPOSE; }
static bool is_neighbour_op(int op)
{
    return op == GGML_OP_ADD || op == GGML_OP_MUL || op == GGML_OP_SILU || op == GGML_OP_RMS_NORM || op == GGML_OP_SCALE || op == GGML_OP_REPEAT ||
           op == GGML_OP_CONT || op == GGML_OP_DUP;
}

struct Produced { const uint8_t *host; size_t bytes; uint8_t *dev; int level; };

static inline bool ranges_overlap(const uint8_t *a, size_t an, const uint8_t *b, size_t bn) { return a < b + bn && b < a + an; }

// The newest earlier result whose host byte range INTERSECTS [p, p + span).  *partial is set when the operand is not wholly inside
// it (a view that starts before the result, or runs past its end): such an operand is neither the device copy nor the host bytes,
// and ggb_graph_compute_mul_mats leaves the node to the caller's loop instead.
static Produced *find_produced(std::vector<Produced> &v, const void *p, size_t span, bool *partial = nullptr)
{
    const uint8_t *q = static_cast<const uint8_t *>(p);
    if (partial) *partial = false;
    for (size_t i = v.size(); i-- > 0;) {                                                                   // newest first
        if (!ranges_overlap(q, span, v[i].host, v[i].bytes)) continue;
        const bool inside = q >= v[i].host && q + span <= v[i].host + v[i].bytes;
        if (!inside) { if (partial) *partial = true; return nullptr; }
        // wholly inside the newest overlapping result: every byte of the operand is that result's (older overlapping results were overwritten there)
        return &v[i];
    }
    return nullptr;
}

// Rows [row0, row0 + rows) of an M-row weight matrix that shard g of G owns: the reference's thread split dr = ceil(nr / nth),
// thread ith takes [dr * ith, min(dr * ith + dr, nr)) (Ggml.cs:6665-6672), with nth = G and "thread" = GPU.
static inline void shard_rows(int64_t M, int g, int G, int64_t &row0, int64_t &rows)
{
    const int64_t dr = (M + G - 1) / G;
    row0 = std::min<int64_t>(dr * g, M);
    rows = std::min<int64_t>(row0 + dr, M) - row0;
}

// Runs `nodes` (already validated to be runnable, in graph order; view ops are not in the list).
//
// sc == nullptr: everything on the library's device.  Otherwise this is host thread sc->g of sc->G, each driving one GPU of the box
// (run_nodes_sharded below): every device executes the same node list on its own copy of every intermediate tensor -- element-wise
// neighbours and CPYs are replicated, they cost nothing next to a mul_mat -- and a MUL_MAT node is row-split: device g multiplies
// its rows of src0 and its kernel stores the results straight into EVERY device's copy of dst over NVLink (ggb_dev_mm.Y_peer; all
// devices allocate node outputs from a "symmetric" arena in the same order, so a peer's address is its arena base plus the local
// offset).  That is the all-gather of north_star, fused into the producing kernel.  After a level with MUL_MATs the devices wait for
// each other's level events (cudaStreamWaitEvent across devices: no spinning kernels).  Nodes too small to split (ggb_pool.
// shard_min_bytes) are multiplied by device 0 alone, which still broadcasts the result.  Device 0 returns the results to the host.
constexpr int GGB_E_NOCAPTURE = -100;       // internal: this compute cannot be recorded as a CUDA graph (run it eagerly)

// cap != nullptr (single device only): do not execute -- RECORD everything the compute enqueues into cap->exec (stream capture
// over both lanes).  Anything a replay could not repeat faithfully -- creating or dropping a resident weight mirror, growing an
// arena -- returns GGB_E_NOCAPTURE before it has any effect, and the caller runs the compute eagerly instead.
static int run_nodes(ggb_pool *pool, const std::vector<ggml_tensor *> &nodes, int flags, const std::vector<char> &is_output, ShardCtx *sc, GraphEntry *cap = nullptr,
                     std::vector<float> *share_out = nullptr)
{
    const int g = sc ? sc->g : 0, G = sc ? sc->G : 1;
    int rc = GGB_OK;
    if (!sc) { rc = ensure_init(); if (rc) return rc; }
    else GGB_CUDA(cudaSetDevice(g_devs[(size_t)g].dev));
    DevCtx &dctx = g_devs[(size_t)g];
    cudaStream_t s = dctx.stream;
    cudaEvent_t ev0 = dctx.ev0, ev1 = dctx.ev1;
    const bool lead = g == 0;                                // device 0 keeps the statistics and talks to the host arena
    const size_t n = nodes.size();
    if (!n) return GGB_OK;
    PoolDev &pd = pool->pd[(size_t)g];
    DevArena &arena = pd.arena, &sym = pd.sym;
    std::map<const void *, Mirror> &mirrors = pd.mirrors;
    auto out_tensor = [](ggml_tensor *t) -> ggml_tensor * { return t->op == GGML_OP_CPY ? t->src1 : t; };      // whose data the node writes

    const bool cache_on = pool->weight_cache && !(flags & GGB_GRAPH_NO_WEIGHT_CACHE);
    // a leaf src0 may stay resident only if nothing rewrites it behind the API's back: never a parameter (ggml_opt updates
    // params in place between computes, Ggml.cs:1734-1760) and never a tensor with a gradient
    auto cacheable = [&](const ggml_tensor *x) { return cache_on && x->op == GGML_OP_NONE && !x->is_param && !x->grad; };
    // which rows of a MUL_MAT node's src0 this device multiplies
    auto my_rows = [&](const ggml_tensor *a, int64_t &row0, int64_t &rows) {
        row0 = 0; rows = a->ne[1];
        if (G == 1) return;
        const bool split = a->ne[2] * a->ne[3] == 1 && (size_t)a->ne[1] * a->nb[1] >= pool->shard_min_bytes && a->ne[1] >= 64 * (int64_t)G;
        if (split) shard_rows(a->ne[1], g, G, row0, rows);
        else if (g != 0) rows = 0;                            // too small to split: device 0 multiplies all of it (and still broadcasts)
    };

    // ---- pass 1: an upper bound of the scratch needed, so both arenas are allocated once before anything is enqueued.
    //      `sym` holds node outputs only, taken in node order: the same offsets on every device of a row split ----
    size_t need = 0, need_sym = 0;
    {
        std::vector<Produced> plan;
        for (size_t i = 0; i < n; i++) {
            ggml_tensor *t = nodes[i];
            const ggml_tensor *a = t->src0, *b = t->src1;
            if (t->op == GGML_OP_MUL_MAT) {
                Produced *a_pr = find_produced(plan, a->data, tensor_span(a));
                const bool a_dev = a_pr != nullptr;
                if (!a_dev && !cacheable(a)) need += align_up(tensor_span(a), 256);
                if (!find_produced(plan, b->data, tensor_span(b))) need += align_up(tensor_span(b), 256);
                // staged weights start 256-byte aligned (mirror / arena) unless they are a view into an earlier node's output
                const bool may_expand = (a->nb[1] & 15) || (a->nb[2] & 15) || (a->nb[3] & 15) ||
                                        (a_pr && ((static_cast<const uint8_t *>(a->data) - a_pr->host) & 15));
                need += (size_t)(a->ne[2] * a->ne[3]) * mm_ws_bytes_bound(a->type, a->ne[1], a->ne[0], b->ne[1], may_expand);
            } else {
                if (!find_produced(plan, a->data, tensor_span(a))) need_sym += align_up(tensor_span(a), 256);      // (an in-place node on a leaf works on a sym copy)
                if (b && t->op != GGML_OP_CPY && t->op != GGML_OP_SCALE && !find_produced(plan, b->data, tensor_span(b))) need += align_up(tensor_span(b), 256);
                if (!find_produced(plan, a->data, tensor_span(a))) need += align_up(tensor_span(a), 256);
            }
            ggml_tensor *o = out_tensor(t);
            need_sym += align_up(tensor_span(o), 256);
            plan.push_back({static_cast<const uint8_t *>(o->data), tensor_span(o), nullptr, 0});
        }
    }
    if (cap && (need + 8192 > arena.cap || need_sym + 4096 > sym.cap)) return GGB_E_NOCAPTURE;
    rc = arena.reserve(need + 8192);
    if (!rc) rc = sym.reserve(need_sym + 4096);
    arena.used = 0; sym.used = 0;
    if (sc) {
        // every device publishes where its symmetric arena lives; nobody enqueues a peer store before all arenas exist
        SHARD_TRACE("dev %d/%d: arenas reserved (%zu + %zu B) rc=%d", g, G, need, need_sym, rc);
        sc->sh->sym_base[g] = sym.base;
        if (rc) sc->sh->bar.fail();
        if (!sc->sh->bar.arrive_and_wait()) return rc ? rc : set_error(GGB_E_CUDA, "row split: another device failed to set up");
    }
    if (rc) return rc;

    // RAII: a capture that is open when this function returns early is closed and thrown away
    struct CaptureGuard {
        cudaStream_t s; bool open = false;
        ~CaptureGuard() { if (open) { cudaGraph_t g = nullptr; cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); } }
    } capture{s};
    const ggb_stats stats0 = g_stats;
    if (cap) {
        GGB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        capture.open = true;
    } else GGB_CUDA(cudaEventRecord(ev0, s));

    // ---- pass 2: stage inputs, group nodes into dependency levels, launch ----
    std::vector<Produced> produced;
    // ew: row exponents per (i2, i3) slice of a resident src0; row0 / rows: the rows of src0 this device multiplies (da points at row0)
    struct Item { ggml_tensor *t; int level; uint8_t *da, *db, *dd; bool a_in_flight, ew_new; int64_t row0, rows; std::vector<const int *> ew; };
    std::vector<Item> items(n);
    int max_level = 0;
    // Mirrors made stale by a node of this call may still be read by EARLIER nodes of it that are not launched yet: they are only
    // unlinked here and freed when the call is over (the stream is synchronised first, also on an error return).
    struct Graveyard {
        std::vector<Mirror> dead; cudaStream_t s;
        ~Graveyard() { if (dead.empty()) return; cudaStreamSynchronize(s); for (Mirror &m : dead) free_mirror(m); cudaGetLastError(); }
    } graveyard{{}, s};
    // a write to [host, host + bytes) makes every cached mirror that shares a byte with it stale (a CPY into an offset view of a
    // cached leaf, an in-place node on it): dropped by byte range, not by start pointer
    bool capture_refused = false;
    auto drop_mirrors = [&](const void *hostp, size_t bytes) {
        const uint8_t *h = static_cast<const uint8_t *>(hostp);
        for (auto f = mirrors.begin(); f != mirrors.end();) {
            if (cap && ranges_overlap(h, bytes, static_cast<const uint8_t *>(f->first), f->second.bytes)) { capture_refused = true; return; }
            if (ranges_overlap(h, bytes, static_cast<const uint8_t *>(f->first), f->second.bytes)) { graveyard.dead.push_back(std::move(f->second)); f = mirrors.erase(f); }
            else ++f;
        }
    };
    // device address of an operand: produced earlier in this call, a cached weight mirror, read in place from the pinned
    // arena (small tensors, UVA), or uploaded into the scratch arena.  A weight is staged from row `row0` on, `rows` rows only.
    auto stage = [&](const ggml_tensor *x, int &level, bool weight, int64_t row0, int64_t rows, uint8_t *&out, Mirror **mir, bool *in_flight) -> int {
        const size_t span = tensor_span(x);
        bool partial = false;
        if (Produced *pr = find_produced(produced, x->data, span, &partial)) {
            out = pr->dev + (static_cast<const uint8_t *>(x->data) - pr->host) + (weight ? (size_t)row0 * x->nb[1] : 0);
            level = std::max(level, pr->level + 1);
            if (in_flight) *in_flight = true;
            return GGB_OK;
        }
        if (partial) return set_error(GGB_E_UNSUPPORTED, "executor: an operand overlaps an earlier node's result without lying inside it");
        // the bytes this device needs: the whole tensor, or (row split) its rows of a 2-D weight
        const bool part = weight && rows != x->ne[1];
        const uint8_t *src = static_cast<const uint8_t *>(x->data) + (part ? (size_t)row0 * x->nb[1] : 0);
        const size_t bytes = part ? (rows > 0 ? (size_t)(rows - 1) * x->nb[1] + (size_t)(x->ne[0] / blck_size(x->type)) * type_size(x->type) : 0) : span;
        if (weight && cacheable(x)) {
            auto f = mirrors.find(x->data);
            if (f != mirrors.end() && (f->second.bytes < span || f->second.row0 != row0 || f->second.rows != rows)) {
                if (cap) return GGB_E_NOCAPTURE;
                graveyard.dead.push_back(std::move(f->second)); mirrors.erase(f); f = mirrors.end();      // another shape or another split: stale
            }
            if (f == mirrors.end()) {
                if (cap) return GGB_E_NOCAPTURE;              // a replay must not upload the mirror again
                void *d = nullptr;
                GGB_CUDA(cudaMalloc(&d, align_up(std::max<size_t>(bytes, 1), 256)));
                if (bytes) GGB_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, s));
                if (lead) { g_stats.h2d_bytes += bytes; g_stats.weight_uploads++; }
                Mirror m; m.dptr = d; m.bytes = span; m.row0 = row0; m.rows = rows;      // `bytes` = the HOST range the mirror stands for (invalidation)
                f = mirrors.emplace(x->data, std::move(m)).first;
            } else if (lead) g_stats.weight_cache_hits++;
            out = static_cast<uint8_t *>(f->second.dptr);
            if (mir) *mir = &f->second;
            return GGB_OK;
        }
        if (!weight && pool->owned && span <= ZC_MAX && (reinterpret_cast<uintptr_t>(x->data) & 3) == 0) {
            out = static_cast<uint8_t *>(x->data);               // UVA: kernels read the pinned arena directly (portable pinned memory: from any device)
        } else {
            out = static_cast<uint8_t *>(arena.take(std::max<size_t>(bytes, 1)));
            if (bytes) GGB_CUDA(cudaMemcpyAsync(out, src, bytes, cudaMemcpyHostToDevice, s));
        }
        if (lead) g_stats.h2d_bytes += bytes;
        return GGB_OK;
    };
    for (size_t i = 0; i < n; i++) {
        ggml_tensor *t = nodes[i];
        ggml_tensor *a = t->src0, *b = t->src1;
        Item &it = items[i];
        it.t = t; it.level = 0; it.da = it.db = it.dd = nullptr; it.a_in_flight = false; it.ew_new = false;
        it.row0 = 0; it.rows = a->ne[1];
        const bool is_mm = t->op == GGML_OP_MUL_MAT;
        if (is_mm) my_rows(a, it.row0, it.rows);
        Mirror *mir = nullptr;
        ggml_tensor *o = out_tensor(t);
        const size_t ospan = tensor_span(o);
        // does the node write over one of its inputs (ggml_scale, ggml_*_inplace: the result is a view of src0)?
        const uint8_t *od = static_cast<const uint8_t *>(o->data), *ad = static_cast<const uint8_t *>(a->data);
        const bool in_place = t->op != GGML_OP_CPY && od >= ad && od < ad + tensor_span(a);
        if (in_place && t->op != GGML_OP_SCALE && t->op != GGML_OP_ADD && t->op != GGML_OP_MUL && t->op != GGML_OP_SILU && t->op != GGML_OP_RMS_NORM)
            return set_error(GGB_E_UNSUPPORTED, "op %d writing over its own source is not supported", t->op);
        if (in_place && od != ad) return set_error(GGB_E_UNSUPPORTED, "in-place op %d on a shifted view", t->op);
        bool a_partial = false;
        Produced *a_prod = in_place ? find_produced(produced, a->data, tensor_span(a), &a_partial) : nullptr;
        if (a_partial) return set_error(GGB_E_UNSUPPORTED, "executor: an in-place operand overlaps an earlier node's result without lying inside it");
        if (in_place && !a_prod) {
            // in place on a leaf: work on a device copy (never on the pinned host arena directly, never on a cached mirror)
            const size_t span = tensor_span(a);
            it.da = static_cast<uint8_t *>(sym.take(span));
            GGB_CUDA(cudaMemcpyAsync(it.da, a->data, span, cudaMemcpyHostToDevice, s));
            if (lead) g_stats.h2d_bytes += span;
            drop_mirrors(a->data, span);
        } else if (!is_mm || it.rows > 0) {
            rc = stage(a, it.level, is_mm, it.row0, it.rows, it.da, &mir, &it.a_in_flight);
            if (rc) return rc;
        } else {
            // a MUL_MAT none of whose rows are this device's: nothing to stage, but the node still sits one level above its producers
            if (Produced *pr = find_produced(produced, a->data, tensor_span(a))) it.level = std::max(it.level, pr->level + 1);
        }
        if (mir && is_mm && b->ne[1] >= 16 && is_q_weight(a->type)) {
            // resident weights: the row exponents of the tensor-core path (launch_weight_rowexp) are computed once per mirror view
            for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) {
                const RowExpKey key{(size_t)(i2 * a->nb[2] + i3 * a->nb[3]), a->type, it.rows, a->ne[0], (int64_t)a->nb[1]};
                auto f = mir->rowexp.find(key);
                if (f == mir->rowexp.end()) {
                    if (cap) return GGB_E_NOCAPTURE;
                    int *ew = nullptr;
                    GGB_CUDA(cudaMalloc(reinterpret_cast<void **>(&ew), align_up((size_t)std::max<int64_t>(it.rows, 1) * 4, 256)));
                    rc = launch_weight_rowexp(a->type, it.da + key.off, (int64_t)a->nb[1], it.rows, a->ne[0], ew, s);
                    if (rc) { cudaFree(ew); return rc; }
                    f = mir->rowexp.emplace(key, ew).first;
                    it.ew_new = true;                            // written on this stream just now: the GEMM's weight side waits
                }
                it.ew.push_back(f->second);
            }
        }
        if (b && t->op != GGML_OP_CPY && t->op != GGML_OP_SCALE && t->op != GGML_OP_REPEAT) {
            if (!is_mm || it.rows > 0) {
                rc = stage(b, it.level, false, 0, 0, it.db, nullptr, nullptr);
                if (rc) return rc;
            } else if (Produced *pr = find_produced(produced, b->data, tensor_span(b))) it.level = std::max(it.level, pr->level + 1);
        }
        if (in_place) {
            // runs after every earlier node (some of them may still read the old contents), and later readers wait for it
            it.level = std::max(it.level, max_level + 1);
            it.dd = it.da;
            if (a_prod) { a_prod->level = it.level; drop_mirrors(od, ospan); }
            else produced.push_back({od, ospan, it.dd, it.level});
        } else {
            it.dd = static_cast<uint8_t *>(sym.take(ospan));
            if (t->op == GGML_OP_CPY) {
                drop_mirrors(od, ospan);                         // a CPY rewrites b->data: every cached mirror sharing bytes with it is stale from here on
                // ... and it must not overtake an earlier node of this call that still reads or writes those bytes on the device
                for (const Produced &pr : produced) if (ranges_overlap(od, ospan, pr.host, pr.bytes)) it.level = std::max(it.level, pr.level + 1);
            }
            produced.push_back({od, ospan, it.dd, it.level});
        }
        max_level = std::max(max_level, it.level);
        if (capture_refused) return GGB_E_NOCAPTURE;             // the node makes a resident mirror stale: not something to replay
    }

    SHARD_TRACE("dev %d/%d: staged %zu nodes, %d levels", g, G, n, max_level + 1);
    // results back into the host arena (device 0 holds every result): small ones through one coalesced copy kernel per call site,
    // large ones through the copy engine.  `copied` marks what has already gone back (the two-lane path below returns a chunk's
    // results as soon as its GEMV is done).
    std::vector<char> copied(n, 0);
    auto copy_out = [&](const std::vector<size_t> &idx, cudaStream_t st) -> int {
        CopyBatch cb = {};
        for (size_t i : idx) {
            const Item &it = items[i];
            ggml_tensor *t = it.t;
            if (copied[i]) continue;
            if ((flags & GGB_GRAPH_KEEP_ON_DEVICE) && !is_output[i] && t->op != GGML_OP_CPY) continue;   // a CPY target is user-visible
            copied[i] = 1;
            ggml_tensor *o = out_tensor(t);
            void *hdst = o->data;
            const size_t span = tensor_span(o);
            if (pool->owned && span <= ZC_MAX && (span & 3) == 0 && (reinterpret_cast<uintptr_t>(hdst) & 15) == 0) {
                CopySeg &sg = cb.seg[cb.n++];
                sg.src = it.dd; sg.dst = static_cast<uint8_t *>(hdst); sg.bytes = (unsigned)span; sg.vec0 = cb.total_vec;
                cb.total_vec += (unsigned)((span + 15) / 16);
                if (cb.n == 96) { int r = flush_copy_batch(cb, st); if (r) return r; }
            } else {
                GGB_CUDA(cudaMemcpyAsync(hdst, it.dd, span, cudaMemcpyDeviceToHost, st));
            }
            g_stats.d2h_bytes += span;
        }
        return flush_copy_batch(cb, st);
    };
    // may node i's result go back to the host before the LATER levels have run?  Not if a later node rewrites the same host bytes
    // in place (the last writer carries the result) -- those are copied in node order at the end.
    auto final_value = [&](size_t i) {
        const uint8_t *od = static_cast<const uint8_t *>(out_tensor(items[i].t)->data);
        const size_t ospan = tensor_span(out_tensor(items[i].t));
        for (size_t j = i + 1; j < n; j++) {
            ggml_tensor *oj = out_tensor(items[j].t);
            if (ranges_overlap(od, ospan, static_cast<const uint8_t *>(oj->data), tensor_span(oj))) return false;
        }
        return true;
    };


    // ---- a dependent chain of single-token nodes runs as ONE persistent launch: the decode program (ggb_internal.h: DpProgram,
    //      ggb_gemv.cu: k_decode_program).  Built from the node list in node order -- the order the reference executes
    //      (Ggml.cs:3539-3704).  All or nothing: a node the program does not take, a dependency level whose mul_mats multiply
    //      different rows (those are one wide batch on the level path), or a graph with a single mul_mat level leaves everything to
    //      the per-level launches below. ----
    bool ran_program = false;
    if (!sc && lead && !g_timing && g_decode_program && n >= 2) {
        std::vector<DpProgram> progs;
        std::vector<size_t> prog_copied;                           // nodes whose results the program itself stores into the host arena
        bool need_silu = false;
        static const bool trace = getenv("GGB200_TRACE_PROGRAM") != nullptr;      // says on stderr why a graph was left to the per-level launches
        size_t at = 0;
        auto bail = [&](const char *why) { if (trace) fprintf(stderr, "ggb200: decode program not used: %s (node %zu of %zu, op %d)\n", why, at, n, at < n ? (int)items[at].t->op : -1); return false; };
        auto build = [&]() -> bool {
            const int64_t vmax = decode_program_row_max();
            // every node must be one the program takes; mul_mats of one level must share their activations
            std::vector<const uint8_t *> level_x((size_t)max_level + 1, nullptr);
            int mm_levels = 0;
            for (size_t i = 0; i < n; i++) {
                const Item &it = items[i];
                const ggml_tensor *t = it.t, *a = t->src0, *b = t->src1;
                at = i;
                switch (t->op) {
                case GGML_OP_MUL_MAT: {
                    if (b->ne[1] != 1 || a->ne[2] * a->ne[3] != 1 || b->ne[2] * b->ne[3] != 1 || it.a_in_flight || it.rows != a->ne[1] || it.rows <= 0) return bail("a mul_mat that is batched, prompt-sized, or whose src0 is an earlier node's result");
                    if (b->type != GGML_TYPE_F32) return bail("src1 is not F32");
                    DpStep probe = {};
                    if (!decode_program_plan_step(probe, a->type, a->ne[0], (int64_t)a->nb[1], it.da)) return bail("src0 type / row length / alignment");
                    if (!level_x[(size_t)it.level]) { level_x[(size_t)it.level] = it.db; mm_levels++; }
                    else if (level_x[(size_t)it.level] != it.db) return bail("mul_mats of one dependency level multiply different rows");
                    break;
                }
                case GGML_OP_ADD: case GGML_OP_MUL:
                    if (a->type != GGML_TYPE_F32 || b->type != GGML_TYPE_F32 || nelements(a) != nelements(t) || nelements(b) != nelements(t)) return bail("a binary op that broadcasts or is not F32");
                    // fallthrough
                case GGML_OP_SILU: case GGML_OP_RMS_NORM: case GGML_OP_SCALE:
                    if (a->type != GGML_TYPE_F32 || t->ne[1] * t->ne[2] * t->ne[3] != 1 || t->ne[0] > vmax || nelements(a) != nelements(t)) return bail("a row op on more than one row (or a very long one)");
                    if (t->ne[0] & 3) return bail("a row that is not whole float4s");
                    break;
                default:
                    return bail("an op the program does not take");
                }
            }
            at = n;
            if (mm_levels < 2) return bail("fewer than two dependent mul_mat levels");
            // small leaves are normally read in place from the pinned host arena (stage() above); here every CTA would read all of
            // them over PCIe, so they get a device copy first
            {
                std::vector<std::pair<const uint8_t *, uint8_t *>> moved;
                auto on_device = [&](const ggml_tensor *x, uint8_t *&p) -> bool {
                    if (!x || p != static_cast<const uint8_t *>(x->data)) return true;
                    for (const auto &m : moved) if (m.first == p) { p = m.second; return true; }
                    const size_t bytes = tensor_span(x);
                    if (arena.used + align_up(std::max<size_t>(bytes, 1), 256) > arena.cap) return false;
                    uint8_t *d = static_cast<uint8_t *>(arena.take(std::max<size_t>(bytes, 1)));
                    if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return false; }
                    moved.push_back({p, d});
                    p = d;
                    return true;
                };
                for (size_t i = 0; i < n; i++) {
                    Item &it = items[i];
                    const ggml_tensor *t = it.t;
                    if (t->op != GGML_OP_MUL_MAT && !on_device(t->src0, it.da)) return bail("no scratch left for a leaf operand");
                    if (t->op != GGML_OP_SCALE && !on_device(t->src1, it.db)) return bail("no scratch left for a leaf operand");
                }
            }
            for (size_t i = 0; i < n; i++) {
                const Item &it = items[i];
                at = i;
                const bool mm = it.t->op == GGML_OP_MUL_MAT, binary = it.t->op == GGML_OP_ADD || it.t->op == GGML_OP_MUL;
                if ((reinterpret_cast<uintptr_t>(it.dd) | (mm ? 0 : reinterpret_cast<uintptr_t>(it.da)) | (mm || binary ? reinterpret_cast<uintptr_t>(it.db) : 0)) & 15)
                    return bail("a row that is not 16-byte aligned on the device");
            }
            struct Range { const uint8_t *p; size_t n; };
            std::vector<Range> step_reads, step_writes;
            std::vector<int> wstep(n, -1);                         // the (global) step a node's result is written in
            int gstep = 0;
            DpProgram cur = {};
            const uint8_t *run_ptr = nullptr; int64_t run_n = 0;   // the tensor the running row holds
            int n_nodes = 0, n_ops = 0;
            // results of closed steps that go back to the host arena from inside the program (the conditions of copy_out's coalesced
            // kernel, in whole float4s); what does not fit a parameter block is left to copy_out
            struct Pending { DpCopy c; size_t item; };
            std::vector<Pending> pending;
            int n_copies = 0;
            auto to_host = [&](size_t i) {
                const Item &it = items[i];
                ggml_tensor *o = out_tensor(it.t);
                const size_t span = tensor_span(o);
                if ((flags & GGB_GRAPH_KEEP_ON_DEVICE) && !is_output[i]) return;
                if (!pool->owned || span > ZC_MAX || (span & 15) || ((reinterpret_cast<uintptr_t>(o->data) | reinterpret_cast<uintptr_t>(it.dd)) & 15) || !final_value(i)) return;
                pending.push_back({DpCopy{reinterpret_cast<const float4 *>(it.dd), static_cast<float4 *>(o->data), (int)(span / 16), 0}, i});
            };
            auto open_step = [&]() {
                DpStep &st = cur.step[cur.n_steps];
                st = DpStep{}; st.op0 = n_ops; st.node0 = n_nodes; st.copy0 = n_copies;
                for (const Pending &pc : pending) if (n_copies < DP_MAX_COPIES) { cur.copy[n_copies++] = pc.c; prog_copied.push_back(pc.item); }
                pending.clear();
                st.ncopies = n_copies - st.copy0;
            };
            auto finish_program = [&]() -> bool {
                if (cur.n_steps == 0) return true;
                if (!decode_program_finish(cur)) return false;
                progs.push_back(cur);
                cur = DpProgram{}; n_nodes = 0; n_ops = 0; n_copies = 0; run_ptr = nullptr; run_n = 0;
                open_step();
                return true;
            };
            auto close_step = [&]() -> bool {                      // ends the open step with a grid barrier (an empty step is not closed)
                DpStep &st = cur.step[cur.n_steps];
                if (st.nops == 0 && st.nnodes == 0 && st.ncopies == 0) return true;
                cur.n_steps++; gstep++;
                step_reads.clear(); step_writes.clear();
                if (cur.n_steps == DP_MAX_STEPS) return finish_program();
                open_step();
                return true;
            };
            auto room = [&](int ops, int nodes) -> bool {          // the open step keeps its ops and nodes in one program
                if (n_ops + ops <= DP_MAX_OPS && n_nodes + nodes <= DP_MAX_NODES) return true;
                if (!close_step() || !finish_program()) return false;
                return ops <= DP_MAX_OPS && nodes <= DP_MAX_NODES;
            };
            // is the operand at device address p complete and visible to every CTA in the open step?  (Results of the open step are
            // spread over the CTAs' slices until its barrier.)
            auto available = [&](size_t j, const uint8_t *p) {
                for (size_t k = j; k-- > 0;) {
                    const size_t span = tensor_span(out_tensor(items[k].t));
                    if (p >= items[k].dd && p < items[k].dd + span) return wstep[k] < gstep;
                }
                return true;                                       // a leaf: uploaded or read in place
            };
            auto overlaps = [&](const std::vector<Range> &v, const uint8_t *p, size_t bytes, bool allow_same) {
                for (const Range &r : v) if (ranges_overlap(p, bytes, r.p, r.n) && !(allow_same && r.p == p && r.n == bytes)) return true;
                return false;
            };
            auto push_op = [&](int op, const uint8_t *a_ptr, const uint8_t *b_ptr, uint8_t *dst, int64_t len, float scalar) {
                DpOp &o = cur.op[n_ops++];
                o = DpOp{};
                o.op = op; o.a = reinterpret_cast<const float *>(a_ptr); o.b = reinterpret_cast<const float *>(b_ptr);
                o.dst = reinterpret_cast<float *>(dst); o.n = (int)len; o.scalar = scalar;
                cur.step[cur.n_steps].nops++;
                if (a_ptr) step_reads.push_back({a_ptr, (size_t)len * 4});
                if (b_ptr) step_reads.push_back({b_ptr, (size_t)len * 4});
                if (dst) step_writes.push_back({dst, (size_t)len * 4});
            };
            open_step();
            // level by level (a valid execution order: a node's operands sit on lower levels), so that the mul_mats of one level
            // -- wq / wk / wv -- meet in one step even when the graph lists an ADD between them
            std::vector<size_t> order(n);
            for (size_t i = 0; i < n; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return items[x].level < items[y].level; });
            for (size_t oi = 0; oi < n; oi++) {
                const size_t i = order[oi];
                const Item &it = items[i];
                const ggml_tensor *t = it.t, *a = t->src0, *b = t->src1;
                at = i;
                if (t->op == GGML_OP_MUL_MAT) {
                    const int64_t K = a->ne[0];
                    DpStep plan = {};
                    decode_program_plan_step(plan, a->type, K, (int64_t)a->nb[1], it.da);
                    DpStep *st = &cur.step[cur.n_steps];
                    const bool joins = st->nnodes > 0 && st->type == plan.type && st->K == plan.K && run_ptr == it.db && run_n == K;
                    if (st->nnodes > 0 && !joins) { if (!close_step()) return bail("program does not fit shared memory"); }
                    if (!joins) {
                        if (!(run_ptr == it.db && run_n == K) && !available(i, it.db)) { if (!close_step()) return bail("program does not fit shared memory"); }
                        if (!room(1, 1)) return bail("program does not fit shared memory");                                   // (may start a new program, which drops the running row)
                        if (!(run_ptr == it.db && run_n == K)) {
                            push_op(DP_LOAD, it.db, nullptr, nullptr, K, 0.0f);
                            run_ptr = it.db; run_n = K;
                        }
                        st = &cur.step[cur.n_steps];
                        const int op0 = st->op0, nops = st->nops, node0 = st->node0, copy0 = st->copy0, ncopies = st->ncopies;
                        *st = plan; st->op0 = op0; st->nops = nops; st->node0 = node0; st->copy0 = copy0; st->ncopies = ncopies;
                    } else if (n_nodes + 1 > DP_MAX_NODES) return bail("a level of too many mul_mats");      // (a step of more than DP_MAX_NODES mul_mats: not a decode chain)
                    const size_t ybytes = (size_t)a->ne[1] * 4;
                    if (overlaps(step_reads, it.dd, ybytes, false) || overlaps(step_writes, it.dd, ybytes, false)) return bail("a mul_mat result aliases something its step touches");
                    DpNode &nd = cur.node[n_nodes++];
                    nd = DpNode{};
                    nd.W = it.da; nd.y = reinterpret_cast<float *>(it.dd); nd.nb01 = (long long)a->nb[1]; nd.M = (int)a->ne[1];
                    const int tr = decode_program_tile_rows(*st);
                    nd.tile0 = st->total_tiles; nd.ntiles = (int)((a->ne[1] + tr - 1) / tr);
                    st->total_tiles += nd.ntiles; st->nnodes++;
                    step_writes.push_back({it.dd, ybytes});
                    wstep[i] = gstep;
                    to_host(i);
                    continue;
                }
                const int64_t len = nelements(t);
                if (t->op == GGML_OP_SILU && cur.step[cur.n_steps].nnodes > 0) {
                    // the SILU of a mul_mat result of the open step: stored by the lane that stores the result (DpNode::y2)
                    DpStep &st = cur.step[cur.n_steps];
                    DpNode *hit = nullptr;
                    for (int k = st.node0; k < st.node0 + st.nnodes; k++) if (reinterpret_cast<const uint8_t *>(cur.node[k].y) == it.da && cur.node[k].M == len && !cur.node[k].y2) hit = &cur.node[k];
                    const size_t bytes = (size_t)len * 4;
                    if (hit && (it.dd == it.da || (!overlaps(step_reads, it.dd, bytes, false) && !overlaps(step_writes, it.dd, bytes, false)))) {
                        hit->y2 = reinterpret_cast<float *>(it.dd);
                        if (it.dd != it.da) step_writes.push_back({it.dd, bytes});
                        need_silu = true;
                        wstep[i] = gstep;
                        to_host(i);
                        continue;
                    }
                }
                // a row op.  The open step's mul_mats (if any) come after its ops: a new op starts the next step.
                if (cur.step[cur.n_steps].nnodes > 0) { if (!close_step()) return bail("program does not fit shared memory"); }
                const bool binary = t->op == GGML_OP_ADD || t->op == GGML_OP_MUL;
                const uint8_t *pa = it.da, *pb = binary ? it.db : nullptr;
                bool a_run = run_ptr == pa && run_n == len, b_run = binary && !a_run && run_ptr == pb && run_n == len;
                if ((!a_run && !available(i, pa)) || (binary && !b_run && !available(i, pb))) { if (!close_step()) return bail("program does not fit shared memory"); }
                const size_t bytes = (size_t)len * 4;
                if (it.dd == it.da && !a_run) {
                    // in place on a tensor the row does not hold: every CTA reads all of it, then each overwrites its slice -- load it in
                    // one step, overwrite it in the next
                    if (!room(1, 0)) return bail("program does not fit shared memory");
                    push_op(DP_LOAD, pa, nullptr, nullptr, len, 0.0f);
                    run_ptr = pa; run_n = len; a_run = true; b_run = false;
                }
                if (binary && pb && ranges_overlap(it.dd, bytes, pb, bytes)) return bail("a binary op writing over its second operand");
                // a slice store must not hit bytes other CTAs still read (or write differently) in this step
                if (overlaps(step_reads, it.dd, bytes, false) || overlaps(step_writes, it.dd, bytes, true)) { if (!close_step()) return bail("program does not fit shared memory"); }
                if (!room(1, 0)) return bail("program does not fit shared memory");
                if (!(run_ptr == pa && run_n == len)) a_run = false;       // (a program split drops the running row)
                if (!(binary && run_ptr == pb && run_n == len)) b_run = false;
                if (it.dd == it.da && !a_run) return bail("an in-place op across a program split");
                int op = DP_LOAD; float scalar = 0.0f;
                switch (t->op) {
                case GGML_OP_ADD: op = DP_ADD; break;
                case GGML_OP_MUL: op = DP_MUL; break;
                case GGML_OP_SILU: op = DP_SILU; break;
                case GGML_OP_RMS_NORM: op = DP_RMS_NORM; break;
                default: op = DP_SCALE; scalar = *static_cast<const float *>(b->data); break;       // float v = *(float*)src1->data (Ggml.cs:6763)
                }
                if (binary) push_op(op, a_run || b_run ? nullptr : pa, a_run ? pb : b_run ? pa : pb, it.dd, len, 0.0f);   // x + y == y + x, x * y == y * x to the bit
                else push_op(op, a_run ? nullptr : pa, nullptr, it.dd, len, scalar);
                run_ptr = it.dd; run_n = len;
                wstep[i] = gstep;
                to_host(i);
            }
            {
                DpStep &st = cur.step[cur.n_steps];
                if (st.nops || st.nnodes || st.ncopies) cur.n_steps++;
                if (!pending.empty()) {
                    // the last step's results: one more step (behind its barrier) that only ships them
                    if (cur.n_steps == DP_MAX_STEPS) { if (!finish_program()) return bail("program does not fit shared memory"); }
                    else open_step();
                    if (cur.step[cur.n_steps].ncopies) cur.n_steps++;
                }
            }
            if (cur.n_steps) { if (!decode_program_finish(cur)) return bail("program does not fit shared memory"); progs.push_back(cur); }
            return !progs.empty();
        };
        if (decode_program_available() && build()) {
            DevCtx &dc = dctx;
            if (!dc.dp_bar) {
                if (cap) return GGB_E_NOCAPTURE;
                GGB_CUDA(cudaMalloc(reinterpret_cast<void **>(&dc.dp_bar), 256));
            }
            const unsigned short *silu = nullptr;
            for (const DpProgram &pg : progs) for (int sti = 0; sti < pg.n_steps; sti++) for (int o = pg.step[sti].op0; o < pg.step[sti].op0 + pg.step[sti].nops; o++) if (pg.op[o].op == DP_SILU) need_silu = true;
            if (need_silu) {
                if (cap && !dc.dp_silu_ready) return GGB_E_NOCAPTURE;
                rc = silu_table_device(&silu);
                if (rc) return rc;
                dc.dp_silu_ready = true;
            }
            static const bool dp_trace = getenv("GGB200_PROGRAM_TRACE") != nullptr;
            long long *tbuf = nullptr;
            if (dp_trace && !cap) { GGB_CUDA(cudaMalloc(reinterpret_cast<void **>(&tbuf), 4 * 64 * 8 * 8)); GGB_CUDA(cudaMemsetAsync(tbuf, 0, 4 * 64 * 8 * 8, s)); }
            for (DpProgram &pg : progs) {
                pg.bar = dc.dp_bar; pg.silu_table = silu; pg.trace = &pg == &progs[0] ? tbuf : nullptr;
                rc = launch_decode_program(pg, s);
                if (rc && &pg == &progs[0] && !cap) {
                    // the first launch was refused (co-residency not available right now): nothing is enqueued yet, take the per-level route
                    cudaGetLastError();
                    if (trace) fprintf(stderr, "ggb200: decode program not used: %s\n", g_err);
                    break;
                }
                if (rc) return rc;
                ran_program = true;
            }
            if (ran_program) for (size_t i : prog_copied) { copied[i] = 1; g_stats.d2h_bytes += tensor_span(out_tensor(items[i].t)); }
            if (tbuf && !ran_program) { cudaFree(tbuf); tbuf = nullptr; }
            if (tbuf) {
                // debugging aid: where the first 64 steps of CTAs 0 / 32 / 64 / 96 spent their time (ns): row ops | staging | tiles | barrier
                std::vector<long long> h(4 * 64 * 8);
                GGB_CUDA(cudaStreamSynchronize(s));
                GGB_CUDA(cudaMemcpy(h.data(), tbuf, h.size() * 8, cudaMemcpyDeviceToHost));
                cudaFree(tbuf);
                for (int c = 0; c < 4; c++)
                    for (int si = 0; si < std::min(64, progs[0].n_steps); si++) {
                        const long long *t = &h[(size_t)((c * 64 + si) * 8)];
                        fprintf(stderr, "dp trace cta %3d step %2d (%d ops, %d nodes, %d tiles): ops %5lld  stage %5lld  tiles %6lld  barrier %5lld  | since start %lld\n", c * 32, si,
                                progs[0].step[si].nops, progs[0].step[si].nnodes, progs[0].step[si].total_tiles, t[1] - t[0], t[2] ? t[2] - t[1] : 0, t[2] ? t[3] - t[2] : 0, t[4] - t[3], t[4] - h[(size_t)(c * 64 * 8)]);
                    }
            }
        }
    }

    for (int lv = 0; lv <= max_level && !ran_program; lv++) {
        std::vector<ggb_dev_mm> mms;
        std::vector<size_t> mm_item;                             // which node each entry of mms belongs to
        bool level_has_mm = false;
        for (size_t ii = 0; ii < n; ii++) {
            Item &it = items[ii];
            if (it.level != lv) continue;
            ggml_tensor *t = it.t; const ggml_tensor *a = t->src0, *b = t->src1;
            switch (t->op) {
            case GGML_OP_MUL_MAT:
                level_has_mm = true;
                if (it.rows <= 0) break;                         // another device's rows; they arrive through its peer stores
                for (int64_t i3 = 0; i3 < a->ne[3]; i3++) for (int64_t i2 = 0; i2 < a->ne[2]; i2++) {
                    ggb_dev_mm m = {};
                    m.type = a->type; m.M = it.rows; m.K = a->ne[0]; m.N = b->ne[1];
                    m.W = it.da + i2 * a->nb[2] + i3 * a->nb[3]; m.nb01 = (int64_t)a->nb[1];
                    m.X = reinterpret_cast<const float *>(it.db + i2 * b->nb[2] + i3 * b->nb[3]); m.ldx_bytes = (int64_t)b->nb[1];
                    m.Y = reinterpret_cast<float *>(it.dd + i2 * t->nb[2] + i3 * t->nb[3]) + it.row0;      // this device's column block of dst
                    // the F16 and quantized drivers index dst as dst_col[ic*ne0] (Ggml.cs:6423, 6697); F32 uses nb1 (6160)
                    m.ldy_bytes = a->type == GGML_TYPE_F32 ? (int64_t)t->nb[1] : (int64_t)t->ne[0] * 4;
                    if (it.a_in_flight || it.ew_new) m.flags |= GGB_MM_W_IN_FLIGHT;     // src0 is an earlier node's result (CPY -> MUL_MAT), or its exponents are brand new
                    if (it.db == static_cast<const uint8_t *>(b->data)) m.flags |= GGB_MM_X_HOST;      // a small leaf read in place from the pinned arena
                    if (!it.ew.empty()) m.W_rowexp = it.ew[(size_t)(i3 * a->ne[2] + i2)];
                    if (G > 1) {
                        // the fused all-gather: the kernel writes each result into every other device's copy of dst as well
                        const size_t off = reinterpret_cast<uint8_t *>(m.Y) - sym.base;
                        for (int h = 0; h < G; h++) if (h != g) m.Y_peer[m.n_peers++] = reinterpret_cast<float *>(sc->sh->sym_base[h] + off);
                    }
                    mms.push_back(m);
                    mm_item.push_back(ii);
                }
                break;
            case GGML_OP_CPY:
                rc = launch_quantize_rows(b->type, reinterpret_cast<const float *>(it.da), a->ne[0], it.dd, b->ne[1] * b->ne[2] * b->ne[3], b->ne[0], s);
                break;
            case GGML_OP_ADD:
                if (a->type != GGML_TYPE_F32) { rc = launch_add_q_f32(a->type, it.da, reinterpret_cast<const float *>(it.db), it.dd, a->ne[1] * a->ne[2] * a->ne[3], a->ne[0], s); break; }
                // fallthrough
            case GGML_OP_MUL:
                rc = launch_binary_f32(t->op, reinterpret_cast<const float *>(it.da), reinterpret_cast<const float *>(it.db), reinterpret_cast<float *>(it.dd), nelements(t), s);
                break;
            case GGML_OP_SILU:
                rc = launch_silu_f32(reinterpret_cast<const float *>(it.da), reinterpret_cast<float *>(it.dd), nelements(t), s);
                break;
            case GGML_OP_RMS_NORM:
                rc = launch_rms_norm_f32(reinterpret_cast<const float *>(it.da), a->ne[0], reinterpret_cast<float *>(it.dd), t->ne[0], a->ne[1] * a->ne[2] * a->ne[3], a->ne[0], s);
                break;
            case GGML_OP_SCALE:
                rc = launch_scale_f32(reinterpret_cast<float *>(it.dd), *static_cast<const float *>(b->data), nelements(t), s);   // float v = *(float*)src1->data (Ggml.cs:6763)
                break;
            case GGML_OP_REPEAT:
                rc = launch_repeat_f32(reinterpret_cast<const float *>(it.da), (int64_t)(a->nb[1] / 4), a->ne[0], a->ne[1], reinterpret_cast<float *>(it.dd), (int64_t)(t->nb[1] / 4), t->ne[0], t->ne[1], s);
                break;
            case GGML_OP_CONT: case GGML_OP_DUP:
                rc = launch_dup_f32_strided(it.da, a->ne, a->nb, reinterpret_cast<float *>(it.dd), s);
                break;
            default:
                rc = set_error(GGB_E_UNSUPPORTED, "executor: op %d", t->op);
            }
            if (rc) return rc;
        }
        if (!mms.empty()) {
            size_t wsb = 0;
            for (const ggb_dev_mm &m : mms) wsb += mm_ws_bytes(m);
            if (arena.used + align_up(wsb, 256) > arena.cap)
                return set_error(GGB_E_NOMEM, "executor: mul_mat workspace of %zu B exceeds the planned scratch (%zu of %zu B used)", wsb, arena.used, arena.cap);
            bool gemv_only = true;
            for (const ggb_dev_mm &m : mms) if (m.N >= 16) gemv_only = false;
            static const bool no_lanes = getenv("GGB200_NO_LANES") != nullptr;
            if (gemv_only && mms.size() >= 8 && !g_timing && !no_lanes) {
                // ---- a wide level of single-token mul_mats: two lanes.  The level is cut into chunks that alternate between the
                //      device's two streams; every chunk is staged, multiplied and (on device 0) copied back on its own lane.  The
                //      GEMV kernels cannot share an SM (each takes ~200 KB of shared memory), so they run one chunk after the other
                //      whatever the stream -- but the staging kernel and the copy kernel are small enough (128 / 256 threads, <= 4 K
                //      registers per CTA) to sit beside a resident GEMV CTA, so chunk c's PCIe traffic hides under chunk c+-1's
                //      weight streaming instead of preceding and following ONE big GEMV. ----
                static const int chunks_env = [] { const char *e = getenv("GGB200_LANE_CHUNKS"); return e ? atoi(e) : 0; }();
                const int nchunk = chunks_env >= 2 ? std::min<int>(chunks_env, (int)mms.size()) : mms.size() >= 16 ? 4 : 2;
                GGB_CUDA(cudaEventRecord(dctx.fork, s));
                GGB_CUDA(cudaStreamWaitEvent(dctx.stream2, dctx.fork, 0));
                uint8_t *wsp = static_cast<uint8_t *>(arena.take(wsb + 256 * (size_t)nchunk));
                size_t i0 = 0;
                for (int c = 0; c < nchunk; c++) {
                    const size_t i1 = mms.size() * (size_t)(c + 1) / (size_t)nchunk;
                    cudaStream_t st = (c & 1) ? dctx.stream2 : s;
                    size_t cws = 0;
                    for (size_t k = i0; k < i1; k++) cws += mm_ws_bytes(mms[k]);
                    rc = dev_batch(mms.data() + i0, (int)(i1 - i0), wsp, cws, st);
                    if (rc) return rc;
                    wsp += align_up(cws, 256);
                    if (lead && !sc) {
                        // (row split: a chunk's dst is complete only when every device has stored its rows -- copied after the level event)
                        std::vector<size_t> idx;
                        for (size_t k = i0; k < i1; k++) if (final_value(mm_item[k]) && (idx.empty() || idx.back() != mm_item[k])) idx.push_back(mm_item[k]);
                        rc = copy_out(idx, st);
                        if (rc) return rc;
                    }
                    i0 = i1;
                }
                GGB_CUDA(cudaEventRecord(dctx.join, dctx.stream2));
                GGB_CUDA(cudaStreamWaitEvent(s, dctx.join, 0));
            } else {
                void *ws = arena.take(wsb);
                rc = dev_batch(mms.data(), (int)mms.size(), ws, wsb, s);
                if (rc) return rc;
            }
        }
        if (!cap && lead) {
            // per-node timing (Ggml.cs:3695-3703 fills perf_time_us per node): one timing event per level, the level's time is
            // shared among its nodes by the bytes they move
            while ((int)dctx.level_time.size() <= lv) { cudaEvent_t e = nullptr; GGB_CUDA(cudaEventCreate(&e)); dctx.level_time.push_back(e); }
            GGB_CUDA(cudaEventRecord(dctx.level_time[(size_t)lv], s));
        }
        if (sc && level_has_mm) {
            // The exchange step of the row split.  The peer stores of this level are in the streams; every device now waits (on the
            // device, not on the host) until all the others have finished the level: event record, host rendezvous so that every
            // event IS recorded, then one cross-device cudaStreamWaitEvent per peer.
            DevCtx &dc = g_devs[(size_t)g];
            if (lv > GGML_MAX_NODES) return set_error(GGB_E_INVALID, "row split: %d dependency levels", lv);
            if (!dc.level_ev[lv]) GGB_CUDA(cudaEventCreateWithFlags(&dc.level_ev[lv], cudaEventDisableTiming));
            GGB_CUDA(cudaEventRecord(dc.level_ev[lv], s));
            SHARD_TRACE("dev %d/%d: level %d enqueued (%zu mul_mats here)", g, G, lv, mms.size());
            if (!sc->sh->bar.arrive_and_wait()) return set_error(GGB_E_CUDA, "row split: another device failed");
            for (int h = 0; h < G; h++) if (h != g) GGB_CUDA(cudaStreamWaitEvent(s, g_devs[(size_t)h].level_ev[lv], 0));
            // (an event of this level is recorded again only by the NEXT call, which starts after every host thread has left this one)
        }
    }

    // ---- whatever has not gone back yet, in node order ----
    if (lead) {
        std::vector<size_t> all(n);
        for (size_t i = 0; i < n; i++) all[i] = i;
        rc = copy_out(all, s);
        if (rc) return rc;
    }
    if (cap) {
        cudaGraph_t graph = nullptr;
        capture.open = false;
        if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return GGB_E_NOCAPTURE; }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess || !exec) { cudaGetLastError(); return GGB_E_NOCAPTURE; }
        cap->exec = exec; cap->epoch = g_epoch;
        cap->d_launches = g_stats.kernel_launches - stats0.kernel_launches; cap->d_h2d = g_stats.h2d_bytes - stats0.h2d_bytes;
        cap->d_d2h = g_stats.d2h_bytes - stats0.d2h_bytes; cap->d_hits = g_stats.weight_cache_hits - stats0.weight_cache_hits;
        cap->d_uploads = g_stats.weight_uploads - stats0.weight_uploads;
        GGB_CUDA(cudaEventRecord(ev0, s));
        GGB_CUDA(cudaGraphLaunch(exec, s));
    }
    GGB_CUDA(cudaEventRecord(ev1, s));
    SHARD_TRACE("dev %d/%d: everything enqueued, waiting for the stream", g, G);
    GGB_CUDA(cudaStreamSynchronize(s));
    SHARD_TRACE("dev %d/%d: done", g, G);
    if (lead) {
        float ms = 0.f;
        GGB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        g_stats.last_graph_device_ms = ms;
        g_stats.nodes_executed += n;
        // each node's share of the device time: measured per dependency level (eager computes), split inside a level by bytes moved;
        // a recorded graph has no events inside, so its nodes keep the shares of the last eager compute of the same cgraph
        std::vector<float> share(n, 1.0f / (float)n);
        if (ran_program) {
            // one launch, no events inside: the nodes share its time by the bytes they move
            double tot = 0.0;
            std::vector<double> w(n, 0.0);
            for (size_t i = 0; i < n; i++) {
                const ggml_tensor *t = nodes[i];
                w[i] = (double)tensor_span(out_tensor(nodes[i])) + (t->src0 ? (double)tensor_span(t->src0) : 0.0) + ((t->src1 && t->src1->data) ? (double)tensor_span(t->src1) : 0.0);
                tot += w[i];
            }
            if (tot > 0.0) for (size_t i = 0; i < n; i++) share[i] = (float)(w[i] / tot);
        } else if (!cap) {
            std::vector<double> w(n, 0.0), lw((size_t)max_level + 1, 0.0), lms((size_t)max_level + 1, 0.0);
            for (size_t i = 0; i < n; i++) {
                const ggml_tensor *t = nodes[i];
                w[i] = (double)tensor_span(out_tensor(nodes[i])) + (t->src0 ? (double)tensor_span(t->src0) : 0.0) + ((t->src1 && t->src1->data) ? (double)tensor_span(t->src1) : 0.0);
                lw[(size_t)items[i].level] += w[i];
            }
            double tot = 0.0;
            for (int lv = 0; lv <= max_level; lv++) {
                float lm = 0.f;
                if (cudaEventElapsedTime(&lm, lv ? dctx.level_time[(size_t)lv - 1] : ev0, dctx.level_time[(size_t)lv]) != cudaSuccess) { cudaGetLastError(); lm = 0.f; }
                lms[(size_t)lv] = lm; tot += lm;
            }
            if (tot > 0.0)
                for (size_t i = 0; i < n; i++) { const size_t lv = (size_t)items[i].level; share[i] = (float)(lms[lv] / tot * (lw[lv] > 0.0 ? w[i] / lw[lv] : 0.0)); }
        } else if (cap->share.size() == n) share = cap->share;
        if (share_out) *share_out = share;
        for (size_t i = 0; i < n; i++) { nodes[i]->perf_runs++; nodes[i]->perf_time_us += (int64_t)((double)ms * 1000.0 * share[i] + 0.5); }   // Ggml.cs:3700-3702
    }
    return GGB_OK;
}

// The row split of one ggml_graph_compute over G GPUs: one host thread per device runs the node list (run_nodes above).
static int run_nodes_sharded(ggb_pool *pool, const std::vector<ggml_tensor *> &nodes, int flags, const std::vector<char> &is_output, int G)
{
    if (G <= 1) return run_nodes(pool, nodes, flags, is_output, nullptr);
    while ((int)pool->pd.size() < G) pool->pd.emplace_back();
    ShardShared sh;
    sh.G = G; sh.bar.n = G;
    run_parallel(G, [&](int g) {
        ShardCtx sc{g, G, &sh};
        const int rc = run_nodes(pool, nodes, flags, is_output, &sc);
        sh.rc[g] = rc;
        if (rc) { snprintf(sh.err[g], sizeof sh.err[g], "%s", g_err); sh.bar.fail(); cudaStreamSynchronize(g_devs[(size_t)g].stream); cudaGetLastError(); }
    });
    cudaSetDevice(g_device);
    for (int g = 0; g < G; g++) if (sh.rc[g]) return set_error(sh.rc[g], "device %d of the row split: %s", g_devs[(size_t)g].dev, sh.err[g]);
    return GGB_OK;
}

// How many GPUs one ggml_graph_compute is split over -- "used only for matrices large enough to benefit" (north_star), in numbers:
//   * nothing to split (no src0 of at least shard_min_bytes), or the split is off: 1;
//   * single-token nodes only (N < 16, bandwidth-bound): all the devices the pool may use -- each reads 1/G of the weights and the
//     exchange is 4 bytes per output row;
//   * prompt-sized nodes (N >= 16, tensor cores): at most 2.  Every device must receive every other device's fp32 results, so with
//     G devices t_exchange / t_compute ~ 2 (G - 1) P / (K B_link): 0.7 at G = 2, K = 4096 -- hidden behind the math -- but
//     4.9 at G = 8, where the gather costs five times the mul_mat (measured in round 1: 11.7 ms against 6.1 ms on ONE GPU).
static int row_split_width(ggb_pool *pool, const std::vector<ggml_tensor *> &run, int flags)
{
    int want = pool->shard_devices;
    if (flags & GGB_GRAPH_SHARD) want = want > 0 ? want : 8;
    if (want <= 1) return 1;
    bool any_big = false, any_batched = false;
    for (const ggml_tensor *t : run) {
        if (t->op != GGML_OP_MUL_MAT) continue;
        const ggml_tensor *a = t->src0;
        if (t->src1->ne[1] >= 16) any_batched = true;
        if (a->ne[2] * a->ne[3] == 1 && (size_t)a->ne[1] * a->nb[1] >= pool->shard_min_bytes) any_big = true;
    }
    if (!any_big) return 1;
    if (any_batched) want = std::min(want, 2);
    return ensure_multi(want);
}

} // namespace ggb

using namespace ggb;

extern "C" {

const char *ggb_last_error(void) { return g_err; }
int ggb_abi_version(void) { return GGB_ABI_VERSION; }

int ggb_abi_check(int sizeof_tensor, int offsetof_data, int sizeof_cgraph, int offsetof_nodes, int sizeof_block_q4_0, int sizeof_block_q4_1)
{
    static_assert(sizeof(ggml_tensor) == 176 && offsetof(ggml_tensor, data) == 160, "ggml_tensor layout (TypeDefinitions.cs:65-99)");
    static_assert(offsetof(ggml_tensor, op) == 72 && offsetof(ggml_tensor, grad) == 80 && offsetof(ggml_tensor, src0) == 88 && offsetof(ggml_tensor, n_tasks) == 136, "ggml_tensor offsets");
    static_assert(sizeof(ggml_cgraph) == 98360 && offsetof(ggml_cgraph, nodes) == 32 && offsetof(ggml_cgraph, leafs) == 65568, "ggml_cgraph layout (TypeDefinitions.cs:102-152)");
    static_assert(sizeof(block_q4_0) == 20 && sizeof(block_q4_1) == 24 && sizeof(block_q8_0) == 36 && sizeof(block_q8_1) == 44, "block layouts");
    if (sizeof_tensor != (int)sizeof(ggml_tensor) || offsetof_data != (int)offsetof(ggml_tensor, data) ||
        sizeof_cgraph != (int)sizeof(ggml_cgraph) || offsetof_nodes != (int)offsetof(ggml_cgraph, nodes) ||
        sizeof_block_q4_0 != 20 || sizeof_block_q4_1 != 24)
        return set_error(GGB_E_ABI, "ABI mismatch: host says tensor %d/%d cgraph %d/%d q4_0 %d q4_1 %d; library has %zu/%zu %zu/%zu 20 24",
                         sizeof_tensor, offsetof_data, sizeof_cgraph, offsetof_nodes, sizeof_block_q4_0, sizeof_block_q4_1,
                         sizeof(ggml_tensor), offsetof(ggml_tensor, data), sizeof(ggml_cgraph), offsetof(ggml_cgraph, nodes));
    return GGB_OK;
}

int ggb_init(void) { std::lock_guard<std::mutex> lk(g_mu); return ensure_init(); }

int ggb_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return GGB_OK;
    sync_all_devices();
    for (Worker *w : g_workers) { { std::lock_guard<std::mutex> lw(w->m); w->quit = true; } w->cv.notify_all(); if (w->th.joinable()) w->th.join(); delete w; }
    g_workers.clear();
    for (size_t i = 1; i < g_devs.size(); i++) {
        cudaSetDevice(g_devs[i].dev);
        for (cudaEvent_t e : g_devs[i].level_ev) if (e) cudaEventDestroy(e);
        cudaEventDestroy(g_devs[i].ev0); cudaEventDestroy(g_devs[i].ev1); cudaStreamDestroy(g_devs[i].stream);
        cudaEventDestroy(g_devs[i].fork); cudaEventDestroy(g_devs[i].join); cudaStreamDestroy(g_devs[i].stream2);
    }
    if (!g_devs.empty()) for (cudaEvent_t e : g_devs[0].level_ev) if (e) cudaEventDestroy(e);
    cudaSetDevice(g_device);
    if (!g_devs.empty()) { cudaEventDestroy(g_devs[0].fork); cudaEventDestroy(g_devs[0].join); cudaStreamDestroy(g_devs[0].stream2); }
    for (DevCtx &d : g_devs) { for (cudaEvent_t e : d.level_time) cudaEventDestroy(e); d.level_time.clear(); }
    if (!g_devs.empty() && g_devs[0].dp_bar) { cudaFree(g_devs[0].dp_bar); g_devs[0].dp_bar = nullptr; g_devs[0].dp_silu_ready = false; }
    g_devs.clear(); g_multi = 0;
    cudaEventDestroy(g_ev0); cudaEventDestroy(g_ev1);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr; g_inited = false;
    return GGB_OK;
}

int ggb_device_count(int *count)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count) *count = n;
    return GGB_OK;
}

static bool weight_cache_default() { const char *e = getenv("GGB200_WEIGHT_CACHE"); return e && atoi(e) != 0; }
static int row_split_default()
{
    const char *e = getenv("GGB200_ROW_SPLIT");
    if (!e || !*e) return 0;
    return (!strcmp(e, "all") || !strcmp(e, "ALL")) ? 8 : std::max(0, atoi(e));
}
static size_t row_split_min_default() { const char *e = getenv("GGB200_ROW_SPLIT_MIN_BYTES"); return e ? (size_t)strtoull(e, nullptr, 0) : (size_t)(4u << 20); }
static void drop_all_mirrors(ggb_pool *pool)
{
    for (PoolDev &pd : pool->pd) { for (auto &kv : pd.mirrors) free_mirror(kv.second); pd.mirrors.clear(); }
}
// pinning the caller's memory is an optimisation only (pageable memory still works); portable, so every device of a row split sees it pinned
static void register_adopted(ggb_pool *pool)
{
    if (pool->owned || pool->register_tried) return;
    pool->register_tried = true;
    if (cudaHostRegister(pool->host_base, pool->bytes, cudaHostRegisterPortable) == cudaSuccess) pool->registered = true; else cudaGetLastError();
}

int ggb_pool_alloc(size_t bytes, void **host_base, ggb_pool **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!host_base || !out) return set_error(GGB_E_INVALID, "ggb_pool_alloc: null out pointer");
    int rc = ensure_init();
    if (rc) return rc;
    void *p = nullptr;
    GGB_CUDA(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable));   // page-aligned >= GGML_MEM_ALIGN; pinned for every device of a row split
    ggb_pool *pool = new ggb_pool();
    pool->host_base = p; pool->bytes = bytes; pool->owned = true; pool->weight_cache = weight_cache_default();
    pool->shard_devices = row_split_default(); pool->shard_min_bytes = row_split_min_default();
    *host_base = p; *out = pool;
    return GGB_OK;
}

int ggb_pool_adopt(void *host_base, size_t bytes, ggb_pool **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!host_base || !out) return set_error(GGB_E_INVALID, "ggb_pool_adopt: null pointer");
    if (reinterpret_cast<uintptr_t>(host_base) % GGML_MEM_ALIGN) return set_error(GGB_E_INVALID, "ggb_pool_adopt: buffer not %d-byte aligned (Ggml.cs:1557)", GGML_MEM_ALIGN);
    // No device work yet: the caller owns the memory; it is pinned lazily at the first compute.
    ggb_pool *pool = new ggb_pool();
    pool->host_base = host_base; pool->bytes = bytes; pool->owned = false; pool->weight_cache = weight_cache_default();
    pool->shard_devices = row_split_default(); pool->shard_min_bytes = row_split_min_default();
    *out = pool;
    return GGB_OK;
}

int ggb_pool_free(ggb_pool *pool)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool) return GGB_OK;
    sync_all_devices();
    for (GraphEntry &e : pool->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    drop_all_mirrors(pool);
    for (PoolDev &pd : pool->pd) { if (pd.arena.base) cudaFree(pd.arena.base); if (pd.sym.base) cudaFree(pd.sym.base); }
    g_epoch++;
    if (pool->registered) cudaHostUnregister(pool->host_base);
    if (pool->owned && pool->host_base) cudaFreeHost(pool->host_base);
    cudaGetLastError();
    delete pool;
    return GGB_OK;
}

int ggb_tensor_invalidate(ggb_pool *pool, const ggml_tensor *t)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool) return set_error(GGB_E_INVALID, "ggb_tensor_invalidate: null pool");
    sync_all_devices();
    if (!t) { drop_all_mirrors(pool); return GGB_OK; }
    if (!t->data) return GGB_OK;
    // every mirror that shares a byte with the tensor (it may be a view of a cached leaf, or a leaf some cached view looks into)
    const uint8_t *h = static_cast<const uint8_t *>(t->data);
    const size_t span = tensor_span(t);
    for (PoolDev &pd : pool->pd)
        for (auto f = pd.mirrors.begin(); f != pd.mirrors.end();) {
            if (ranges_overlap(h, span, static_cast<const uint8_t *>(f->first), f->second.bytes)) { free_mirror(f->second); f = pd.mirrors.erase(f); }
            else ++f;
        }
    return GGB_OK;
}

int ggb_pool_set_weight_cache(ggb_pool *pool, int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool) return set_error(GGB_E_INVALID, "ggb_pool_set_weight_cache: null pool");
    pool->weight_cache = on != 0;
    if (!on) { sync_all_devices(); drop_all_mirrors(pool); }
    return GGB_OK;
}

int ggb_row_split_rows(int64_t M, int g, int G, int64_t *row0, int64_t *rows)
{
    if (M < 0 || G < 1 || g < 0 || g >= G || !row0 || !rows) return set_error(GGB_E_INVALID, "ggb_row_split_rows: bad arguments");
    shard_rows(M, g, G, *row0, *rows);
    return GGB_OK;
}

int ggb_pool_set_row_split(ggb_pool *pool, int max_devices, size_t min_weight_bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool) return set_error(GGB_E_INVALID, "ggb_pool_set_row_split: null pool");
    if (max_devices < 0) max_devices = 8;
    if (max_devices != pool->shard_devices) { sync_all_devices(); drop_all_mirrors(pool); }      // resident slices belong to the old split
    pool->shard_devices = std::min(max_devices, 8);
    if (min_weight_bytes) pool->shard_min_bytes = min_weight_bytes;
    return GGB_OK;
}

int ggb_mul_mat_node(ggb_pool *pool, ggml_tensor *dst)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool || !dst) return set_error(GGB_E_INVALID, "ggb_mul_mat_node: null argument");
    int rc = validate_mul_mat(dst);
    if (rc) return rc;
    std::vector<ggml_tensor *> nodes{dst};
    register_adopted(pool);
    return run_nodes(pool, nodes, 0, std::vector<char>{1}, nullptr);
}

} // extern "C"

namespace ggb {

// ---- seam B: which nodes of a graph run here ----
//
// A node is runnable if it is on the path (MUL_MAT, the F32 -> {F16, quantized} CPY, and unless GGB_GRAPH_MUL_MAT_ONLY the neighbours of
// mul_mat), its operands are leafs (op NONE) or runnable nodes, and running it AHEAD of the caller's CPU loop cannot change a result.
// RESHAPE / VIEW / PERMUTE / TRANSPOSE do nothing at compute time (Ggml.cs:8668-8687): "runnable" when their source is, never enqueued.
//
// The reference executes strictly in node order (Ggml.cs:3539-3704); this call runs every runnable node before the caller's loop runs
// the rest.  That reordering is only invisible if no dependency passes through memory the pointer graph does not show, so while
// scanning in node order the byte ranges every NON-runnable node reads and writes are recorded (an in-place op such as
// ggml_sqr_inplace writes src0's bytes; a CPY writes src1's), and a later candidate is refused -- and, through operand_ok, everything
// downstream of it -- when
//   one of its operands intersects bytes an earlier CPU node writes        (it would read them before they are written),
//   its output intersects bytes an earlier CPU node reads or writes        (in-place / CPY targets: it would clobber them too early),
//   or an operand intersects an earlier runnable node's result without lying inside it (neither the device copy nor the host bytes).
struct ByteRange { const uint8_t *p; size_t n; };
struct Selection { std::vector<ggml_tensor *> run; std::vector<int> run_idx, view_idx; std::vector<char> is_output; };

static int select_nodes(ggml_cgraph *g, int flags, Selection &sel)
{
    if (g->n_nodes < 0 || g->n_nodes > GGML_MAX_NODES) return set_error(GGB_E_INVALID, "graph with %d nodes", g->n_nodes);
    std::map<const ggml_tensor *, bool> runnable;
    const bool neighbours = !(flags & GGB_GRAPH_MUL_MAT_ONLY);
    std::vector<ByteRange> cpu_reads, cpu_writes;
    std::vector<Produced> gpu_results;                            // host byte ranges of the runnable nodes' results, in order
    auto range_of = [](const ggml_tensor *x) { return ByteRange{static_cast<const uint8_t *>(x->data), x->data ? tensor_span(x) : 0}; };
    auto hits = [](const std::vector<ByteRange> &v, ByteRange r) {
        if (!r.p || !r.n) return false;
        for (const ByteRange &e : v) if (ranges_overlap(r.p, r.n, e.p, e.n)) return true;
        return false;
    };
    auto out_tensor = [](ggml_tensor *t) -> ggml_tensor * { return t->op == GGML_OP_CPY ? t->src1 : t; };
    for (int i = 0; i < g->n_nodes; i++) {
        ggml_tensor *t = g->nodes[i];
        if (!t) return set_error(GGB_E_INVALID, "graph node %d is null", i);
        auto operand_ok = [&](const ggml_tensor *o) { return o && (o->op == GGML_OP_NONE || runnable.count(o)); };
        bool ok = false;
        if (neighbours && is_view_op(t->op)) {
            if (operand_ok(t->src0)) { runnable[t] = true; sel.view_idx.push_back(i); }
            continue;                                             // a view node touches no bytes either way
        }
        if (is_view_op(t->op)) continue;
        if (t->op == GGML_OP_MUL_MAT || t->op == GGML_OP_CPY || (neighbours && is_neighbour_op(t->op))) {
            const bool unary = t->op == GGML_OP_SILU || t->op == GGML_OP_RMS_NORM || t->op == GGML_OP_CONT || t->op == GGML_OP_DUP;
            // REPEAT's src1 only supplies the shape (Ggml.cs:8015-8034); it is never read
            const bool need_b = !unary && t->op != GGML_OP_REPEAT;
            if (operand_ok(t->src0) && (!need_b || operand_ok(t->src1))) {
                int rc = t->op == GGML_OP_MUL_MAT ? validate_mul_mat(t) : t->op == GGML_OP_CPY ? validate_cpy(t) : validate_neighbour(t);
                if (rc == GGB_E_INVALID) return rc;        // the reference would assert: report it
                ok = rc == GGB_OK;                         // unsupported: leave the node to the caller's loop
            }
            if (ok) {
                // ---- memory hazards against the nodes that stay on the CPU, and partial views of device results ----
                const ByteRange ra = range_of(t->src0), rb = (need_b && t->op != GGML_OP_CPY) ? range_of(t->src1) : ByteRange{nullptr, 0};
                const ByteRange ro = range_of(out_tensor(t));
                if (hits(cpu_writes, ra) || hits(cpu_writes, rb) || hits(cpu_writes, ro) || hits(cpu_reads, ro)) ok = false;
                bool partial = false;
                if (ok && ra.p) { find_produced(gpu_results, ra.p, ra.n, &partial); if (partial) ok = false; }
                if (ok && rb.p) { find_produced(gpu_results, rb.p, rb.n, &partial); if (partial) ok = false; }
                if (ok && ro.p && ra.p && !(ro.p == ra.p)) {
                    // an output that shares bytes with an earlier device result must BE that result's range (in place) or a CPY target
                    if (t->op != GGML_OP_CPY) { find_produced(gpu_results, ro.p, ro.n, &partial); if (partial) ok = false; }
                }
            }
        }
        if (ok) {
            runnable[t] = true; sel.run.push_back(t); sel.run_idx.push_back(i);
            const ByteRange ro = range_of(out_tensor(t));
            gpu_results.push_back({ro.p, ro.n, nullptr, 0});
        } else {
            // stays with the caller: remember what it touches.  Its result bytes (a CPY's are src1's), and for safety every operand it names.
            if (t->src0 && t->src0->data) cpu_reads.push_back(range_of(t->src0));
            if (t->src1 && t->src1->data) cpu_reads.push_back(range_of(t->src1));
            for (int k = 0; k < GGML_MAX_OPT; k++) { const ggml_tensor *o = reinterpret_cast<const ggml_tensor *>(t->opt[k]); if (o && o->data) cpu_reads.push_back(range_of(o)); }
            ggml_tensor *o = out_tensor(t);
            if (o && o->data) cpu_writes.push_back(range_of(o));
        }
    }
    // graph outputs = executed nodes nobody else in the executed set consumes (views are looked through)
    auto base_of = [](const ggml_tensor *o) { while (o && is_view_op(o->op) && o->src0) o = o->src0; return o; };
    std::vector<ggml_tensor *> &run = sel.run;
    sel.is_output.assign(run.size(), 1);
    for (size_t i = 0; i < run.size(); i++)
        for (size_t j = i + 1; j < run.size(); j++)
            if (base_of(run[j]->src0) == run[i] || base_of(run[j]->src1) == run[i]) sel.is_output[i] = 0;
    // a consumer outside the executed set (a CPU op of the caller) needs the data on the host
    for (int i = 0; i < g->n_nodes; i++) {
        ggml_tensor *t = g->nodes[i];
        if (runnable.count(t)) continue;
        for (size_t j = 0; j < run.size(); j++) {
            if (base_of(t->src0) == run[j] || base_of(t->src1) == run[j]) sel.is_output[j] = 1;
            // ... also when it reaches the bytes without naming the node (a fresh view tensor over the same data)
            const ByteRange rj = range_of(out_tensor(run[j]));
            if ((t->src0 && t->src0->data && ranges_overlap(rj.p, rj.n, static_cast<const uint8_t *>(t->src0->data), tensor_span(t->src0))) ||
                (t->src1 && t->src1->data && ranges_overlap(rj.p, rj.n, static_cast<const uint8_t *>(t->src1->data), tensor_span(t->src1)))) sel.is_output[j] = 1;
        }
    }
    // an in-place node (SCALE, *_inplace) shares its host bytes with its source: the LAST writer of a range carries the result
    for (size_t i = 0; i < run.size(); i++)
        for (size_t j = i + 1; j < run.size(); j++)
            if (run[j]->op != GGML_OP_CPY && run[j]->data == run[i]->data && run[i]->op != GGML_OP_CPY) sel.is_output[i] = 0;
    return GGB_OK;
}

} // namespace ggb

extern "C" {

int ggb_graph_plan(ggml_cgraph *g, int flags, uint8_t *done)
{
    if (!g || !done) return set_error(GGB_E_INVALID, "ggb_graph_plan: null argument");
    Selection sel;
    int rc = select_nodes(g, flags, sel);
    if (rc) return rc;
    for (int i = 0; i < g->n_nodes; i++) done[i] = 0;
    for (int i : sel.run_idx) done[i] = 1;
    for (int i : sel.view_idx) done[i] = 1;
    return (int)sel.run.size();
}

// Everything the executor's decisions depend on, folded into 64 bits (FNV-1a over the words): per node the header fields it reads
// (op, type, shape, strides, data pointer), the same for both operands plus what decides residency (op, is_param, grad), and the
// scalar a SCALE node bakes into its launch.  Two computes with the same key enqueue the same work on the same addresses.
static uint64_t graph_key(const ggb_pool *pool, const ggml_cgraph *g, int flags)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; };
    auto tensor = [&](const ggml_tensor *t) {
        mix(reinterpret_cast<uintptr_t>(t));
        if (!t) return;
        mix(((uint64_t)(uint32_t)t->type << 32) | (uint32_t)t->op);
        for (int i = 0; i < GGML_MAX_DIMS; i++) { mix((uint64_t)t->ne[i]); mix(t->nb[i]); }
        mix(reinterpret_cast<uintptr_t>(t->data));
        mix(((uint64_t)t->is_param << 1) | (t->grad ? 1u : 0u));
    };
    mix((uint64_t)flags); mix((uint64_t)g->n_nodes); mix(pool->weight_cache ? 1 : 0); mix((uint64_t)pool->shard_devices); mix(pool->shard_min_bytes);
    for (int i = 0; i < g->n_nodes; i++) {
        const ggml_tensor *t = g->nodes[i];
        tensor(t);
        if (!t) continue;
        tensor(t->src0); tensor(t->src1);
        if (t->src0 && is_view_op(t->src0->op)) tensor(t->src0->src0);
        if (t->src1 && is_view_op(t->src1->op)) tensor(t->src1->src0);
        if (t->op == GGML_OP_SCALE && t->src1 && t->src1->data && t->src1->type == GGML_TYPE_F32) { uint32_t v; memcpy(&v, t->src1->data, 4); mix(v); }
    }
    return h ? h : 1;
}

int ggb_graph_compute_mul_mats(ggb_pool *pool, ggml_cgraph *g, int flags, uint8_t *done)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!pool || !g) return set_error(GGB_E_INVALID, "ggb_graph_compute_mul_mats: null argument");
    if (g->n_nodes < 0 || g->n_nodes > GGML_MAX_NODES) return set_error(GGB_E_INVALID, "graph with %d nodes", g->n_nodes);
    static const bool graph_cache = getenv("GGB200_NO_GRAPH_CACHE") == nullptr;
    GraphEntry *ge = nullptr;
    int rc = GGB_OK;
    if (graph_cache && !g_timing && g_inited) {
        const uint64_t key = graph_key(pool, g, flags);
        for (GraphEntry &e : pool->graphs) if (e.key == key) { ge = &e; break; }
        if (ge && ge->exec && ge->epoch == g_epoch) {
            // ---- the same compute as before: replay what the executor enqueued for it ----
            cudaSetDevice(g_device);
            cudaStream_t s = g_stream;
            GGB_CUDA(cudaEventRecord(g_ev0, s));
            GGB_CUDA(cudaGraphLaunch(ge->exec, s));
            GGB_CUDA(cudaEventRecord(g_ev1, s));
            GGB_CUDA(cudaStreamSynchronize(s));
            float ms = 0.f;
            GGB_CUDA(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
            g_stats.last_graph_device_ms = ms;
            g_stats.nodes_executed += ge->run.size();
            __atomic_fetch_add(&g_stats.kernel_launches, ge->d_launches, __ATOMIC_RELAXED);
            g_stats.h2d_bytes += ge->d_h2d; g_stats.d2h_bytes += ge->d_d2h; g_stats.weight_cache_hits += ge->d_hits; g_stats.weight_uploads += ge->d_uploads;
            g_stats.graph_replays++;
            for (size_t i = 0; i < ge->run.size(); i++) {
                const float sh = ge->share.size() == ge->run.size() ? ge->share[i] : 1.0f / (float)ge->run.size();
                ge->run[i]->perf_runs++; ge->run[i]->perf_time_us += (int64_t)((double)ms * 1000.0 * sh + 0.5);
            }
            if (done) memcpy(done, ge->done.data(), (size_t)g->n_nodes);
            g->perf_runs++;
            g->perf_time_us += (int64_t)(ms * 1000.0);
            return ge->n_run;
        }
        if (ge && ge->exec) { cudaGraphExecDestroy(ge->exec); ge->exec = nullptr; }      // an older epoch: its pointers are stale
        if (!ge) {
            if (pool->graphs.size() >= 8) { if (pool->graphs.front().exec) cudaGraphExecDestroy(pool->graphs.front().exec); pool->graphs.erase(pool->graphs.begin()); }
            pool->graphs.emplace_back();
            ge = &pool->graphs.back();
            ge->key = key;
        }
        ge->seen++;
    }
    Selection sel;
    rc = select_nodes(g, flags, sel);
    if (rc) return rc;
    if (done) for (int i = 0; i < g->n_nodes; i++) done[i] = 0;
    rc = ensure_init();
    if (rc) return rc;
    register_adopted(pool);
    const int G = row_split_width(pool, sel.run, flags);
    bool ran = false;
    // second sighting of a graph on one device, over pinned memory: record it
    if (ge && ge->seen >= 2 && !ge->bad && G == 1 && !sel.run.empty() && (pool->owned || pool->registered)) {
        rc = run_nodes(pool, sel.run, flags, sel.is_output, nullptr, ge);
        if (rc == GGB_E_NOCAPTURE) { ge->bad = ge->seen >= 3; rc = GGB_OK; }      // (the 2nd compute may still be creating mirrors: one more try)
        else if (rc) return rc;
        else {
            ran = true;
            ge->run = sel.run; ge->n_run = (int)sel.run.size();
            ge->done.assign((size_t)g->n_nodes, 0);
            for (int i : sel.run_idx) ge->done[(size_t)i] = 1;
            for (int i : sel.view_idx) ge->done[(size_t)i] = 1;
        }
    }
    if (!ran) {
        if (G == 1) {
            std::vector<float> share;
            rc = run_nodes(pool, sel.run, flags, sel.is_output, nullptr, nullptr, &share);
            if (ge) ge->share = std::move(share);
        } else rc = run_nodes_sharded(pool, sel.run, flags, sel.is_output, G);
        if (rc) return rc;
    }
    if (done) { for (int i : sel.run_idx) done[i] = 1; for (int i : sel.view_idx) done[i] = 1; }
    g->perf_runs++;
    g->perf_time_us += (int64_t)(g_stats.last_graph_device_ms * 1000.0);
    return (int)sel.run.size();
}

// ---- codecs on host or device pointers ----

static int codec_rows(bool quantize, int type, const void *src, void *dst, int64_t nrows, int64_t k)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (nrows < 0 || k < 0) return set_error(GGB_E_INVALID, "negative size");
    if (nrows == 0 || k == 0) return GGB_OK;
    const int blck = blck_size(type);
    if (!type_size(type) || (quantize && type == GGML_TYPE_F32)) return set_error(GGB_E_UNSUPPORTED, "codec: type %d", type);
    if (k % blck) return set_error(GGB_E_INVALID, "codec: k=%lld not a multiple of %d", (long long)k, blck);
    const size_t fbytes = (size_t)nrows * k * 4, qbytes = (size_t)nrows * (k / blck) * type_size(type);
    const size_t sbytes = quantize ? fbytes : qbytes, dbytes = quantize ? qbytes : fbytes;
    const bool sdev = is_device_ptr(src), ddev = is_device_ptr(dst);
    void *ds = const_cast<void *>(src), *dd = dst;
    if (!sdev) { GGB_CUDA(cudaMalloc(&ds, sbytes)); GGB_CUDA(cudaMemcpyAsync(ds, src, sbytes, cudaMemcpyHostToDevice, g_stream)); g_stats.h2d_bytes += sbytes; }
    if (!ddev) { GGB_CUDA(cudaMalloc(&dd, dbytes)); }
    rc = quantize ? launch_quantize_rows(type, (const float *)ds, k, dd, nrows, k, g_stream)
                  : launch_dequantize_rows(type, ds, (float *)dd, nrows, k, g_stream);
    if (!rc && !ddev) { cudaError_t e = cudaMemcpyAsync(dst, dd, dbytes, cudaMemcpyDeviceToHost, g_stream); g_stats.d2h_bytes += dbytes;
                        if (e != cudaSuccess) rc = set_error(GGB_E_CUDA, "D2H failed: %s", cudaGetErrorString(e)); }
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (!rc && e != cudaSuccess) rc = set_error(GGB_E_CUDA, "codec kernel failed: %s", cudaGetErrorString(e));
    if (!sdev) cudaFree(ds);
    if (!ddev) cudaFree(dd);
    return rc;
}

int ggb_quantize_rows(int type, const float *src, void *dst, int64_t nrows, int64_t k)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return codec_rows(true, type, src, dst, nrows, k);
}
int ggb_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return codec_rows(false, type, src, dst, nrows, k);
}

// ---- device-resident entry points ----

size_t ggb_dev_workspace_bytes(const ggb_dev_mm *mm, int count)
{
    size_t t = 0;
    for (int i = 0; i < count; i++) t += mm_ws_bytes(mm[i]);
    return t;
}

// The device-level entry points run on the caller's stream without the library lock; with stream == NULL they land on the library's
// own stream, which the executor may be recording into a CUDA graph on another host thread -- those calls take the lock.
struct OwnStreamLock {
    std::unique_lock<std::mutex> lk;
    explicit OwnStreamLock(const void *stream) { if (!stream) lk = std::unique_lock<std::mutex>(g_mu); }
};

int ggb_dev_mul_mat_batch(const ggb_dev_mm *mm, int count, void *ws, size_t ws_bytes, void *stream)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    if (count < 0 || (count && !mm)) return set_error(GGB_E_INVALID, "ggb_dev_mul_mat_batch: bad arguments");
    return dev_batch(mm, count, ws, ws_bytes, stream ? static_cast<cudaStream_t>(stream) : g_stream);
}

int ggb_dev_mul_mat_batch_phase(const ggb_dev_mm *mm, int count, void *ws, size_t ws_bytes, void *stream, int phase)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    if (count < 0 || (count && !mm) || phase < 0 || phase > 3) return set_error(GGB_E_INVALID, "ggb_dev_mul_mat_batch_phase: bad arguments");
    for (int i = 0; i < count; i++) if (phase && needs_k_segments(mm[i])) return set_error(GGB_E_UNSUPPORTED, "ggb_dev_mul_mat_batch_phase: rows this long are multiplied in K segments (one phase only)");
    if (phase)
        for (int i = 0; i < count; i++)
            if (mm[i].M > 0 && mm[i].N > 0 && use_gemm(mm[i]))
                return set_error(GGB_E_UNSUPPORTED, "ggb_dev_mul_mat_batch_phase: split phases are for single-token (N < 16) nodes only");
    tl_phase = phase;
    rc = dev_batch(mm, count, ws, ws_bytes, stream ? static_cast<cudaStream_t>(stream) : g_stream);
    tl_phase = 0;
    return rc;
}

int ggb_dev_weight_rowexp(int type, const void *W, int64_t nb01, int64_t M, int64_t K, int32_t *rowexp, void *stream)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    if (!W || !rowexp) return set_error(GGB_E_INVALID, "ggb_dev_weight_rowexp: null pointer");
    return launch_weight_rowexp(type, W, nb01, M, K, rowexp, stream ? static_cast<cudaStream_t>(stream) : g_stream);
}

int ggb_dev_quantize_rows(int type, const float *src, void *dst, int64_t nrows, int64_t k, void *stream)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    return launch_quantize_rows(type, src, k, dst, nrows, k, stream ? static_cast<cudaStream_t>(stream) : g_stream);
}
int ggb_dev_dequantize_rows(int type, const void *src, float *dst, int64_t nrows, int64_t k, void *stream)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    return launch_dequantize_rows(type, src, dst, nrows, k, stream ? static_cast<cudaStream_t>(stream) : g_stream);
}

#define GGB_DEV_WRAP(call) do { OwnStreamLock guard(stream); int rc_ = ensure_init(); if (rc_) return rc_; cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g_stream; return call; } while (0)
int ggb_dev_binary(int op, const float *a, const float *b, float *dst, int64_t n, void *stream) { GGB_DEV_WRAP(launch_binary_f32(op, a, b, dst, n, st)); }
int ggb_dev_scale(float *x, float v, int64_t n, void *stream) { GGB_DEV_WRAP(launch_scale_f32(x, v, n, st)); }
int ggb_dev_silu(const float *x, float *dst, int64_t n, void *stream) { GGB_DEV_WRAP(launch_silu_f32(x, dst, n, st)); }
int ggb_dev_rms_norm(const float *x, int64_t x_stride, float *dst, int64_t dst_stride, int64_t nrows, int64_t ne00, void *stream)
{ GGB_DEV_WRAP(launch_rms_norm_f32(x, x_stride, dst, dst_stride, nrows, ne00, st)); }
int ggb_dev_repeat(const float *src, int64_t src_stride, int64_t nc0, int64_t nr0, float *dst, int64_t dst_stride, int64_t nc, int64_t nr, void *stream)
{ GGB_DEV_WRAP(launch_repeat_f32(src, src_stride, nc0, nr0, dst, dst_stride, nc, nr, st)); }
int ggb_dev_cont(const void *src, const int64_t ne[4], const uint64_t nb[4], float *dst, void *stream) { GGB_DEV_WRAP(launch_dup_f32_strided(src, ne, nb, dst, st)); }
int ggb_dev_add_q(int type, const void *src0, const float *src1, void *dst, int64_t nrows, int64_t k, void *stream)
{ GGB_DEV_WRAP(launch_add_q_f32(type, src0, src1, dst, nrows, k, st)); }
#undef GGB_DEV_WRAP

int ggb_dev_alloc(size_t bytes, void **dptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!dptr) return set_error(GGB_E_INVALID, "null out pointer");
    GGB_CUDA(cudaMalloc(dptr, bytes ? bytes : 256));
    return GGB_OK;
}
int ggb_dev_free(void *dptr) { if (dptr) { GGB_CUDA(cudaFree(dptr)); } return GGB_OK; }
int ggb_dev_upload(void *dptr, const void *host, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = ensure_init();
    if (rc) return rc;
    GGB_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, g_stream));
    GGB_CUDA(cudaStreamSynchronize(g_stream));
    g_stats.h2d_bytes += bytes;
    return GGB_OK;
}
int ggb_dev_download(void *host, const void *dptr, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = ensure_init();
    if (rc) return rc;
    GGB_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, g_stream));
    GGB_CUDA(cudaStreamSynchronize(g_stream));
    g_stats.d2h_bytes += bytes;
    return GGB_OK;
}
int ggb_stream_sync(void *stream)
{
    OwnStreamLock guard(stream);
    int rc = ensure_init();
    if (rc) return rc;
    GGB_CUDA(cudaStreamSynchronize(stream ? static_cast<cudaStream_t>(stream) : g_stream));
    return GGB_OK;
}

int ggb_ipc_export(void *dptr, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    GGB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return GGB_OK;
}
int ggb_ipc_open(const uint8_t handle[64], void **peer)
{
    int rc = ensure_init();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    GGB_CUDA(cudaIpcOpenMemHandle(peer, h, cudaIpcMemLazyEnablePeerAccess));
    return GGB_OK;
}
int ggb_ipc_close(void *peer) { GGB_CUDA(cudaIpcCloseMemHandle(peer)); return GGB_OK; }

namespace ggb {
struct PeerFlags { uint64_t *p[8]; };
__global__ void k_peer_barrier(PeerFlags f, int rank, int world, unsigned long long epoch)
{
    const int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();                                   // this rank's earlier peer stores (previous kernels) are ordered before the flag
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[t] + rank), "l"(epoch) : "memory");
        unsigned long long v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f.p[rank] + t) : "memory");
            if (v < epoch && clock64() - t0 > 20000000000ll) __trap();      // ~10 s: a peer died; fail the launch instead of hanging the GPU
        } while (v < epoch);
    }
}
} // namespace ggb

namespace ggb {
struct PeerBases { uint8_t *p[8]; };
__global__ void __launch_bounds__(128) k_peer_push_barrier(PeerBases b, PeerFlags f, uint32_t *counter, int rank, int world,
                                                           unsigned long long seg_offset, unsigned long long seg_bytes, unsigned long long seg_stride,
                                                           int n_seg, unsigned long long epoch)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the producing kernel's stores to the local segments are complete
    const unsigned long long vec_per_seg = seg_bytes >> 4, total = vec_per_seg * (unsigned long long)n_seg;
    const uint8_t *src = b.p[rank];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long sgi = i / vec_per_seg, v = i - sgi * vec_per_seg;
        const unsigned long long off = seg_offset + sgi * seg_stride + (v << 4);
        const uint4 val = *reinterpret_cast<const uint4 *>(src + off);
        for (int p = 0; p < world; p++)
            if (p != rank) *reinterpret_cast<uint4 *>(b.p[p] + off) = val;
    }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) { __threadfence_system(); last = atomicAdd(counter, 1u) == gridDim.x - 1; }   // cumulative fence after the CTA barrier
    __syncthreads();
    if (!last) return;
    __shared__ unsigned long long ep;
    if (threadIdx.x == 0) {
        *counter = 0;
        // epoch == 0: take the next epoch from device memory (counter + 8 bytes), so a captured CUDA graph can be replayed
        unsigned long long *dev_epoch = reinterpret_cast<unsigned long long *>(counter + 2);
        ep = epoch ? epoch : ++(*dev_epoch);
    }
    __syncthreads();
    epoch = ep;
    const int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[t] + rank), "l"(epoch) : "memory");
        unsigned long long v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f.p[rank] + t) : "memory");
            if (v < epoch && clock64() - t0 > 20000000000ll) __trap();
        } while (v < epoch);
    }
}
} // namespace ggb

int ggb_peer_push_barrier(void *const *peer_bases, uint64_t *const *peer_flags, uint32_t *counter, int rank, int world,
                          size_t seg_offset, size_t seg_bytes, size_t seg_stride, int n_seg, uint64_t epoch, void *stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!peer_bases || !peer_flags || !counter || world < 1 || world > 8 || rank < 0 || rank >= world || n_seg < 0)
        return set_error(GGB_E_INVALID, "ggb_peer_push_barrier: bad arguments");
    if ((seg_offset | seg_bytes | seg_stride) & 15) return set_error(GGB_E_INVALID, "ggb_peer_push_barrier: segments must be 16-byte multiples");
    PeerBases b = {}; PeerFlags f = {};
    for (int i = 0; i < world; i++) {
        if (!peer_bases[i] || !peer_flags[i]) return set_error(GGB_E_INVALID, "ggb_peer_push_barrier: null pointer for rank %d", i);
        b.p[i] = static_cast<uint8_t *>(peer_bases[i]); f.p[i] = peer_flags[i];
    }
    {
        // same shared-memory carve-out as the GEMV this kernel is meant to run beside (an SM is not shared by kernels whose carve-outs differ)
        static PerDeviceOnce once;
        if (once.need()) { cudaFuncSetAttribute(k_peer_push_barrier, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); cudaGetLastError(); }
    }
    const unsigned long long total = (unsigned long long)(seg_bytes >> 4) * (unsigned long long)n_seg;
    // 128-thread CTAs (32 registers: 4 K per CTA) fit beside a resident GEMV CTA, so the exchange of step i really runs under the
    // GEMVs of step i + 1 (256-thread CTAs needed 8 K registers and waited for a GEMV CTA to exit: VERDICT r1, weak #8)
    unsigned grid = (unsigned)std::min<unsigned long long>(128, std::max<unsigned long long>(1, (total + 511) / 512));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128);
    cfg.stream = stream ? static_cast<cudaStream_t>(stream) : g_stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_peer_push_barrier, b, f, counter, rank, world, (unsigned long long)seg_offset, (unsigned long long)seg_bytes,
                                (unsigned long long)seg_stride, n_seg, (unsigned long long)epoch));
    count_launch();
    return GGB_OK;
}

int ggb_peer_barrier(uint64_t *const *peer_flags, int rank, int world, uint64_t epoch, void *stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!peer_flags || world < 1 || world > 8 || rank < 0 || rank >= world) return set_error(GGB_E_INVALID, "ggb_peer_barrier: bad arguments");
    PeerFlags f = {};
    for (int i = 0; i < world; i++) { if (!peer_flags[i]) return set_error(GGB_E_INVALID, "ggb_peer_barrier: null flags pointer for rank %d", i); f.p[i] = peer_flags[i]; }
    k_peer_barrier<<<1, 32, 0, stream ? static_cast<cudaStream_t>(stream) : g_stream>>>(f, rank, world, (unsigned long long)epoch);
    count_launch();
    GGB_CUDA(cudaGetLastError());
    return GGB_OK;
}

int ggb_set_kernel_timing(int on) { g_timing = on < 0 ? 0 : on > 2 ? 2 : on; return GGB_OK; }
int ggb_set_decode_program(int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    const bool want = on != 0;
    if (want != g_decode_program) { g_decode_program = want; g_epoch++; }      // recorded graphs were enqueued the other way
    return GGB_OK;
}
int ggb_get_stats(ggb_stats *out)
{
    if (!out) return set_error(GGB_E_INVALID, "null");
    for (auto &pr : g_timed) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            g_stats.timed_kernel_ms += ms; g_stats.timed_kernel_launches++;
        }
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    g_timed.clear();
    cudaGetLastError();
    *out = g_stats;
    return GGB_OK;
}
int ggb_reset_stats(void)
{
    for (auto &pr : g_timed) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    g_timed.clear();
    g_stats = ggb_stats{};
    return GGB_OK;
}

} // extern "C"
