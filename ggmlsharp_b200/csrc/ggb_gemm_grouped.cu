// Batched (N >= 16) mul_mat over a whole GROUP of independent nodes in one launch: persistent CTA pairs walk the 256 x 128
// output tiles of up to GGB_GEMM_GROUP_NODES Q4_0 / Q4_1 / F16 nodes (ggml_compute_forward_mul_mat_q_f32 / _f16_f32,
// Ggml.cs:6440-6712, 6180-6438).
//
// Why (profiles/README.md, round-1 GEMM notes): the per-node kernel k_gemm_q spends ~550 cycles per 128-wide K step against a
// 512-cycle tensor-pipe floor, but around that mainloop every launch pays ~3 k cycles of prologue (barrier init, TMEM
// allocation, descriptor fetch, first TMA round trip), a launch gap, and -- worst -- whatever part of the machine the node's
// own tile count leaves idle: 4096 x 512 gives 64 pair-tiles for 74 SM pairs, a row-split shard of it on 8 GPUs gives 8.  A
// graph level of MUL_MAT nodes is independent work, so here the tile list of ALL its nodes is dealt round-robin to the
// resident pairs: the pipeline (TMA rings, TMEM A stages, barrier phases) keeps running across tile and node boundaries,
// and the TMA producers prefetch the next tile's operands while the 16 dequant warps drain the accumulators of the
// current one.
//
// Tile = one CTA pair (tcgen05 cta_group::2): 256 weight rows x 128 activation rows, K step 128.  Warp roles, TMEM layout,
// the nibble -> fp16 -> tcgen05.st A path and the even/odd dual MMA issuers are those of k_gemm_q (ggb_gemm.cu); what is
// new is the tile loop, the running K-step counter that indexes every ring, and the ACC_EMPTY barrier that lets the issuers
// overwrite the accumulators only after both CTAs have read them.
#include "ggb_tc.cuh"

#include <algorithm>
#include <vector>
#include <cstdio>
#include <cstdlib>

namespace ggb {

namespace {

using namespace tc;

struct alignas(64) GroupNode {
    CUtensorMap map_w, map_x;      // raw quant bytes [M][row_bytes] (box 128 rows x 4 blocks); fp16 activations [Npad][K] (box 64 x 64, swizzle-128B)
    float *Y; long long ldy;       // dst, row stride in floats
    int M, N, ksteps;
    int tile0, tile_end, nt;       // this node's tiles in the launch are [tile0, tile_end); nt = tiles along the activation rows
    int n_peers, pad_;
    long long peer_delta[7];       // row split: byte offset from Y to the same element of peer p's dst (CUDA-IPC mapped)
    const int *ew, *ex;            // power-of-two row exponents of the weights / staged activations (null, null: true F16 weights)
};
// CAP = 8 keeps the parameter block at 3 KB for small groups (parameters above 4 KB add microseconds to the launch)
// wait_w: some node's weights (or ew) were written earlier in this stream by a kernel of the same batch, so the weight side waits too
template <int CAP> struct alignas(64) GemmGroupT { int n_nodes, total_tiles, wait_w; int pad_[13]; GroupNode node[CAP]; };
using GemmGroup = GemmGroupT<GGB_GEMM_GROUP_NODES>;
constexpr int SMALL_GROUP = 8;
static_assert(sizeof(GemmGroup) <= 32000, "kernel parameter space");

// BN = 128: two issuer warps (even / odd K steps) with a 128-column accumulator each, as in k_gemm_q.
// BN = 256: every dequantized A stage (and every raw byte) is reused for twice the columns, and one m256n256k16 instruction keeps
// the tensor pipe busy for ~128 cycles, so ONE issuer thread is enough (a thread that also waits and commits issues an MMA per
// ~125 cycles, profiles/r01_mma_rate.txt) and the single 256-column accumulator takes the TMEM the two half-width ones had.
template <int BN, int RAW_ROW> struct GroupedCfg {
    static constexpr int NISSUE = BN == 256 ? 1 : 2;
    static constexpr int B_BYTES = (BN / 2) * BK * 2;
    // raw ring: a multiple of the 4 dequant groups (see ggb_gemm.cu); 8 stages unless shared memory runs out (Q8_0 with 32 KB B stages)
    static constexpr int RAW_STAGES = (1024 + 4 * B_BYTES + 8 * BM * RAW_ROW + 64 * 8 + 16 <= 227 * 1024) ? 8 : 4;
    static constexpr size_t SMEM = 1024 + (size_t)4 * B_BYTES + (size_t)RAW_STAGES * (BM * RAW_ROW) + 64 * 8 + 16;
};

template <int TYPE, int CAP, int BN>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_q_grouped(const __grid_constant__ GemmGroupT<CAP> G)
{
    constexpr int BNL = BN / 2;
    constexpr int RAW_ROW = RawRow<TYPE>::BYTES;
    using Cfg = GroupedCfg<BN, RAW_ROW>;
    constexpr int NISSUE = Cfg::NISSUE;
    constexpr int RAW_BYTES = BM * RAW_ROW, B_BYTES = BNL * BK * 2;
    constexpr int RAW_STAGES = Cfg::RAW_STAGES, A_STAGES = 4;
    constexpr int TMEM_COLS = 512, A_COL0 = NISSUE * BN;      // [0,256): one 256-column accumulator or the even-K / odd-K pair; then 4 x 64 columns of A stages

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;
    uint8_t *sRaw = sB + A_STAGES * B_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRaw + RAW_STAGES * RAW_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    constexpr int RAW_FULL = 0, RAW_EMPTY = 8, A_FULL = 16, A_EMPTY = 20, ACC_FULL = 24, ACC_EMPTY = 25, NBARS = 26;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    auto BAR = [&](int i) { return bar0 + 8 * i; };
    auto LBAR = [&](int i) { return mapa_u32(BAR(i), 0); };   // the same barrier in the leader CTA
    constexpr int RAW_SH = RAW_STAGES == 8 ? 3 : 2;           // log2(RAW_STAGES): phase = (gk >> RAW_SH) & 1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RAW_STAGES; i++) { mbar_init(BAR(RAW_FULL + i), 1); mbar_init(BAR(RAW_EMPTY + i), 4); }
        for (int i = 0; i < A_STAGES; i++) { mbar_init(BAR(A_FULL + i), 4 * 2 + 2); mbar_init(BAR(A_EMPTY + i), 1); }
        mbar_init(BAR(ACC_FULL), NISSUE);
        mbar_init(BAR(ACC_EMPTY), 2 * NDQ_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next group's activation kernel may start beside this GEMM

    // Every role walks the same tile sequence: t = pair, pair + npairs, ...; tiles of a node are ordered n-tile fastest so
    // that concurrently running pairs share weight rows in L2.
    struct Tile { const GroupNode *nd; int m0, n0, ksteps; };
    int scan_n = 0;
    auto locate = [&](int t) -> Tile {
        while (t >= G.node[scan_n].tile_end) scan_n++;
        const GroupNode &nd = G.node[scan_n];
        const int local = t - nd.tile0, mt = local / nd.nt, ntile = local - mt * nd.nt;
        return Tile{&nd, (mt * 2 + (int)rank) * BM, ntile * BN, nd.ksteps};
    };

    if (warp == 0) {
        // ===== TMA producer 1: this CTA's 128 rows of raw quant blocks =====
        if (lane == 0) {
            if (G.wait_w) asm volatile("griddepcontrol.wait;" ::: "memory");
            uint32_t gk = 0;
            for (int t = pair; t < G.total_tiles; t += npairs) {
                const Tile tl = locate(t);
                for (int ks = 0; ks < tl.ksteps; ks++, gk++) {
                    const int s = gk & (RAW_STAGES - 1);
                    mbar_wait(BAR(RAW_EMPTY + s), ((gk >> RAW_SH) & 1) ^ 1);
                    mbar_expect_tx(BAR(RAW_FULL + s), RAW_BYTES);
                    tma_load_2d(smem_u32(sRaw + s * RAW_BYTES), &tl.nd->map_w, BAR(RAW_FULL + s), RawRow<TYPE>::box_x(ks), tl.m0);
                }
            }
        }
    } else if (warp == 2) {
        // ===== TMA producer 2: this CTA's half (64 rows) of the activation tile =====
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");       // the activation kernel must have finished; weights never wait
            uint32_t gk = 0;
            for (int t = pair; t < G.total_tiles; t += npairs) {
                const Tile tl = locate(t);
                const int n0 = tl.n0 + (int)rank * BNL;
                for (int ks = 0; ks < tl.ksteps; ks++, gk++) {
                    const int sb = gk & (A_STAGES - 1);
                    mbar_wait(BAR(A_EMPTY + sb), ((gk >> 2) & 1) ^ 1);
                    const uint32_t full = LBAR(A_FULL + sb);
                    if (leader) mbar_expect_tx(BAR(A_FULL + sb), 2 * B_BYTES); else mbar_arrive_cluster(full);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES), &tl.nd->map_x, full, ks * BK, n0);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES + BNL * 128), &tl.nd->map_x, full, ks * BK + 64, n0);
                }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (leader CTA): even / odd K steps of every tile, own accumulator each (BN = 256: warp 1 alone) =====
        if (leader && lane == 0 && (NISSUE == 2 || warp == 1)) {
            const int me = warp >> 1;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * 2) >> 4) << 24);
            const uint32_t acc = tmem + (uint32_t)(me * BN);
            const uint64_t bdesc0 = make_sdesc(smem_u32(sB));
            uint32_t gk0 = 0, it = 0;
            for (int t = pair; t < G.total_tiles; t += npairs, it++) {
                const Tile tl = locate(t);
                if (it) { mbar_wait(BAR(ACC_EMPTY), (it - 1) & 1); tc_fence_after(); }    // both CTAs have drained the previous tile's accumulators
                for (int ks = me; ks < tl.ksteps; ks += NISSUE) {
                    const uint32_t gk = gk0 + (uint32_t)ks;
                    const int g = gk & 3;
                    mbar_wait(BAR(A_FULL + g), (gk >> 2) & 1);           // A rows in TMEM (both CTAs) + both halves of the B tile
                    tc_fence_after();
                    const uint64_t bd0 = bdesc0 + (uint64_t)((g * B_BYTES) >> 4);
                    const uint32_t a_base = tmem + A_COL0 + g * 64;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        tc_mma_f16_ts(acc, a_base + k * 8, bd0 + (uint64_t)(((k >> 2) * (BNL * 128) + (k & 3) * 32) >> 4), idesc, (ks >= NISSUE) || k != 0, true);
                    tc_commit_cg2(BAR(A_EMPTY + g));                      // frees TMEM stage g and B stage g in both CTAs
                }
                tc_commit_cg2(BAR(ACC_FULL));
                gk0 += (uint32_t)tl.ksteps;
            }
        }
    } else {
        // ===== 16 dequant warps (TMEM lane quadrant q, K-step group g), which are also the epilogue =====
        const int q = warp & 3, g = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t raw_row = smem_u32(sRaw) + (uint32_t)(r * RAW_ROW);
        const uint32_t a_full = LBAR(A_FULL + g), acc_empty = LBAR(ACC_EMPTY);
        const uint32_t a_tmem = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_COL0 + g * 64);
        uint32_t mk_lo = 0x000F000Fu, mk_hi = 0x00F000F0u, mg_lo = 0x64006400u, mg_hi = 0x54005400u;
        asm volatile("" : "+r"(mk_lo), "+r"(mk_hi), "+r"(mg_lo), "+r"(mg_hi));
        uint32_t gk0 = 0, it = 0;
        int ew_next = 0;
        if (G.wait_w) asm volatile("griddepcontrol.wait;" ::: "memory");       // ew comes from a kernel launched earlier in this batch
        for (int t = pair; t < G.total_tiles; t += npairs, it++) {
            const Tile tl = locate(t);
            // this thread's weight row of the tile: its exponent was fetched while the previous tile's accumulators were drained
            const int ewr = it ? ew_next : ((tl.m0 + r) < tl.nd->M ? __ldcg(tl.nd->ew + tl.m0 + r) : 0);
            const float rs = exp2i(-ewr);
            for (int ks = (int)((g - gk0) & 3); ks < tl.ksteps; ks += 4) {
                const uint32_t gk = gk0 + (uint32_t)ks;
                const int s = gk & (RAW_STAGES - 1);
                mbar_wait(BAR(RAW_FULL + s), (gk >> RAW_SH) & 1);
                uint32_t w[RAW_ROW / 4];
                load_raw_row<TYPE>(raw_row + (uint32_t)(s * RAW_BYTES), ks, w);
                // the raw stage is released only after the dequant below has CONSUMED these registers (ggb_gemm.cu)
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {                         // two blocks (64 K = 32 columns) per TMEM store
                    uint32_t v[32];
#pragma unroll
                    for (int jb = 0; jb < 2; jb++) {
                        dequant_group<TYPE>(w, hb * 2 + jb, rs, mk_lo, mk_hi, mg_lo, mg_hi, &v[jb * 16]);
                    }
                    if (hb == 0) {
                        // the nibble expansion of the first half ran ahead of this wait: only the TMEM store needs the stage free
                        mbar_wait(BAR(A_EMPTY + g), ((gk >> 2) & 1) ^ 1);
                        tc_fence_after();
                    }
                    tmem_st_x32(a_tmem + (uint32_t)(hb * 32), v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive_cluster(a_full); mbar_arrive(BAR(RAW_EMPTY + s)); }
            }
            gk0 += (uint32_t)tl.ksteps;

            // ---- epilogue of this tile: warp -> (lane quadrant q, 32-column groups g, g+4, ...) ----
            const GroupNode &nd = *tl.nd;
            // Issued BEFORE the wait for the accumulators, so their latency hides behind the last MMAs: lane l's factor 2^ex of column
            // cb*32 + l of each of this warp's column groups (broadcast per column by SHFL below), and the NEXT tile's weight-row exponent.
            if (it == 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // ex was written by the activation kernel (complete: the B tiles came from it)
            float fxl[BN / 128];
#pragma unroll
            for (int j = 0; j < BN / 128; j++) fxl[j] = exp2i(__ldcg(nd.ex + tl.n0 + (g + 4 * j) * 32 + lane));
            if (t + npairs < G.total_tiles) {
                const Tile nx = locate(t + npairs);
                ew_next = (nx.m0 + r) < nx.nd->M ? __ldcg(nx.nd->ew + nx.m0 + r) : 0;
            }
            const float fw = exp2i(ewr);
            mbar_wait(BAR(ACC_FULL), it & 1);
            tc_fence_after();
            const int m = tl.m0 + q * 32 + lane;
            const bool two = NISSUE == 2 && tl.ksteps > 1;
#pragma unroll 1
            for (int cb = g; cb < BN / 32; cb += NDQ_WARPS / 4) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32);
                const float fx_cb = (BN == 256 && cb != g) ? fxl[BN / 128 - 1] : fxl[0];     // a select, not a dynamically indexed register array
#pragma unroll 1
                for (int hc = 0; hc < 2; hc++) {                      // 16 columns at a time keeps both accumulators in 32 registers
                    uint32_t v[16], u[16];
#define GGB_TMEM_LD16(ARR, ADDR) \
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                                 : "=r"(ARR[0]), "=r"(ARR[1]), "=r"(ARR[2]), "=r"(ARR[3]), "=r"(ARR[4]), "=r"(ARR[5]), "=r"(ARR[6]), "=r"(ARR[7]), \
                                   "=r"(ARR[8]), "=r"(ARR[9]), "=r"(ARR[10]), "=r"(ARR[11]), "=r"(ARR[12]), "=r"(ARR[13]), "=r"(ARR[14]), "=r"(ARR[15]) \
                                 : "r"(ADDR) : "memory")
                    GGB_TMEM_LD16(v, taddr + (uint32_t)(hc * 16));
                    if (NISSUE == 2) GGB_TMEM_LD16(u, taddr + (uint32_t)(BN + hc * 16));
#undef GGB_TMEM_LD16
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    {
#pragma unroll
                        for (int c = 0; c < 16; c++) {
                            const int n = tl.n0 + cb * 32 + hc * 16 + c;
                            float res = two ? __uint_as_float(v[c]) + __uint_as_float(u[c]) : __uint_as_float(v[c]);   // even-K + odd-K partial sums
                            res = scale_pair(res, fw, __shfl_sync(0xffffffffu, fx_cb, hc * 16 + c));   // undo the operands' power-of-two pre-scaling (exact); every lane shuffles
                            if (m < nd.M && n < nd.N) {
                                float *yp = nd.Y + (long long)n * nd.ldy + m;
                                *yp = res;
                                for (int pp = 0; pp < nd.n_peers; pp++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + nd.peer_delta[pp]) = res;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty);            // this warp's TMEM reads of the tile are complete
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// F16 weights: both operands come from shared memory (TMA, swizzle-128B), so the 16 warps that dequantize in the Q4 kernel
// have nothing to do during the main loop and TMEM has room for TWO accumulator sets (2 x (even-K + odd-K) x 128 columns):
// the epilogue of tile i runs under the main loop of tile i + 1.
template <int CAP>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_f16_grouped(const __grid_constant__ GemmGroupT<CAP> G)
{
    constexpr int BN = 128, BNL = 64;
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNL * BK * 2, STAGES = 4;
    constexpr int TMEM_COLS = 512;                             // accumulator set b: columns [256 b, 256 b + 128) even-K, [+128, +256) odd-K

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + STAGES * A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + STAGES * B_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    constexpr int FULL = 0, EMPTY = 4, ACC_FULL = 8, ACC_EMPTY = 10, NBARS = 12;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    auto BAR = [&](int i) { return bar0 + 8 * i; };
    auto LBAR = [&](int i) { return mapa_u32(BAR(i), 0); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(BAR(FULL + i), 2); mbar_init(BAR(EMPTY + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(BAR(ACC_FULL + i), 2); mbar_init(BAR(ACC_EMPTY + i), 2 * NDQ_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    struct Tile { const GroupNode *nd; int m0, n0, ksteps; };
    int scan_n = 0;
    auto locate = [&](int t) -> Tile {
        while (t >= G.node[scan_n].tile_end) scan_n++;
        const GroupNode &nd = G.node[scan_n];
        const int local = t - nd.tile0, mt = local / nd.nt, ntile = local - mt * nd.nt;
        return Tile{&nd, (mt * 2 + (int)rank) * BM, ntile * BN, nd.ksteps};
    };

    if (warp == 0) {
        // ===== TMA producer: this CTA's 128 weight rows and its half of the activation rows, per K step =====
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");       // activations come from the preceding kernel
            uint32_t gk = 0;
            for (int t = pair; t < G.total_tiles; t += npairs) {
                const Tile tl = locate(t);
                const int n0 = tl.n0 + (int)rank * BNL;
                for (int ks = 0; ks < tl.ksteps; ks++, gk++) {
                    const int sb = gk & (STAGES - 1);
                    mbar_wait(BAR(EMPTY + sb), ((gk >> 2) & 1) ^ 1);
                    const uint32_t full = LBAR(FULL + sb);
                    if (leader) mbar_expect_tx(BAR(FULL + sb), 2 * (A_BYTES + B_BYTES)); else mbar_arrive_cluster(full);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES), &tl.nd->map_x, full, ks * BK, n0);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES + BNL * 128), &tl.nd->map_x, full, ks * BK + 64, n0);
                    tma_load_2d_cg2(smem_u32(sA + sb * A_BYTES), &tl.nd->map_w, full, ks * BK, tl.m0);
                    tma_load_2d_cg2(smem_u32(sA + sb * A_BYTES + BM * 128), &tl.nd->map_w, full, ks * BK + 64, tl.m0);
                }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (leader CTA): even / odd K steps, accumulator set it & 1 =====
        if (leader && lane == 0) {
            const int me = warp >> 1;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * 2) >> 4) << 24);
            const uint64_t adesc0 = make_sdesc(smem_u32(sA)), bdesc0 = make_sdesc(smem_u32(sB));
            uint32_t gk0 = 0, it = 0;
            for (int t = pair; t < G.total_tiles; t += npairs, it++) {
                const Tile tl = locate(t);
                const int ab = it & 1;
                const uint32_t acc = tmem + (uint32_t)(ab * 2 * BN + me * BN);
                if (it >= 2) { mbar_wait(BAR(ACC_EMPTY + ab), ((it >> 1) - 1) & 1); tc_fence_after(); }
                for (int ks = me; ks < tl.ksteps; ks += 2) {
                    const uint32_t gk = gk0 + (uint32_t)ks;
                    const int sb = gk & (STAGES - 1);
                    mbar_wait(BAR(FULL + sb), (gk >> 2) & 1);
                    tc_fence_after();
                    const uint64_t ad0 = adesc0 + (uint64_t)((sb * A_BYTES) >> 4), bd0 = bdesc0 + (uint64_t)((sb * B_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        const uint64_t ad = ad0 + (uint64_t)(((k >> 2) * (BM * 128) + (k & 3) * 32) >> 4);
                        const uint64_t bd = bd0 + (uint64_t)(((k >> 2) * (BNL * 128) + (k & 3) * 32) >> 4);
                        tc_mma_f16_cg2(acc, ad, bd, idesc, (ks >= 2) || k != 0);
                    }
                    tc_commit_cg2(BAR(EMPTY + sb));
                }
                tc_commit_cg2(BAR(ACC_FULL + ab));
                gk0 += (uint32_t)tl.ksteps;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warps: tile it drains accumulator set it & 1 while the issuers fill the other =====
        const int q = warp & 3, g = (warp - 4) >> 2;
        uint32_t it = 0;
        for (int t = pair; t < G.total_tiles; t += npairs, it++) {
            const Tile tl = locate(t);
            const int ab = it & 1;
            const uint32_t acc_empty = LBAR(ACC_EMPTY + ab);
            mbar_wait(BAR(ACC_FULL + ab), (it >> 1) & 1);
            tc_fence_after();
            const GroupNode &nd = *tl.nd;
            const int m = tl.m0 + q * 32 + lane;
            const bool two = tl.ksteps > 1;
            // expanded quantized weights (ggb_shim.cu: use_gemm_expanded) carry row exponents; true F16 weights do not
            const int *__restrict__ exn = nd.ex;
            float fw = 1.0f, fxl = 1.0f;
            if (exn) {
                if (it == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
                if (m < nd.M) fw = exp2i(__ldcg(nd.ew + m));
                fxl = exp2i(__ldcg(exn + tl.n0 + g * 32 + lane));      // BN = 128: one column group per warp
            }
#pragma unroll 1
            for (int cb = g; cb < BN / 32; cb += NDQ_WARPS / 4) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 2 * BN + cb * 32);
#pragma unroll 1
                for (int hc = 0; hc < 2; hc++) {
                    uint32_t v[16], u[16];
#define GGB_TMEM_LD16(ARR, ADDR) \
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                                 : "=r"(ARR[0]), "=r"(ARR[1]), "=r"(ARR[2]), "=r"(ARR[3]), "=r"(ARR[4]), "=r"(ARR[5]), "=r"(ARR[6]), "=r"(ARR[7]), \
                                   "=r"(ARR[8]), "=r"(ARR[9]), "=r"(ARR[10]), "=r"(ARR[11]), "=r"(ARR[12]), "=r"(ARR[13]), "=r"(ARR[14]), "=r"(ARR[15]) \
                                 : "r"(ADDR) : "memory")
                    GGB_TMEM_LD16(v, taddr + (uint32_t)(hc * 16));
                    GGB_TMEM_LD16(u, taddr + (uint32_t)(BN + hc * 16));
#undef GGB_TMEM_LD16
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    {
#pragma unroll
                        for (int c = 0; c < 16; c++) {
                            const int n = tl.n0 + cb * 32 + hc * 16 + c;
                            float res = two ? __uint_as_float(v[c]) + __uint_as_float(u[c]) : __uint_as_float(v[c]);
                            if (exn) res = scale_pair(res, fw, __shfl_sync(0xffffffffu, fxl, hc * 16 + c));      // warp-uniform branch; every lane shuffles
                            if (m < nd.M && n < nd.N) {
                                float *yp = nd.Y + (long long)n * nd.ldy + m;
                                *yp = res;
                                for (int pp = 0; pp < nd.n_peers; pp++) *reinterpret_cast<float *>(reinterpret_cast<char *>(yp) + nd.peer_delta[pp]) = res;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}


template <int TYPE, int CAP, int BN = 128>
int launch_grouped(const GemmGroupT<CAP> &G, cudaStream_t s)
{
    constexpr int RAW_ROW = RawRow<TYPE>::BYTES;
    constexpr size_t smem = TYPE == GGML_TYPE_F16 ? 1024 + (size_t)4 * (BM * BK * 2) + (size_t)4 * (64 * BK * 2) + 64 * 8 + 16
                                                  : GroupedCfg<BN, RAW_ROW>::SMEM;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    void (*kern)(const GemmGroupT<CAP>) = nullptr;
    if constexpr (TYPE == GGML_TYPE_F16) kern = k_gemm_f16_grouped<CAP>; else kern = k_gemm_q_grouped<TYPE, CAP, BN>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    // The tile lists are static, so every pair of the grid must be resident at once: ask how many 2-CTA clusters the device
    // can actually co-schedule (a GPC with an odd number of usable SMs leaves one unpaired) instead of assuming SMs / 2.
    static int max_pairs_dev[16] = {};                          // per device: the GPUs of one box need not have the same usable pairs
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    int &max_pairs = max_pairs_dev[cur_dev & 15];
    if (!max_pairs) {
        cfg.gridDim = dim3((unsigned)(device_sm_count() / 2 * 2));
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = device_sm_count() / 2; }
        if (const char *e = getenv("GGB200_GEMM_PAIRS")) nc = std::max(1, std::min(nc, atoi(e)));
        max_pairs = nc;
        if (getenv("GGB200_VERBOSE")) fprintf(stderr, "[ggb200] grouped GEMM: %d resident CTA pairs\n", nc);
    }
    const int pairs = std::min(G.total_tiles, max_pairs);
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // prologue + weight streaming overlap the activation kernel
    at[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("GGB200_NO_PDL") != nullptr;
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 1 : 2;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, kern, G));
    count_launch();
    return GGB_OK;
}

} // namespace

bool gemm_grouped_supported(int type) { return type == GGML_TYPE_Q4_0 || type == GGML_TYPE_Q4_1 || type == GGML_TYPE_F16 || is_q_weight(type); }

int launch_gemm_grouped(const GemmArgs *args, int count, cudaStream_t s)
{
    if (count <= 0) return GGB_OK;
    if (count > GGB_GEMM_GROUP_NODES) return set_error(GGB_E_INVALID, "grouped GEMM: %d nodes > %d", count, GGB_GEMM_GROUP_NODES);
    static thread_local GemmGroup G;                                   // ~25 KB: kept off the stack
    const int type = args[0].type;
    G.n_nodes = count; G.total_tiles = 0; G.wait_w = 0;
    // Quantized weights: 256-column tiles halve the dequant work and the weight traffic per flop (a K step of a 256-wide tile costs
    // ~1.78x a 128-wide one, measured on cfg 4: 986 -> 1 074 TFLOP/s), but they also halve the number of tiles, and the static
    // round-robin deal makes the slowest pair the launch time (cfg 3 batch of 8: 512 tiles = 6.9 rounds of 74 pairs, 256 tiles = 3.5
    // -> 4 rounds).  So both tilings are dealt out on paper -- K steps plus an epilogue per tile, in units of a 128-wide K step --
    // and the cheaper one is launched.
    int bn = 128;
    if (type != GGML_TYPE_F16) {
        static const int pairs_est = [] { int p = std::max(1, device_sm_count() / 2); if (const char *e = getenv("GGB200_GEMM_PAIRS")) p = std::max(1, std::min(p, atoi(e))); return p; }();
        auto dealt = [&](int tile_n, double kstep_cost, double epi_cost) {
            std::vector<double> load((size_t)pairs_est, 0.0);
            long long t = 0;
            for (int i = 0; i < count; i++) {
                const long long tiles = ((args[i].M + 2 * BM - 1) / (2 * BM)) * ((args[i].N + tile_n - 1) / tile_n);
                const double per_tile = (double)((args[i].K + BK - 1) / BK) * kstep_cost + epi_cost;
                for (long long k = 0; k < tiles; k++, t++) load[(size_t)(t % pairs_est)] += per_tile;
            }
            return *std::max_element(load.begin(), load.end());
        };
        if (dealt(256, 1.78, 13.0) < dealt(128, 1.0, 6.5)) bn = 256;
        static const int force = [] { const char *e = getenv("GGB200_GEMM_BN"); return e ? atoi(e) : 0; }();
        if (force == 128 || force == 256) bn = force;
    }
    for (int i = 0; i < count; i++) {
        const GemmArgs &a = args[i];
        if (a.type != type) return set_error(GGB_E_INVALID, "grouped GEMM: mixed weight types in one group");
        if (a.n_peers < 0 || a.n_peers > 7) return set_error(GGB_E_INVALID, "grouped GEMM: n_peers=%d", a.n_peers);
        GroupNode &nd = G.node[i];
        int rc;
        if (type == GGML_TYPE_F16) {
            rc = make_map_2d(&nd.map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.W, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.nb01, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        } else {
            const int raw_row = (type == GGML_TYPE_Q4_0 || type == GGML_TYPE_Q4_2) ? 80 : type == GGML_TYPE_Q8_0 ? 144 : 96;
            const uint64_t row_bytes = (uint64_t)(a.K / GGB_QK) * q32_bytes(type);
            rc = make_map_2d(&nd.map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.W, row_bytes, (uint64_t)a.M, (uint64_t)a.nb01, (uint32_t)raw_row, BM, CU_TENSOR_MAP_SWIZZLE_NONE);
        }
        if (rc) return rc;
        rc = make_map_2d(&nd.map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.Xh, (uint64_t)a.K, (uint64_t)a.Npad, (uint64_t)a.K * 2, 64, (uint32_t)(bn / 2), CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        nd.Y = a.Y; nd.ldy = a.ldy; nd.M = (int)a.M; nd.N = (int)a.N; nd.ksteps = (int)((a.K + BK - 1) / BK);
        nd.nt = (int)((a.N + bn - 1) / bn);
        const int mt = (int)((a.M + 2 * BM - 1) / (2 * BM));
        nd.tile0 = G.total_tiles; nd.tile_end = nd.tile0 + mt * nd.nt;
        G.total_tiles = nd.tile_end;
        nd.ew = a.ew; nd.ex = a.ex;
        if (type != GGML_TYPE_F16 && (!a.ew || !a.ex)) return set_error(GGB_E_INVALID, "grouped GEMM: quantized weights need their row exponents (ew / ex)");
        if ((a.ew == nullptr) != (a.ex == nullptr)) return set_error(GGB_E_INVALID, "grouped GEMM: ew and ex come together");
        if (a.wait_w) G.wait_w = 1;
        nd.n_peers = a.n_peers;
        for (int p = 0; p < a.n_peers; p++) nd.peer_delta[p] = (long long)(reinterpret_cast<char *>(a.ypeer[p]) - reinterpret_cast<char *>(a.Y));
    }
    if (G.total_tiles == 0) return GGB_OK;
#define GGB_GROUPED_Q(CAPV, GV) \
    switch (type) { \
    case GGML_TYPE_Q4_0: return bn == 256 ? launch_grouped<GGML_TYPE_Q4_0, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q4_0, CAPV, 128>(GV, s); \
    case GGML_TYPE_Q4_1: return bn == 256 ? launch_grouped<GGML_TYPE_Q4_1, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q4_1, CAPV, 128>(GV, s); \
    case GGML_TYPE_Q4_2: return bn == 256 ? launch_grouped<GGML_TYPE_Q4_2, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q4_2, CAPV, 128>(GV, s); \
    case GGML_TYPE_Q8_0: return bn == 256 ? launch_grouped<GGML_TYPE_Q8_0, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q8_0, CAPV, 128>(GV, s); \
    case GGML_TYPE_Q5_0: return bn == 256 ? launch_grouped<GGML_TYPE_Q5_0, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q5_0, CAPV, 128>(GV, s); \
    default:             return bn == 256 ? launch_grouped<GGML_TYPE_Q5_1, CAPV, 256>(GV, s) : launch_grouped<GGML_TYPE_Q5_1, CAPV, 128>(GV, s); \
    }
    if (count <= SMALL_GROUP) {
        static thread_local GemmGroupT<SMALL_GROUP> S;
        S.n_nodes = G.n_nodes; S.total_tiles = G.total_tiles; S.wait_w = G.wait_w;
        for (int i = 0; i < count; i++) S.node[i] = G.node[i];
        if (type == GGML_TYPE_F16) return launch_grouped<GGML_TYPE_F16, SMALL_GROUP>(S, s);
        GGB_GROUPED_Q(SMALL_GROUP, S)
    }
    if (type == GGML_TYPE_F16) return launch_grouped<GGML_TYPE_F16, GGB_GEMM_GROUP_NODES>(G, s);
    GGB_GROUPED_Q(GGB_GEMM_GROUP_NODES, G)
#undef GGB_GROUPED_Q
}

} // namespace ggb
