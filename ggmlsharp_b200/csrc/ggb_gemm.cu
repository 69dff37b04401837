// Batched (N >= 16) mul_mat on the 5th-generation tensor cores: tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA, one warp-specialised CTA per 128 x BN output tile.
//
// Replaces ggml_compute_forward_mul_mat_q_f32 / _f16_f32 (Ggml.cs:6440-6712, 6180-6438) for prompt-sized
// batches.  dst[n][m] = sum_k W[m][k] * X[n][k]; both operands are K-major, so the MMA "M" dimension (the
// 128 TMEM lanes) is a tile of weight rows and the MMA "N" dimension (TMEM columns) a tile of activation
// rows; the epilogue's 32x32b TMEM load then gives each thread one m and 32 consecutive n, and a warp's store
// of one n is 32 consecutive floats of dst -- coalesced with no transpose.
//
//   warp 0       TMA producer: raw quant blocks (2-D byte tensor map, box = 128 rows x 4 blocks) or, for F16
//                weights, swizzled fp16 tiles; plus the activation tiles (fp16, swizzle-128B)
//   warp 1       MMA issuer: one elected lane, 8 x tcgen05.mma (K = 16 each) per 128-wide K step
//   warp 2       TMEM allocator
//   warps 4..19  dequant: each thread expands one 32-weight block per K step into the swizzle-128B K-major A
//                tile (nibble -> fp16 with the 0x6400 / 0x5400 magic-number trick, exact q-8, then * d),
//                fence.proxy.async, arrive; warps 4..7 are also the epilogue (tcgen05.ld -> global)
//
// Numerics: activations are quantized exactly as quantize_row_q8_0/q8_1 do (same d, same quants) and enter
// the MMA as fp16(d1*q); weights as fp16(d0*(q-8)) (Q4_0), fp16(d0*(q-8) + (m0+8*d0)) (Q4_1) or the stored fp16 (F16);
// products are exact in the tensor core and accumulate in fp32.  Measured rel-L2 vs the CPU oracle is in
// tests/test_gpu_gemm.py (bound 1e-3).  The magic-number unpack yields the 8 weights of a 32-bit word in the
// order 0,4,1,5,2,6,3,7; the activation kernel stores K in the same order, so no re-pairing is needed.
#include "ggb_tc.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace ggb {

namespace {

using namespace tc;

// Row split: the epilogue stores each result into this rank's dst AND the same element of every peer's dst
// (CUDA-IPC mapped pointers); a warp store is 128 contiguous bytes, a full NVLink write packet.
struct GemmPeers { int n; float *y[7]; };

// F16 weights: both operands are fp16 tiles fetched by TMA (swizzle-128B), A from shared memory.
// CG = CTAs per tile: 2 -> tcgen05 cta_group::2, the pair computes 256 weight rows x BN; each CTA stages its own 128 rows of A
// and only HALF of the activation tile.  Two issuer warps (even / odd K steps, separate accumulators) as in k_gemm_q.
template <int BN, int CG>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_f16(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
           float *__restrict__ Y, long long ldy, int M, int N, int K, const __grid_constant__ GemmPeers peers,
           const int *__restrict__ ew, const int *__restrict__ ex, int wait_w)
{
    // ew / ex: power-of-two row exponents of an EXPANDED quantized weight matrix and of its d*q activations (ggb_internal.h:
    // launch_weight_rowexp); both null for true F16 weights, whose operands the reference itself rounds to Half.
    constexpr int BNL = BN / CG;
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNL * BK * 2;
    constexpr int STAGES = CG == 2 ? 4 : 3, NISSUE = 2;
    constexpr int TMEM_COLS = NISSUE * BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                          // STAGES x A_BYTES (1024-aligned)
    uint8_t *sB = sA + STAGES * A_BYTES;                         // STAGES x B_BYTES
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + STAGES * B_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    constexpr int FULL = 0, EMPTY = 4, ACC_FULL = 8, NBARS = 9;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    auto BAR = [&](int i) { return bar0 + 8 * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN + (int)rank * BNL;
    const int ksteps = (K + BK - 1) / BK;
    auto LBAR = [&](int i) { return CG == 2 ? mapa_u32(BAR(i), 0) : BAR(i); };

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(BAR(FULL + i), CG); mbar_init(BAR(EMPTY + i), 1); }
        mbar_init(BAR(ACC_FULL), NISSUE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 2) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next node's activation kernel may run beside this GEMM

    if (warp == 0) {
        // ===== TMA producer: this CTA's 128 weight rows and its half of the activation rows, per K step =====
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");       // activations come from the preceding kernel
            int sb = 0; uint32_t ph = 1;
            for (int ks = 0; ks < ksteps; ks++) {
                mbar_wait(BAR(EMPTY + sb), ph);
                const uint32_t full = LBAR(FULL + sb);
                constexpr uint32_t per_cta = B_BYTES + A_BYTES;
                if (CG == 2) {
                    if (leader) mbar_expect_tx(BAR(FULL + sb), CG * per_cta); else mbar_arrive_cluster(full);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES), &map_x, full, ks * BK, n0);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES + BNL * 128), &map_x, full, ks * BK + 64, n0);
                    tma_load_2d_cg2(smem_u32(sA + sb * A_BYTES), &map_w, full, ks * BK, m0);
                    tma_load_2d_cg2(smem_u32(sA + sb * A_BYTES + BM * 128), &map_w, full, ks * BK + 64, m0);
                } else {
                    mbar_expect_tx(BAR(FULL + sb), per_cta);
                    tma_load_2d(smem_u32(sB + sb * B_BYTES), &map_x, full, ks * BK, n0);
                    tma_load_2d(smem_u32(sB + sb * B_BYTES + BNL * 128), &map_x, full, ks * BK + 64, n0);
                    tma_load_2d(smem_u32(sA + sb * A_BYTES), &map_w, full, ks * BK, m0);
                    tma_load_2d(smem_u32(sA + sb * A_BYTES + BM * 128), &map_w, full, ks * BK + 64, m0);
                }
                if (++sb == STAGES) { sb = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (leader CTA): even / odd K steps, own accumulator each =====
        if (leader && lane == 0) {
            const int me = warp >> 1;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
            const uint32_t acc = tmem + (uint32_t)(me * BN);
            const uint64_t adesc0 = make_sdesc(smem_u32(sA)), bdesc0 = make_sdesc(smem_u32(sB));
            for (int ks = me; ks < ksteps; ks += NISSUE) {
                const int sb = ks % STAGES;
                mbar_wait(BAR(FULL + sb), (uint32_t)((ks / STAGES) & 1));
                tc_fence_after();
                const uint64_t ad0 = adesc0 + (uint64_t)((sb * A_BYTES) >> 4), bd0 = bdesc0 + (uint64_t)((sb * B_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    // sub-tile k/4 (64 K each), then 32 bytes per K=16 step inside the 128-byte swizzle row
                    const uint64_t ad = ad0 + (uint64_t)(((k >> 2) * (BM * 128) + (k & 3) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)(((k >> 2) * (BNL * 128) + (k & 3) * 32) >> 4);
                    if (CG == 2) tc_mma_f16_cg2(acc, ad, bd, idesc, (ks >= NISSUE) || k != 0);
                    else tc_mma_f16(acc, ad, bd, idesc, (ks >= NISSUE) || k != 0);
                }
                if (CG == 2) tc_commit_cg2(BAR(EMPTY + sb)); else tc_commit(BAR(EMPTY + sb));
            }
            if (CG == 2) tc_commit_cg2(BAR(ACC_FULL)); else tc_commit(BAR(ACC_FULL));
        }
    } else if (warp >= 4) {
        // ===== epilogue on all 16 warps: warp -> (TMEM lane quadrant warp%4, 32-column group) =====
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        const int nbase = blockIdx.y * BN;
        float fw = 1.0f, fxl = 1.0f;
        if (ew) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if (m < M) fw = exp2i(__ldcg(ew + m));
            fxl = exp2i(__ldcg(ex + nbase + ((warp - 4) >> 2) * 32 + lane));   // BN = 128: one 32-column group per warp; lane l holds column l's factor
        }
        mbar_wait(BAR(ACC_FULL), 0);
        tc_fence_after();
#pragma unroll 1
        for (int cb = (warp - 4) >> 2; cb < BN / 32; cb += NDQ_WARPS / 4) {
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32);
#pragma unroll 1
            for (int hc = 0; hc < 2; hc++) {                      // 16 columns at a time keeps both accumulators in 32 registers
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
                        const int n = nbase + cb * 32 + hc * 16 + c;
                        float r = ksteps > 1 ? __uint_as_float(v[c]) + __uint_as_float(u[c]) : __uint_as_float(v[c]);
                        if (ew) r = scale_pair(r, fw, __shfl_sync(0xffffffffu, fxl, hc * 16 + c));      // warp-uniform branch; every lane takes part in the shuffle
                        if (m < M && n < N) {
                            Y[(long long)n * ldy + m] = r;
                            for (int pp = 0; pp < peers.n; pp++) peers.y[pp][(long long)n * ldy + m] = r;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Q4_0 / Q4_1 weights: the dequantized A operand lives in TENSOR MEMORY, not shared memory.
//
// profiles/r01_gemm_trace.txt: with A staged in shared memory the K step took ~1150 cycles against a 512-cycle MMA floor,
// and the clock64 trace showed neither TMA latency nor instruction issue but shared-memory BANDWIDTH as the limit:
// per K step a CTA moved raw 10 KB in + 10 KB out, A 32 KB in (STS) + 32 KB out (tensor core), B 16 KB in + 16 KB out
// = 116 KB at 128 B/clk.  tcgen05.mma can take A from TMEM, and the dequant warps already own one weight row per lane --
// exactly TMEM's lane = row layout -- so they write fp16 pairs with tcgen05.st and 64 KB of shared traffic per step
// disappears (52 KB left = ~400 cycles, below the MMA floor).
//
// TMEM columns: [0, BN) fp32 accumulator; [BN + 64*g, +64) A stage g (128 K halfs = 64 x 32-bit), g = K step mod 4.
// Dequant warp (4 + 4*g + q): TMEM lane quadrant q = warp % 4 (hardware rule), K steps ks == g (mod 4).  Each thread reads its
// row's 80 / 96 raw bytes of the step with LDS.128 (lane stride 20 / 24 words), expands 4 blocks to 64 half2 registers
// in the K order 0,4,1,5,2,6,3,7 per 8 (the activation buffer uses the same order) and stores them to its lane.
// ------------------------------------------------------------------------------------------------

// RAW must be a multiple of the 4 dequant groups: then raw stage s is always consumed by group s % 4, which also consumed
// its previous phase, so a parity wait can never alias a stale phase (with RAW = 6 a group running ahead of the TMA read
// a stage one revolution early -- rel-L2 1.5e-2 on large Q4_1 shapes, caught by tests/test_gpu_gemm.py).
template <int TYPE, int CG> struct QStages {
    static constexpr int RAW = 8;
    static constexpr int B = 4;
};
static_assert(QStages<GGML_TYPE_Q4_1, 2>::RAW % 4 == 0, "raw ring must be a multiple of the dequant groups");

template <int TYPE, int BN, int CG>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_q(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
         float *__restrict__ Y, long long ldy, int M, int N, int K, long long *__restrict__ dbg, int dbg_flags,
         const __grid_constant__ GemmPeers peers, const int *__restrict__ ew, const int *__restrict__ ex, int wait_w)
{
    constexpr int RAW_ROW = RawRow<TYPE>::BYTES;
    constexpr int RAW_BYTES = BM * RAW_ROW;
    constexpr int BNL = BN / CG;
    constexpr int B_BYTES = BNL * BK * 2;
    constexpr int RAW_STAGES = QStages<TYPE, CG>::RAW, A_STAGES = 4, B_STAGES = A_STAGES;   // A (TMEM) and B (smem) stages share one barrier ring
    // two MMA issuer warps (even / odd K steps) accumulate into separate TMEM regions that the epilogue adds
    constexpr int TMEM_COLS = 512, NISSUE = 2, A_COL0 = NISSUE * BN;
    static_assert(NISSUE * BN + A_STAGES * 64 <= TMEM_COLS, "TMEM budget");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;
    uint8_t *sRaw = sB + B_STAGES * B_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRaw + RAW_STAGES * RAW_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    // FULL[g]: 4*CG dequant-warp arrivals + CG activation-producer arrivals (+ the TMA bytes); EMPTY[g]: one commit per K step
    constexpr int RAW_FULL = 0, RAW_EMPTY = 8, A_FULL = 16, A_EMPTY = 20, ACC_FULL = 24, NBARS = 25;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    auto BAR = [&](int i) { return bar0 + 8 * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *const tdbg = dbg ? dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 128 : nullptr;
    if (tdbg && threadIdx.x == 0) { tdbg[0] = clock64(); unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); tdbg[4] = (long long)gt; }
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN + (int)rank * BNL;
    const int ksteps = (K + BK - 1) / BK;
    auto LBAR = [&](int i) { return CG == 2 ? mapa_u32(BAR(i), 0) : BAR(i); };

    if (threadIdx.x == 0) {
        for (int i = 0; i < RAW_STAGES; i++) { mbar_init(BAR(RAW_FULL + i), 1); mbar_init(BAR(RAW_EMPTY + i), 4); }
        for (int i = 0; i < A_STAGES; i++) { mbar_init(BAR(A_FULL + i), 4 * CG + CG); mbar_init(BAR(A_EMPTY + i), 1); }
        mbar_init(BAR(ACC_FULL), NISSUE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 2) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next node's activation kernel may run beside this GEMM
    if (tdbg && threadIdx.x == 0) tdbg[1] = clock64();

    if (warp == 0) {
        // ===== TMA producer 1: raw quant blocks (own barrier ring, never blocked by the activation ring) =====
        if (lane == 0) {
            if (wait_w) asm volatile("griddepcontrol.wait;" ::: "memory");   // the weights were written by an earlier kernel of this batch (CPY / quantize)
            int s = 0; uint32_t ph_s = 1;
            for (int ks = 0; ks < ksteps; ks++) {
                mbar_wait(BAR(RAW_EMPTY + s), ph_s);
                mbar_expect_tx(BAR(RAW_FULL + s), RAW_BYTES);
                tma_load_2d(smem_u32(sRaw + s * RAW_BYTES), &map_w, BAR(RAW_FULL + s), RawRow<TYPE>::box_x(ks), m0);
                if (++s == RAW_STAGES) { s = 0; ph_s ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===== TMA producer 2: activation tiles (this CTA's half of the BN rows) =====
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");       // PDL: the activation kernel must have finished; weights never wait
            int sb = 0; uint32_t ph_b = 1;
            for (int ks = 0; ks < ksteps; ks++) {
                mbar_wait(BAR(A_EMPTY + sb), ph_b);
                if (CG == 2) {
                    const uint32_t full = LBAR(A_FULL + sb);
                    if (leader) mbar_expect_tx(BAR(A_FULL + sb), CG * B_BYTES); else mbar_arrive_cluster(full);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES), &map_x, full, ks * BK, n0);
                    tma_load_2d_cg2(smem_u32(sB + sb * B_BYTES + BNL * 128), &map_x, full, ks * BK + 64, n0);
                } else {
                    mbar_expect_tx(BAR(A_FULL + sb), B_BYTES);
                    tma_load_2d(smem_u32(sB + sb * B_BYTES), &map_x, BAR(A_FULL + sb), ks * BK, n0);
                    tma_load_2d(smem_u32(sB + sb * B_BYTES + BNL * 128), &map_x, BAR(A_FULL + sb), ks * BK + 64, n0);
                }
                if (++sb == B_STAGES) { sb = 0; ph_b ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers (leader CTA): A from TMEM, B from shared memory =====
        // benchmarks/micro/mma_rate.cu: tcgen05.mma issue is effectively synchronous for the issuing thread -- a lone
        // issuer that also waits on barriers and commits ran at ~125 cycles per 128x128x16 MMA against the 64-cycle
        // floor, two issuers with the same overhead at ~78.  So warps 1 and 3 take the even / odd K steps, each with its
        // own accumulator (columns [0,BN) and [BN,2BN)); descriptors are one 64-bit add off a per-stage base.
        if (leader && lane == 0) {
            const int me = warp >> 1;                                  // 0 or 1
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
            const uint32_t acc = tmem + (uint32_t)(me * BN);
            const uint64_t bdesc0 = make_sdesc(smem_u32(sB));
            for (int ks = me; ks < ksteps; ks += NISSUE) {
                const int g = ks & 3;
                mbar_wait(BAR(A_FULL + g), (uint32_t)((ks >> 2) & 1));       // A rows in TMEM (both CTAs) + both halves of the B tile
                tc_fence_after();
                const uint64_t bd0 = bdesc0 + (uint64_t)((g * B_BYTES) >> 4);
                const uint32_t a_base = tmem + A_COL0 + g * 64;
#pragma unroll
                for (int k = 0; k < BK / 16; k++)
                    tc_mma_f16_ts(acc, a_base + k * 8, bd0 + (uint64_t)(((k >> 2) * (BNL * 128) + (k & 3) * 32) >> 4), idesc, (ks >= NISSUE) || k != 0, CG == 2);
                if (CG == 2) tc_commit_cg2(BAR(A_EMPTY + g)); else tc_commit(BAR(A_EMPTY + g));     // frees TMEM stage g and B stage g
            }
            if (CG == 2) tc_commit_cg2(BAR(ACC_FULL)); else tc_commit(BAR(ACC_FULL));
        }
    } else if (warp >= 4) {
        // ===== dequant warps: raw blocks (shared) -> fp16 A rows (tensor memory) =====
        {
            const int q = warp & 3, g = (warp - 4) >> 2;
            const int r = q * 32 + lane;
            const uint32_t raw_row = smem_u32(sRaw) + (uint32_t)(r * RAW_ROW);
            const uint32_t a_full = LBAR(A_FULL + g);
            const uint32_t a_tmem = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_COL0 + g * 64);
            uint32_t mk_lo = 0x000F000Fu, mk_hi = 0x00F000F0u, mg_lo = 0x64006400u, mg_hi = 0x54005400u;
            asm volatile("" : "+r"(mk_lo), "+r"(mk_hi), "+r"(mg_lo), "+r"(mg_hi));
            uint32_t ph_a = 1;
            if (wait_w) asm volatile("griddepcontrol.wait;" ::: "memory");   // ew comes from a kernel launched earlier in this batch
            const float rs = exp2i(-((m0 + r) < M ? __ldcg(ew + m0 + r) : 0));   // this thread's weight row, pre-scaled by 2^-ew
            for (int ks = g; ks < ksteps; ks += 4) {
                const int s = ks % RAW_STAGES;
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 0] = clock64();
                mbar_wait(BAR(RAW_FULL + s), (uint32_t)((ks / RAW_STAGES) & 1));
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 1] = clock64();
                uint32_t w[RAW_ROW / 4] = {};
                if (!(dbg_flags & 2)) load_raw_row<TYPE>(raw_row + (uint32_t)(s * RAW_BYTES), ks, w);
                // NOTE: the raw stage is released only after the dequant below has CONSUMED these registers.  Arriving right after
                // issuing the LDS (data still in flight) let the next TMA overwrite the stage under the loads: intermittent
                // rel-L2 ~5e-3 on large Q4_1 shapes (benchmarks/q41_bisect.sh).
                mbar_wait(BAR(A_EMPTY + g), ph_a);
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 2] = clock64();
                tc_fence_after();
                if (!(dbg_flags & 1))
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {                         // two blocks (64 K = 32 columns) per TMEM store
                    uint32_t v[32];
#pragma unroll
                    for (int jb = 0; jb < 2; jb++) {
                        dequant_group<TYPE>(w, hb * 2 + jb, rs, mk_lo, mk_hi, mg_lo, mg_hi, &v[jb * 16]);
                    }
                    tmem_st_x32(a_tmem + (uint32_t)(hb * 32), v);
                }
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 3] = clock64();
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 4] = clock64();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_cluster(a_full); else mbar_arrive(BAR(A_FULL + g)); mbar_arrive(BAR(RAW_EMPTY + s)); }
                if (tdbg && ks < 32 && (threadIdx.x == 128)) tdbg[(threadIdx.x == 128 ? 48 : 96) + (ks >> 2) * 6 + 5] = clock64();
                ph_a ^= 1;
            }
        }
        {
            // epilogue on all 16 warps: warp -> (TMEM lane quadrant warp%4, 32-column group)
            const int q = warp & 3;
            const int m = m0 + q * 32 + lane;
            const int nbase = blockIdx.y * BN;
            // before the wait for the accumulators, so the loads hide behind the last MMAs: lane l's factor 2^ex of column l of this warp's group
            asm volatile("griddepcontrol.wait;" ::: "memory");         // ex was written by the activation kernel (long complete: the B tiles came from it)
            const float fw = exp2i(m < M ? __ldcg(ew + m) : 0);
            const float fxl = exp2i(__ldcg(ex + nbase + ((warp - 4) >> 2) * 32 + lane));   // BN = 128: one 32-column group per warp
            mbar_wait(BAR(ACC_FULL), 0);
            if (tdbg && threadIdx.x == 128) tdbg[2] = clock64();
            tc_fence_after();
#pragma unroll 1
            for (int cb = (warp - 4) >> 2; cb < BN / 32; cb += NDQ_WARPS / 4) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32);
#pragma unroll 1
                for (int hc = 0; hc < 2; hc++) {                      // 16 columns at a time keeps both accumulators in 32 registers
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
                            const int n = nbase + cb * 32 + hc * 16 + c;
                            float r = ksteps > 1 ? __uint_as_float(v[c]) + __uint_as_float(u[c]) : __uint_as_float(v[c]);   // even-K + odd-K partial sums
                            r = scale_pair(r, fw, __shfl_sync(0xffffffffu, fxl, hc * 16 + c));   // undo the operands' power-of-two pre-scaling (exact); every lane shuffles
                            if (m < M && n < N) {
                                Y[(long long)n * ldy + m] = r;
                                for (int pp = 0; pp < peers.n; pp++) peers.y[pp][(long long)n * ldy + m] = r;
                            }
                        }
                    }
                }
            }
        }
    }
    if (tdbg && threadIdx.x == 128) { tdbg[3] = clock64(); unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); tdbg[5] = (long long)gt; }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

} // namespace (reopened below)

// ---- host side: tensor maps ----

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_map_2d(CUtensorMap *map, CUtensorMapDataType dt, const void *base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                uint32_t box0, uint32_t box1, CUtensorMapSwizzle sw)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(GGB_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {dim0, dim1};
    const cuuint64_t strides[1] = {stride1_bytes};
    const cuuint32_t box[2] = {box0, box1};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(GGB_E_CUDA, "cuTensorMapEncodeTiled failed (%d) dims=%llu x %llu stride=%llu box=%u x %u",
                                            (int)r, (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)stride1_bytes, box0, box1);
    return GGB_OK;
}

namespace {

template <int BN, int CG>
int launch_f16(const GemmArgs &a, cudaStream_t s)
{
    constexpr int BNL = BN / CG;
    CUtensorMap mw, mx;
    int rc = make_map_2d(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.W, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.nb01, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.Xh, (uint64_t)a.K, (uint64_t)a.Npad, (uint64_t)a.K * 2, 64, BNL, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    constexpr int STAGES = CG == 2 ? 4 : 3;
    constexpr size_t smem = 1024 + STAGES * (size_t)(BM * BK * 2) + STAGES * (size_t)(BNL * BK * 2) + 64 * 8 + 16;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(k_gemm_f16<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
    const unsigned mt = (unsigned)((a.M + BM - 1) / BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mt + CG - 1) / CG * CG, (unsigned)((a.N + BN - 1) / BN));    // whole CTA pairs; a padding CTA sees only out-of-bounds (zero) rows
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // prologue overlaps the activation kernel
    at[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("GGB200_NO_PDL") != nullptr;
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 1 : 2;
    GemmPeers peers = {};
    peers.n = a.n_peers;
    for (int p = 0; p < a.n_peers; p++) peers.y[p] = a.ypeer[p];
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_f16<BN, CG>, mw, mx, a.Y, (long long)a.ldy, (int)a.M, (int)a.N, (int)a.K, peers, a.ew, a.ex, a.wait_w));
    count_launch();
    return GGB_OK;
}

template <int TYPE, int BN, int CG>
int launch_q(const GemmArgs &a, cudaStream_t s)
{
    if (!a.ew || !a.ex) return set_error(GGB_E_INVALID, "batched path: quantized weights need their row exponents (ew / ex)");
    constexpr int RAW_ROW = RawRow<TYPE>::BYTES;
    constexpr int BNL = BN / CG;
    CUtensorMap mw, mx;
    const uint64_t row_bytes = (uint64_t)(a.K / GGB_QK) * q32_bytes(TYPE);
    int rc = make_map_2d(&mw, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.W, row_bytes, (uint64_t)a.M, (uint64_t)a.nb01, RAW_ROW, BM, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.Xh, (uint64_t)a.K, (uint64_t)a.Npad, (uint64_t)a.K * 2, 64, BNL, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    constexpr size_t smem = 1024 + (size_t)QStages<TYPE, CG>::B * (BNL * BK * 2) + (size_t)QStages<TYPE, CG>::RAW * (BM * RAW_ROW) + 64 * 8 + 16;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { GGB_CUDA(cudaFuncSetAttribute(k_gemm_q<TYPE, BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
    const unsigned mt = (unsigned)((a.M + BM - 1) / BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mt + CG - 1) / CG * CG, (unsigned)((a.N + BN - 1) / BN));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // prologue + weight streaming overlap the activation kernel
    at[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("GGB200_NO_PDL") != nullptr;
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 1 : 2;
    static const int dbg_flags = [] { const char *e = getenv("GGB200_GEMM_DBG"); return e ? atoi(e) : 0; }();   // perf experiments only (wrong results)
    GemmPeers peers = {};
    peers.n = a.n_peers;
    for (int p = 0; p < a.n_peers; p++) peers.y[p] = a.ypeer[p];
    GGB_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_q<TYPE, BN, CG>, mw, mx, a.Y, (long long)a.ldy, (int)a.M, (int)a.N, (int)a.K, static_cast<long long *>(a.trace), dbg_flags, peers, a.ew, a.ex, a.wait_w));
    count_launch();
    return GGB_OK;
}

} // namespace

bool gemm_supported(int type, int64_t M, int64_t K, int64_t N, int64_t nb01, const void *W)
{
    // F32 weights stay on FFMA (1e-5 bar)
    if (type != GGML_TYPE_F16 && !is_q_weight(type)) return false;
    if (M <= 0 || N < 16 || K <= 0 || K % GGB_QK) return false;
    if ((reinterpret_cast<uintptr_t>(W) & 15) || (nb01 & 15)) return false;                            // TMA: 16-byte base and strides
    if (type == GGML_TYPE_F16) return K % 8 == 0;
    return K % 128 == 0;                                                                               // whole 4-block boxes
}

size_t gemm_workspace_bytes(int type, int64_t M, int64_t K, int64_t N)
{
    // fp16 activations; quantized weights add the two exponent arrays (ggb_internal.h: gemm_ws_ex_offset / gemm_ws_ew_offset)
    if (type == GGML_TYPE_F16) return gemm_ws_ex_offset(K, N);
    return align_up(gemm_ws_ew_offset(K, N) + (size_t)(M > 0 ? M : 0) * 4, 256);
}

int gemm_act_perm(int type) { return type == GGML_TYPE_F16 ? 0 : 1; }

int launch_gemm(const GemmArgs &a, void *ws, cudaStream_t s)
{
    (void)ws;
    if (a.n_peers < 0 || a.n_peers > 7) return set_error(GGB_E_INVALID, "batched path: n_peers=%d", a.n_peers);
    static const int cg = [] { const char *e = getenv("GGB200_GEMM_CG"); return e ? atoi(e) : 2; }();     // debugging: 1 = no CTA pairs
    switch (a.type) {
    case GGML_TYPE_Q4_0: return cg == 1 ? launch_q<GGML_TYPE_Q4_0, 128, 1>(a, s) : launch_q<GGML_TYPE_Q4_0, 128, 2>(a, s);
    case GGML_TYPE_Q4_1: return cg == 1 ? launch_q<GGML_TYPE_Q4_1, 128, 1>(a, s) : launch_q<GGML_TYPE_Q4_1, 128, 2>(a, s);
    case GGML_TYPE_Q4_2: return launch_q<GGML_TYPE_Q4_2, 128, 2>(a, s);
    case GGML_TYPE_Q5_1: return launch_q<GGML_TYPE_Q5_1, 128, 2>(a, s);
    case GGML_TYPE_Q8_0: return launch_q<GGML_TYPE_Q8_0, 128, 2>(a, s);
    case GGML_TYPE_Q5_0: return launch_q<GGML_TYPE_Q5_0, 128, 2>(a, s);
    case GGML_TYPE_F16: return cg == 1 ? launch_f16<128, 1>(a, s) : launch_f16<128, 2>(a, s);
    default: return set_error(GGB_E_UNSUPPORTED, "batched path: type %d", a.type);
    }
}

} // namespace ggb
