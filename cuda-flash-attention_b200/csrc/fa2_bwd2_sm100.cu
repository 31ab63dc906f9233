// fa2_bwd2_sm100.cu -- CTA-pair FlashAttention-2 backward for sm_100a at D = 128 (same maths as fa2_bwd_sm100.cu and
// the reference's flash_attention2_backward_kernel, kernels/f-attn2-backward.cu:33-339).
//
// A cluster of two CTAs (one SM pair) owns TWO neighbouring 128-row KV tiles of a (batch, head) slab -- CTA r the
// tile r -- and walks the Q tiles like the single-CTA kernel.  Every product is a tcgen05.mma.cta_group::2 issued by the
// leader CTA (rank 0), which lifts the two walls the single-CTA kernel sits on:
//   * operand fetch: a pair MMA reads per SM its own 128 rows of A and only HALF of B, so the 128x128x128 shares run at
//     ~550 cycles instead of 715 (tools/mma2_probe.cu);
//   * dQ reduce traffic: dQ = dS K contracts over the KV rows of BOTH tiles inside one M = 128 pair MMA (each CTA
//     receives 64 of the 128 query rows), so each Q step reduce-adds 32 KB per SM into L2 instead of 64 KB.
//     S^T  = K Q^T    M = 256: A = own K tile (K-major),       B = Q rows [64 r, 64 r + 64) (K-major)      -> TMEM [384,512)
//     dP^T = V dO^T   M = 256: A = own V tile,                 B = dO rows [64 r, +64)                     -> TMEM [256,384)
//     dV  += P^T dO   M = 256: A = P^T from TMEM [384,448),    B = dO columns [64 r, +64) (MN-major)        -> TMEM [128,256)
//     dK  += dS^T Q   M = 256: A = dS^T from TMEM (inside dP), B = Q columns [64 r, +64) (MN-major)         -> TMEM [0,128)
//     dQ   = dS K     M = 128: A = dS^T atoms of BOTH tiles, query columns [64 r, +64), read MN-major;
//                              B = K columns [64 r, +64) of both tiles (MN-major)                          -> TMEM [448,512)
// TMEM is full, so dQ shares columns with S^T: dQ(G-1) is issued once the compute warps have copied S^T(G) out and its
// drain ends long before S^T(G+1) is issued -- the ~900-cycle DSMEM copy, the dQ MMA and its drain all run in the shadow
// of the P phase instead of sitting between dS(G-1) and dP(G).
// The dQ MMA needs, in CTA r, the dS^T atom "query half r" of the PEER's tile: the compute warpgroup that produces the
// other half writes it to a send buffer and ships it into the peer's shared memory with one DSMEM bulk copy
// (cp.async.bulk.shared::cluster, completion on an mbarrier of the receiver; tools/mma2b_probe.cu).  dQ lands in the
// "2x2" TMEM layout (lanes 0-63: head-dim columns 0..63, lanes 64-127: columns 64..127 of query rows 64 r + lane % 64).
// Barriers the leader's MMA warp waits on ("full", p / dS ready, dQ drained) live in the leader and are credited /
// arrived on by both CTAs through the cluster window; every tcgen05.commit is multicast to both CTAs.
//
// Warp roles (16 warps): 0-3 / 4-7 compute (P^T, dS^T; warpgroup h owns query columns [64 h, +64)), 8-11 dQ drain,
// 12 MMA issuer (leader only), 13 TMA producer (also stages LSE / D_i), 14 forwards "peer atom received" from the
// follower to the leader, 15 issues the DSMEM copy of the dS^T send atom.  512 x 128 registers at launch; setmaxnreg: compute 160, drain 104, rest 56.
#include <cstdlib>

#include "fa2_common.h"
#include "ptx.cuh"

#ifndef FA2_BWD_PAIR_DEFAULT
#define FA2_BWD_PAIR_DEFAULT 1
#endif
#ifndef FA2_BWD2_POLY_DEFAULT
#define FA2_BWD2_POLY_DEFAULT 1
#endif

namespace fa2 {
namespace {

constexpr int BT = 128;
constexpr int DP = 128;
constexpr int ATOM = BT * 128;          // [128 rows][64 x 16-bit], 128B-swizzled: 16 KB
constexpr int HATOM = 64 * 128;         // [64 rows][64 x 16-bit]: 8 KB
constexpr int NUM_THREADS = 512;
constexpr int D_WARP0 = 8, MMA_WARP = 12, TMA_WARP = 13, FWD_WARP = 14;

// Debug timeline (only with -DFA2_TIMELINE): lane 0 of a role of the LEADER CTA stamps clock64 into slot `slot` of step i
// of work item 0 (tools/timeline_bwd.py).
#ifdef FA2_TIMELINE
#define TL(slot) do { if (p.timeline && w == 0 && rank == 0 && lane == 0 && i < 32) p.timeline[i * 32 + (slot)] = clock64(); } while (0)
#define TLC(k) do { if (p.timeline && rank == 0) p.timeline[1024 + 8 * w + (k)] = clock64(); } while (0)
// the same for the follower CTA (its clock is another SM's: only differences within the follower are meaningful,
// plus the offset to the leader that the shared barriers imply)
#define TLF(slot) do { if (p.timeline && w == 0 && rank == 1 && lane == 0 && i < 32) p.timeline[i * 32 + (slot)] = clock64(); } while (0)
#else
#define TL(slot) do { } while (0)
#define TLC(k) do { } while (0)
#define TLF(slot) do { } while (0)
#endif

struct L {
    static constexpr int OFF_K = 0;                        // own K tile, 2 atoms (A of S^T)
    static constexpr int OFF_V = OFF_K + 2 * ATOM;         // own V tile (A of dP^T)
    static constexpr int OFF_KB = OFF_V + 2 * ATOM;        // K columns [64 r, +64) of tile 0 / tile 1 (B of dQ)
    static constexpr int OFF_QN = OFF_KB + 2 * ATOM;       // Q rows [64 r, +64), 2 half atoms (B of S^T)
    static constexpr int OFF_QD = OFF_QN + 2 * HATOM;      // Q columns [64 r, +64), all rows (B of dK)
    static constexpr int OFF_DON = OFF_QD + ATOM;          // dO rows [64 r, +64) (B of dP^T)
    static constexpr int OFF_DOD = OFF_DON + 2 * HATOM;    // dO columns [64 r, +64) (B of dV)
    static constexpr int OFF_DQA = OFF_DOD + ATOM;         // dS^T atoms "query half r" of tile 0 / tile 1 (A of dQ)
    static constexpr int OFF_SEND = OFF_DQA + 2 * ATOM;    // own dS^T atom "query half 1 - r", shipped to the peer
    static constexpr int OFF_DQS = OFF_SEND + ATOM;        // dQ staging: 2 boxes [64 rows][32 fp32]
    static constexpr int OFF_LSE = OFF_DQS + ATOM;         // 2 stages x 128 fp32
    static constexpr int OFF_DELTA = OFF_LSE + 2 * BT * 4;
    static constexpr int OFF_BAR = OFF_DELTA + 2 * BT * 4;
    static constexpr int NUM_BARS = 24;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int BYTES = OFF_TMEM_PTR + 16;
};
static_assert(L::BYTES <= 232448, "shared memory budget of one CTA");

// POLY: which of the exponentials of P^T go through the FMA-pipe polynomial (ex2_poly2) instead of MUFU.EX2 -- one bit
// per group of 4 columns inside a 32-column chunk, low byte for the first pair of the group, high byte for the second
// (the P phase is MUFU-bound: 128 x 128 exponentials per step at 16 per clock are 1024 cycles).
template <bool BF16, unsigned POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
fa2_bwd2_kernel(const __grid_constant__ BwdParams p) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t COL_DK = 0, COL_DV = 128, COL_DP = 256, COL_S = 384, COL_DQ = 448;
    constexpr int KSTEPS_D = DP / 16, KSTEPS_T = BT / 16;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* kv_full = bars + 0;       // [leader] K, V, KB of both CTAs landed
    uint64_t* kv_empty = bars + 1;
    uint64_t* qn_full = bars + 2;       // [leader]
    uint64_t* qn_empty = bars + 3;
    uint64_t* qd_full = bars + 4;       // [leader]
    uint64_t* qd_empty = bars + 5;
    uint64_t* don_full = bars + 6;      // [leader]
    uint64_t* don_empty = bars + 7;
    uint64_t* dod_full = bars + 8;      // [leader]
    uint64_t* dod_empty = bars + 9;
    // (bars 10, 11 unused)
    uint64_t* s_full = bars + 12;
    uint64_t* p_full = bars + 13;       // [leader] 16 compute warps
    uint64_t* dp_full = bars + 14;
    uint64_t* ds_full = bars + 15;      // [leader] 16 compute warps
    uint64_t* ds_empty = bars + 16;
    uint64_t* dqa_recv = bars + 17;     // the peer's dS^T atom has landed here (complete_tx)
    uint64_t* dqa_fwd = bars + 18;      // [leader] ... and in the follower
    uint64_t* dq_full = bars + 19;
    uint64_t* dq_empty = bars + 20;     // [leader] 8 drain warps
    uint64_t* dkdv_full = bars + 21;
    uint64_t* epi_issued = bars + 22;
    uint64_t* s_read = bars + 23;       // [leader] 16 compute warps have copied S^T out of TMEM
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
    float* lse_s = reinterpret_cast<float*>(smem + L::OFF_LSE);
    float* delta_s = reinterpret_cast<float*>(smem + L::OFF_DELTA);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int unit = blockIdx.x >> 1, n_units = gridDim.x >> 1;

    // work item = (bh, pair of KV tiles); every item has n_q steps (Q tiles)
    const int n_q = (p.S_q + BT - 1) / BT;
    const int n_kvt = (p.S_kv + BT - 1) / BT;
    const int n_pairs = (n_kvt + 1) >> 1;
    const int n_work = p.BH * n_pairs;
    auto q_row_at = [&](int u, int i) { return ((i + u) % n_q) * BT; };     // rotated per pair: concurrent pairs of a slab hit different dQ tiles

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v); tma_prefetch_desc(&p.tm_do);
        tma_prefetch_desc(&p.tm_q64); tma_prefetch_desc(&p.tm_do64); tma_prefetch_desc(&p.tm_dq64);
        tma_prefetch_desc(&p.tm_dk); tma_prefetch_desc(&p.tm_dv);
    }
    if (warp == MMA_WARP) {
        if (lane == 0) {
            if (smem_u32(smem) & 1023u) __trap();
            for (int i = 0; i < L::NUM_BARS; ++i) mbar_init(&bars[i], 1);
            mbar_init(s_full, 2);             // tcgen05.commit (multicast) + this CTA's producer warp (LSE / D_i staged)
            mbar_init(p_full, 16);
            mbar_init(ds_full, 16);
            mbar_init(s_read, 16);
            mbar_init(dq_empty, 8);
            mbar_init(epi_issued, 2);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair(tmem_holder, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == TMA_WARP) {
        // ------------------------------------------------------------------ producer (both CTAs)
        setmaxnreg_dec<56>();
        const uint32_t kv_full_l = cluster_map(kv_full, 0), qn_full_l = cluster_map(qn_full, 0), qd_full_l = cluster_map(qd_full, 0);
        const uint32_t don_full_l = cluster_map(don_full, 0), dod_full_l = cluster_map(dod_full, 0);
        const int c64 = 64 * static_cast<int>(rank);
        auto stage_side = [&](int G, int bh, int q_row) {      // LSE (log2 domain) and D_i of a Q tile -> stage G & 1
            const float delta_mul = p.range != nullptr ? ldg_scalar_volatile(p.range + kDeltaMul) : 1.0f;
#pragma unroll
            for (int r = 0; r < BT / 32; ++r) {
                const int row = q_row + r * 32 + lane;
                const bool ok = row < p.S_q;
                const size_t g = static_cast<size_t>(bh) * p.S_q + (ok ? row : 0);
                lse_s[(G & 1) * BT + r * 32 + lane] = ok ? __ldg(p.lse_log2 + g) : INFINITY;     // rows past S: P = 0
                delta_s[(G & 1) * BT + r * 32 + lane] = ok ? __ldg(p.delta + g) * delta_mul : 0.0f;
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(s_full);              // s_full(G) = S^T(G) complete (commit) + its LSE / D_i staged
            __syncwarp();
        };
        int it = 0;
        for (int w = unit; w < n_work; w += n_units, ++it) {
            const int bh = w / n_pairs, u = w % n_pairs;
            const int kv_pair_row0 = u * 2 * BT, kv_row0 = kv_pair_row0 + static_cast<int>(rank) * BT;
            const int G0 = it * n_q;
            mbar_wait(kv_empty, (it & 1) ^ 1);                  // the previous item's MMAs are done with K / V / KB
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(kv_full, 2 * 6 * ATOM);
                for (int a = 0; a < 2; ++a) {
                    tma_load_3d_pair(smem + L::OFF_K + a * ATOM, &p.tm_k, kv_full_l, a * 64, kv_row0, bh);
                    tma_load_3d_pair(smem + L::OFF_V + a * ATOM, &p.tm_v, kv_full_l, a * 64, kv_row0, bh);
                    tma_load_3d_pair(smem + L::OFF_KB + a * ATOM, &p.tm_k, kv_full_l, c64, kv_pair_row0 + a * BT, bh);
                }
            }
            __syncwarp();
            // Loads are issued in the order their buffers come free within a step (each operand has ONE buffer):
            //   dOD(G) after dV(G-1) | QN(G+1) after S(G) | QD(G) after dK(G-1), then LSE / D_i of step G+1 | dON(G+1) after dP(G)
            auto load_qn = [&](int G, int q_row) {
                mbar_wait(qn_empty, (G & 1) ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(qn_full, 2 * 2 * HATOM);
                    for (int a = 0; a < 2; ++a)
                        tma_load_3d_pair(smem + L::OFF_QN + a * HATOM, &p.tm_q64, qn_full_l, a * 64, q_row + c64, bh);
                }
                __syncwarp();
            };
            auto load_don = [&](int G, int q_row) {
                mbar_wait(don_empty, (G & 1) ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(don_full, 2 * 2 * HATOM);
                    for (int a = 0; a < 2; ++a)
                        tma_load_3d_pair(smem + L::OFF_DON + a * HATOM, &p.tm_do64, don_full_l, a * 64, q_row + c64, bh);
                }
                __syncwarp();
            };
            // (stage G0 & 1 was last read in step G0 - 2, and this warp has waited for qd_empty of that step)
            stage_side(G0, bh, q_row_at(u, 0));
            load_qn(G0, q_row_at(u, 0));
            load_don(G0, q_row_at(u, 0));
            for (int i = 0; i < n_q; ++i) {
                const int G = G0 + i, q_row = q_row_at(u, i);
                const uint32_t ph = (G & 1) ^ 1;
                mbar_wait(dod_empty, ph);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(dod_full, 2 * ATOM);
                    tma_load_3d_pair(smem + L::OFF_DOD, &p.tm_do, dod_full_l, c64, q_row, bh);
                }
                __syncwarp();
                if (i + 1 < n_q) load_qn(G + 1, q_row_at(u, i + 1));
                mbar_wait(qd_empty, ph);                         // dK(G-1) is done: the compute warps are past step G-1
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(qd_full, 2 * ATOM);
                    tma_load_3d_pair(smem + L::OFF_QD, &p.tm_q, qd_full_l, c64, q_row, bh);
                }
                __syncwarp();
                if (i + 1 < n_q) {
                    stage_side(G + 1, bh, q_row_at(u, i + 1));   // its stage was last read in step G-1
                    load_don(G + 1, q_row_at(u, i + 1));
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        setmaxnreg_dec<56>();
        if (rank == 0) {
            const uint32_t id_ss = umma_idesc_f16(2 * BT, BT, 0, 0, BF16 ? 1 : 0);      // S^T, dP^T : M = 256, N = 128
            const uint32_t id_kmn = umma_idesc_f16(2 * BT, DP, 0, 1, BF16 ? 1 : 0);     // dV, dK    : A K-major (TMEM), B MN-major
            const uint32_t id_dq = umma_idesc_f16(BT, DP, 1, 1, BF16 ? 1 : 0);          // dQ        : M = 128 over the pair, A / B MN-major
            const uint32_t hi = umma_desc_hi(1024);
            const uint32_t k_k = umma_desc_lo(smem_u32(smem + L::OFF_K), 16), v_k = umma_desc_lo(smem_u32(smem + L::OFF_V), 16);
            const uint32_t qn_k = umma_desc_lo(smem_u32(smem + L::OFF_QN), 16), don_k = umma_desc_lo(smem_u32(smem + L::OFF_DON), 16);
            const uint32_t qd_mn = umma_desc_lo(smem_u32(smem + L::OFF_QD), ATOM), dod_mn = umma_desc_lo(smem_u32(smem + L::OFF_DOD), ATOM);
            const uint32_t dqa_mn = umma_desc_lo(smem_u32(smem + L::OFF_DQA), ATOM), kb_mn = umma_desc_lo(smem_u32(smem + L::OFF_KB), ATOM);
            const uint32_t tS = tmem_base + COL_S, tDP = tmem_base + COL_DP, tDQ = tmem_base + COL_DQ;
            const uint32_t tDK = tmem_base + COL_DK, tDV = tmem_base + COL_DV;

            auto issue_s = [&]() {
                static_for<KSTEPS_D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_pair_ss_off<koff_kmajor(k, ATOM), koff_kmajor(k, HATOM)>(tS, k_k, qn_k, hi, id_ss, k > 0);
                });
            };
            auto issue_dp = [&]() {
                static_for<KSTEPS_D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_pair_ss_off<koff_kmajor(k, ATOM), koff_kmajor(k, HATOM)>(tDP, v_k, don_k, hi, id_ss, k > 0);
                });
            };
            auto issue_dv = [&](bool first) {       // P^T: the first 64 columns of the S region
                static_for<KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_pair_ts_off<k * 8, koff_mnmajor(k)>(tDV, tS, dod_mn, hi, id_kmn, (!first || k > 0) ? 1u : 0u);
                });
            };
            auto issue_dk = [&](bool first) {       // dS^T: two 32-column runs of the dP region (one per compute warpgroup)
                static_for<KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_pair_ts_off<(k >> 2) * 64 + (k & 3) * 8, koff_mnmajor(k)>(tDK, tDP, qd_mn, hi, id_kmn, (!first || k > 0) ? 1u : 0u);
                });
            };
            auto issue_dq = [&]() {                 // contraction over the 256 KV rows of the pair: tile 0 atoms, then tile 1
                static_for<2 * KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    constexpr uint32_t off = (k >> 3) * (ATOM >> 4) + koff_mnmajor(k & 7);
                    umma_pair_ss_off<off, off>(tDQ, dqa_mn, kb_mn, hi, id_dq, k > 0);
                });
            };
            // Issue order per step G:   S(G)  |  dK(G-1)  |  dP(G)  |  dQ(G-1)  |  dV(G)
            // dK(G-1) consumes dS^T(G-1) before dP(G) overwrites it; dQ(G-1) lands in the upper half of the S region once
            // the compute warps have copied S^T(G) out (s_read) and is drained before S^T(G+1) is issued.
            auto issue_tail_dk = [&](int w, int i, int Gp, bool first) {
                (void)w; (void)i;
                mbar_wait_spin(ds_full, Gp & 1);
                mbar_wait_spin(qd_full, Gp & 1);
                tc_fence_after();
                TL(1);
                if (elect_one()) {
                    issue_dk(first);
                    umma_commit_pair(qd_empty);
                    mbar_expect_tx(dqa_recv, ATOM);            // the follower's atom for this CTA's half of dQ(Gp)
                }
                __syncwarp();
                TL(6);
            };
            auto issue_tail_dq = [&](int w, int i, int Gp, bool last) {
                (void)w; (void)i;
                mbar_wait_spin(dqa_recv, Gp & 1);
                mbar_wait_spin(dqa_fwd, Gp & 1);               // ... and this CTA's atom has landed in the follower
                tc_fence_after();
                TL(5);
                if (elect_one()) {
                    issue_dq();
                    umma_commit_pair(dq_full);
                    umma_commit_pair(ds_empty);
                    if (last) {
                        umma_commit_pair(dkdv_full);
                        umma_commit_pair(kv_empty);
                    }
                }
                __syncwarp();
            };

            int it = 0;
            int dq_issued = -1, dq_waited = -1;                // last step whose dQ was issued / whose drain was awaited
            for (int w = unit; w < n_work; w += n_units, ++it) {
                const int G0 = it * n_q;
                mbar_wait_spin(kv_full, it & 1);
                for (int i = 0; i < n_q; ++i) {
                    const int G = G0 + i;
                    mbar_wait_spin(qn_full, G & 1);
                    if (dq_waited < dq_issued) {                // the last dQ issued has left TMEM: S^T may overwrite its columns
                        mbar_wait_spin(dq_empty, dq_issued & 1);
                        dq_waited = dq_issued;
                        TL(2);
                    }
                    tc_fence_after();
                    TL(0);
                    if (elect_one()) {
                        issue_s();                              // dV(G-1) was issued before: P^T(G-1) is consumed
                        umma_commit_pair(s_full);
                        umma_commit_pair(qn_empty);
                    }
                    __syncwarp();
                    if (i > 0) issue_tail_dk(w, i, G - 1, i == 1);
                    mbar_wait_spin(don_full, G & 1);
                    tc_fence_after();
                    TL(3);
                    if (elect_one()) {
                        issue_dp();
                        umma_commit_pair(dp_full);
                        umma_commit_pair(don_empty);
                    }
                    __syncwarp();
                    mbar_wait_spin(s_read, G & 1);              // (every step: keeps the barrier's phases in step)
                    if (i > 0) {
                        issue_tail_dq(w, i, G - 1, false);
                        dq_issued = G - 1;
                    }
                    mbar_wait_spin(p_full, G & 1);
                    mbar_wait_spin(dod_full, G & 1);
                    tc_fence_after();
                    TL(4);
                    if (elect_one()) {
                        issue_dv(i == 0);
                        umma_commit_pair(dod_empty);
                    }
                    __syncwarp();
                }
                issue_tail_dk(w, n_q, G0 + n_q - 1, n_q == 1);
                issue_tail_dq(w, n_q, G0 + n_q - 1, true);
                dq_issued = G0 + n_q - 1;
            }
        }
    } else if (warp == FWD_WARP) {
        // ------------------------------------------------------------------ follower: "the leader's atom has landed here"
        setmaxnreg_dec<56>();
        if (rank == 1) {
            const uint32_t fwd_l = cluster_map(dqa_fwd, 0);
            int it = 0;
            for (int w = unit; w < n_work; w += n_units, ++it) {
                for (int i = 0; i < n_q; ++i) {
                    const int G = it * n_q + i;
                    if (elect_one()) mbar_expect_tx(dqa_recv, ATOM);
                    __syncwarp();
                    mbar_wait(dqa_recv, G & 1);
                    if (elect_one()) mbar_arrive_cluster(fwd_l);
                    __syncwarp();
                }
            }
        }
    } else if (warp < D_WARP0) {
        // ------------------------------------------------------------------ compute warps: P^T and dS^T
        setmaxnreg_inc<160>();
        const int h = warp >> 2;                               // which 64 Q-columns of the tile
        const int n = (warp & 3) * 32 + lane;                  // kv row within the own tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + COL_S + h * 64;      // this warpgroup's S^T columns
        const uint32_t tPT = tmem_base + lane_addr + COL_S + h * 32;     // its P^T (16 bit): [384,448) holds both warpgroups'
        const uint32_t tDP = tmem_base + lane_addr + COL_DP + h * 64;    // its dP^T columns, then (first 32) its dS^T
        const float c2 = p.range != nullptr ? __ldg(p.range + kC2) : p.scale_log2;
        const float2 c2v = make_float2(c2, c2);
        const bool keep_local = static_cast<uint32_t>(h) == rank;            // this warpgroup's dS^T atom feeds this CTA's half of dQ
        uint8_t* ds_atom = keep_local ? smem + L::OFF_DQA + rank * ATOM : smem + L::OFF_SEND;
        const uint32_t p_full_l = cluster_map(p_full, 0), ds_full_l = cluster_map(ds_full, 0), s_read_l = cluster_map(s_read, 0);
        const bool issuer = ((warp & 3) == 0) && lane == 0;    // owns this warpgroup's dK / dV store groups
        const uint32_t ep_bar = 5 + 2 * h;                      // named barriers private to this warpgroup

        int it = 0;
        for (int w = unit; w < n_work; w += n_units, ++it) {
        const int bh = w / n_pairs, u = w % n_pairs;
        const int kv_row0 = u * 2 * BT + static_cast<int>(rank) * BT;
        const int G0 = it * n_q;
        const bool ragged_kv = kv_row0 + BT > p.S_kv;
        const uint32_t kv_keep = (kv_row0 + n < p.S_kv) ? 0xffffffffu : 0u;   // padded KV lane: P = dS = 0
#ifdef FA2_TIMELINE
        if (threadIdx.x == 0 && p.timeline && rank == 0) {
            uint32_t smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            p.timeline[1024 + 8 * w + 4] = smid;
            TLC(0);
        }
#endif
        for (int i = 0; i < n_q; ++i) {
            const int G = G0 + i;
            const float4* lse_t = reinterpret_cast<const float4*>(lse_s + (G & 1) * BT + h * 64);
            const float4* dl_t = reinterpret_cast<const float4*>(delta_s + (G & 1) * BT + h * 64);
            mbar_wait(s_full, G & 1);                           // S^T(G) complete and its LSE / D_i staging visible
            tc_fence_after();
            if (warp == 0) { TL(8); TLF(20); }
            if (warp == 4) { TL(24); TLF(26); }
            if (i == 0 && threadIdx.x == 0) TLC(1);
            uint32_t pk[32];                                    // P^T row (64 values) rounded to 16 bit
            {
                uint32_t sr[2][32];
                tmem_ld32(tS, sr[0]);
                tmem_ld32(tS + 32, sr[1]);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (rank == 0) mbar_arrive(s_read); else mbar_arrive_cluster(s_read_l); }   // dQ(G-1) may take S^T's upper columns
                named_bar_sync(9, 256);     // both warpgroups have copied S^T out: P^T of warpgroup 1 goes into warpgroup 0's columns
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float4 l4[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) l4[c] = lse_t[sub * 8 + hf * 4 + c];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int cc = hf * 4 + c;
                            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[sub][4 * cc]), __uint_as_float(sr[sub][4 * cc + 1])), c2v,
                                                         make_float2(-l4[c].x, -l4[c].y));
                            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[sub][4 * cc + 2]), __uint_as_float(sr[sub][4 * cc + 3])), c2v,
                                                         make_float2(-l4[c].z, -l4[c].w));
                            const float2 ea = ((POLY >> cc) & 1u) ? ex2_poly2(x0) : make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
                            const float2 eb = ((POLY >> (8 + cc)) & 1u) ? ex2_poly2(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                            pk[sub * 16 + 2 * cc] = BF16 ? pack_bf16x2(ea.x, ea.y) : pack_half2(ea.x, ea.y);
                            pk[sub * 16 + 2 * cc + 1] = BF16 ? pack_bf16x2(eb.x, eb.y) : pack_half2(eb.x, eb.y);
                        }
                    }
                }
            }
            if (ragged_kv) {
#pragma unroll
                for (int x = 0; x < 32; ++x) pk[x] &= kv_keep;
            }
            tmem_st32(tPT, pk);                                 // over S^T columns both warpgroups have consumed
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(p_full); else mbar_arrive_cluster(p_full_l); }
            if (warp == 0) { TL(9); TLF(21); }
            if (warp == 4) { TL(25); TLF(27); }

            mbar_wait(dp_full, G & 1);
            tc_fence_after();
            if (warp == 0) { TL(10); TLF(22); }
            uint32_t dr[2][32];
            tmem_ld32(tDP, dr[0]);
            tmem_ld32(tDP + 32, dr[1]);
            tmem_wait_ld();
            tc_fence_before();
            if (warp == 0) TL(12);
            // dS^T = P^T o (dP^T - D_i): the difference in fp32, the product in packed 16-bit, kept in registers until
            // the previous step's dQ is done with the shared-memory atoms
            uint32_t dsp[2][16];
            auto ds_math = [&](auto subc) {                     // 32 of this warpgroup's 64 query columns
                constexpr int sub = decltype(subc)::value;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float4 d4[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) d4[c] = dl_t[sub * 8 + hf * 4 + c];
#pragma unroll
                    for (int c8l = 0; c8l < 2; ++c8l) {        // 8 columns -> one 16-byte chunk of dS^T
                        const int c8 = hf * 2 + c8l;
#pragma unroll
                        for (int q2 = 0; q2 < 2; ++q2) {
                            const float4 dd = d4[c8l * 2 + q2];
                            const int c = c8 * 8 + q2 * 4;
                            const float2 t0 = __fadd2_rn(make_float2(__uint_as_float(dr[sub][c]), __uint_as_float(dr[sub][c + 1])), make_float2(-dd.x, -dd.y));
                            const float2 t1 = __fadd2_rn(make_float2(__uint_as_float(dr[sub][c + 2]), __uint_as_float(dr[sub][c + 3])), make_float2(-dd.z, -dd.w));
                            const uint32_t p0 = pk[sub * 16 + c8 * 4 + q2 * 2], p1 = pk[sub * 16 + c8 * 4 + q2 * 2 + 1];
                            dsp[sub][c8 * 4 + q2 * 2] = BF16 ? mul_bf16x2(p0, pack_bf16x2(t0.x, t0.y)) : mul_half2(p0, pack_half2(t0.x, t0.y));
                            dsp[sub][c8 * 4 + q2 * 2 + 1] = BF16 ? mul_bf16x2(p1, pack_bf16x2(t1.x, t1.y)) : mul_half2(p1, pack_half2(t1.x, t1.y));
                        }
                    }
                }
                if (ragged_kv) {
#pragma unroll
                    for (int x = 0; x < 16; ++x) dsp[sub][x] &= kv_keep;
                }
            };
            // one copy to TMEM (A of dK; over this warpgroup's own dP^T columns), one to the shared-memory atom that feeds dQ
            auto ds_store = [&](auto subc) {
                constexpr int sub = decltype(subc)::value;
                tmem_st16(tDP + sub * 16, dsp[sub]);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                    *reinterpret_cast<uint4*>(ds_atom + swz128(n, sub * 4 + c8)) =
                        make_uint4(dsp[sub][c8 * 4], dsp[sub][c8 * 4 + 1], dsp[sub][c8 * 4 + 2], dsp[sub][c8 * 4 + 3]);
            };
            // The first half is stored while the second is computed: its STS / STTM drain behind the math instead of
            // in front of the proxy fence below, which waits for every store of the thread (all 32 KB of the CTA's
            // dS^T going out at once cost ~350 cycles per step there; backward -2.2 %, profiles/r02/experiments.md).
            ds_math(std::integral_constant<int, 0>{});
            if (G > 0) mbar_wait(ds_empty, (G - 1) & 1);        // dQ(G-1) is done with the dS^T atoms (of both CTAs)
            if (i == 0 && it > 0) {
                // ... and so has the previous item's last dK / dV store, which was staged in this warpgroup's atom
                if (issuer) tma_store_wait_read<0>();
                named_bar_sync(ep_bar, 128);
            }
            ds_store(std::integral_constant<int, 0>{});
            ds_math(std::integral_constant<int, 1>{});
            ds_store(std::integral_constant<int, 1>{});
            tmem_wait_st();
            fence_proxy_async_smem();       // dS smem writes -> visible to the async proxy (tensor core / bulk copy)
            tc_fence_before();
            // the whole atom is written: warp 15 ships it into the peer's dQ operand buffer (issuing the 16 KB bulk copy
            // blocks the issuing thread for several hundred cycles, which a compute warp cannot afford)
            if (!keep_local) named_bar_arrive(10, 160);
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(ds_full); else mbar_arrive_cluster(ds_full_l); }
            if (warp == 0) { TL(11); TLF(23); }
        }

        // epilogue: warpgroup 0 stores dK (scaled), warpgroup 1 stores dV, 32 columns at a time through this warpgroup's
        // (now dead) dS atom as a 128B-swizzled fp32 box -> TMA store (rows past S are clipped by the tensor map)
        mbar_wait(dkdv_full, it & 1);
        tc_fence_after();
        if (threadIdx.x == 0) TLC(2);
        const float mul = (h == 0) ? (p.range != nullptr ? __ldg(p.range + kDkMul) : p.scale)
                                   : (p.range != nullptr ? __ldg(p.range + kDvMul) : 1.0f);
        const uint32_t tsrc = tmem_base + lane_addr + (h == 0 ? COL_DK : COL_DV);
        const CUtensorMap* tm_out = (h == 0) ? &p.tm_dk : &p.tm_dv;
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tsrc + c * 32, r);
            tmem_wait_ld();
            if (c > 0) {
                if (issuer) tma_store_wait_read<0>();       // the previous chunk has left the staging atom
                named_bar_sync(ep_bar, 128);
            }
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
                float4 v4;
                v4.x = __uint_as_float(r[q4 * 4]) * mul;
                v4.y = __uint_as_float(r[q4 * 4 + 1]) * mul;
                v4.z = __uint_as_float(r[q4 * 4 + 2]) * mul;
                v4.w = __uint_as_float(r[q4 * 4 + 3]) * mul;
                *reinterpret_cast<float4*>(ds_atom + swz128(n, q4)) = v4;
            }
            fence_proxy_async_smem();
            named_bar_sync(ep_bar + 1, 128);
            if (issuer) {
                tma_store_3d(tm_out, ds_atom, c * 32, kv_row0, bh);
                tma_store_commit();
            }
        }
        if (issuer) mbar_arrive(epi_issued);
        tc_fence_before();                  // dK / dV reads are complete before the next item's MMAs overwrite them
        if (threadIdx.x == 0) TLC(3);
        }
        if (issuer) tma_store_wait<0>();    // global writes done before the CTA retires
    } else if (warp < MMA_WARP) {
        // ------------------------------------------------------------------ dQ drain warps
        setmaxnreg_dec<104>();
        const int n = (warp & 3) * 32 + lane;                  // TMEM lane: query row 64 rank + n % 64, head-dim half n / 64
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tDQ = tmem_base + lane_addr + COL_DQ;
        uint8_t* stage = smem + L::OFF_DQS + (n >> 6) * HATOM;  // box of this lane half: [64 rows][32 fp32], 128B-swizzled
        const int row = n & 63;
        const bool issuer = warp == D_WARP0 && lane == 0;
        const uint32_t dq_empty_l = cluster_map(dq_empty, 0);

        int it = 0;
        for (int w = unit; w < n_work; w += n_units, ++it) {
        const int bh = w / n_pairs, u = w % n_pairs;
        for (int i = 0; i < n_q; ++i) {
            const int G = it * n_q + i;
            mbar_wait(dq_full, G & 1);
            tc_fence_after();
            if (warp == D_WARP0) TL(15);
            uint32_t r[2][32];
            tmem_ld32(tDQ, r[0]);
            tmem_ld32(tDQ + 32, r[1]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(dq_empty); else mbar_arrive_cluster(dq_empty_l); }   // dP(G+1) may overwrite the columns
            if (warp == D_WARP0) TL(16);

            // The dK / dV stores at the end of an item gate the next item; the last tile's reduce gates nothing and waits
            // until they have all been issued (the SM's write path is shared).
            if (i == n_q - 1) mbar_wait(epi_issued, it & 1);
            const int q_row = q_row_at(u, i) + 64 * static_cast<int>(rank);
#pragma unroll
            for (int c = 0; c < 2; ++c) {                       // 32 columns of each lane half per pass
                const float dq_mul = p.range != nullptr ? ldg_scalar_volatile(p.range + kDqMul) : p.scale;
                if (issuer) tma_store_wait_read<0>();           // the previous reduces have read the staging boxes
                named_bar_sync(1, 128);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    float4 v4;
                    v4.x = __uint_as_float(r[c][q4 * 4]) * dq_mul;
                    v4.y = __uint_as_float(r[c][q4 * 4 + 1]) * dq_mul;
                    v4.z = __uint_as_float(r[c][q4 * 4 + 2]) * dq_mul;
                    v4.w = __uint_as_float(r[c][q4 * 4 + 3]) * dq_mul;
                    *reinterpret_cast<float4*>(stage + swz128(row, q4)) = v4;
                }
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (issuer) {
                    tma_reduce_add_3d(&p.tm_dq64, smem + L::OFF_DQS, c * 32, q_row, bh);              // head-dim columns [32 c, +32)
                    tma_reduce_add_3d(&p.tm_dq64, smem + L::OFF_DQS + HATOM, 64 + c * 32, q_row, bh); // ... and [64 + 32 c, +32)
                    tma_store_commit();
                }
                if (warp == D_WARP0) TL(18 + c);
            }
            if (warp == D_WARP0) TL(17);
        }
        }
        if (issuer) tma_store_wait<0>();
    } else {
        // ------------------------------------------------------------------ warp 15: ships this CTA's dS^T send atom to the peer
        setmaxnreg_dec<56>();
        const uint32_t peer_dqa = cluster_map(smem + L::OFF_DQA + rank * ATOM, rank ^ 1);
        const uint32_t peer_recv = cluster_map(dqa_recv, rank ^ 1);
        for (int w = unit; w < n_work; w += n_units) {
            for (int i = 0; i < n_q; ++i) {
                named_bar_sync(10, 160);    // the sending warpgroup has written (and proxy-fenced) the atom
                if (elect_one()) dsmem_bulk_copy(peer_dqa, smem + L::OFF_SEND, ATOM, peer_recv);
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // nobody retires while its peer may still signal into its shared memory
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

}  // namespace

// FA2_BWD_PAIR=0 selects the single-CTA kernel of fa2_bwd_sm100.cu at D = 128 as well (A/B runs, debugging).
bool bwd_uses_pair(int D) {
    static const bool allowed = [] { const char* ev = getenv("FA2_BWD_PAIR"); return ev ? atoi(ev) != 0 : FA2_BWD_PAIR_DEFAULT; }();
    return allowed && D == 128;
}

cudaError_t launch_bwd2(const BwdParams& p, cudaStream_t st) {
    const int n_kvt = (p.S_kv + BT - 1) / BT;
    const long long n_work = static_cast<long long>(p.BH) * ((n_kvt + 1) / 2);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t err = ensure_smem_optin(reinterpret_cast<const void*>(kern), L::BYTES);
        if (err != cudaSuccess) return err;
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = L::BYTES; cfg.stream = st;
        static int max_clusters[64] = {0};
        if (dev < 64 && max_clusters[dev] == 0) {
            cfg.gridDim = dim3(static_cast<unsigned>(sm_count_current() & ~1));
            int n = 0;
            if ((err = cudaOccupancyMaxActiveClusters(&n, kern, &cfg)) != cudaSuccess) return err;
            max_clusters[dev] = n > 0 ? n : 1;
        }
        const long long cap = dev < 64 ? max_clusters[dev] : sm_count_current() / 2;
        cfg.gridDim = dim3(static_cast<unsigned>(2 * (n_work < cap ? n_work : cap)));
        return cudaLaunchKernelEx(&cfg, kern, p);
    };
    static const int poly = [] { const char* ev = getenv("FA2_BWD2_POLY"); return ev ? atoi(ev) : FA2_BWD2_POLY_DEFAULT; }();
    if (p.bf16) e = go(fa2_bwd2_kernel<true, 0u>);
    else if (poly == 1) e = go(fa2_bwd2_kernel<false, 0x5500u>);       // 25 %: second pair of every other group
    else if (poly == 2) e = go(fa2_bwd2_kernel<false, 0xFF00u>);       // 50 %: second pair of every group
    else if (poly == 3) e = go(fa2_bwd2_kernel<false, 0xAA00u | 0x00AAu | 0x5500u>);   // 75 %
    else e = go(fa2_bwd2_kernel<false, 0u>);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t warm_bwd2() {
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd2_kernel<false, 0u>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd2_kernel<false, 0x5500u>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd2_kernel<false, 0xFF00u>)) != cudaSuccess) return e;
    return cudaFuncGetAttributes(&a, fa2_bwd2_kernel<true, 0u>);
}

}  // namespace fa2
