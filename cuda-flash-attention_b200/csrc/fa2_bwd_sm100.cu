// fa2_bwd_sm100.cu -- FlashAttention-2 backward for sm_100a (replaces the reference's
// flash_attention2_backward_kernel, kernels/f-attn2-backward.cu:33-339).
//
// Like the reference, one work item is one KV tile of a (batch, head) slab: the CTA walks the Q
// tiles, keeps dK / dV on chip and adds its dQ contribution into global memory (reference:
// atomicAdd, f-attn2-backward.cu:298; here: TMA reduce-add in L2).  The kernel is persistent (one
// CTA per SM, items taken round-robin, all mbarrier parities from running counters): the next
// item's K/V load and S / dP MMAs overlap the dK / dV epilogue of the current one.  Tiles are
// 128x128 and the five products run on the tensor cores, transposed so that the KV row is
// the TMEM lane:
//     S^T  = K Q^T          (SS, both K-major)            -> TMEM [384,512)
//     dP^T = V dO^T         (SS, both K-major)            -> TMEM [256,384)
//     dV  += P^T dO         (A = P^T from TMEM, B = dO tile as MN-major)   -> TMEM [128,128+DP)
//     dK  += dS^T Q         (A = dS^T smem K-major,  B = Q tile MN-major)  -> TMEM [0,DP)
//     dQ   = dS K           (A = dS^T smem read MN-major, B = K tile MN-major) -> TMEM [256,256+DP)
// dQ shares its TMEM columns with dP^T (dP is dead once dS has been formed).
//
// Warp roles (20 warps = 5 warpgroups; within a pair of warpgroups, warpgroup h owns one half of the columns):
//   0-3 / 4-7    compute:  S^T -> P^T = 2^(S^T c - lse2[q]) -> 16-bit into TMEM for the dV MMA, then
//                          dS^T = P^T o (dP^T - D_i[q]) -> 128B-swizzled smem; at the end of an item they
//                          store dK / dV (TMEM -> the dead dS atoms as fp32 staging boxes -> TMA store)
//   8-11 / 12-15 dQ drain: TMEM -> registers (frees the dP/dQ columns at once) -> scaled, swizzled smem ->
//                          cp.reduce.async.bulk.tensor add.  fp32 reduce-add sustains only ~24 B/clk per SM
//                          (tools/reduce_probe.cu), i.e. >= 2730 cycles per 64 KB dQ tile -- more than the 2560
//                          tensor cycles of the five GEMMs -- so only the drain warps ever wait for it.
//   16 MMA issuer, 17 TMA producer (also stages LSE and D_i per Q tile), 18-19 register donors.
// The CTA launches with 640 x 96 registers; setmaxnreg moves them to compute 128 / drain 88 / rest 40.
// Padding: TMA zero-fills out-of-range Q/K/V/dO rows, out-of-range LSE (padded Q columns) is staged as
// +inf (P = 0), out-of-range dK/dV rows are not stored, and padded KV lanes (kv row >= S, where S^T = 0 and
// P = 2^(-lse) can overflow 16 bit for strongly negative LSE) are forced to P = dS = 0 by the compute
// warps, so that no inf * 0 reaches the dQ contraction over the KV rows.
#include "fa2_common.h"
#include "ptx.cuh"

namespace fa2 {
namespace {

constexpr int BT = 128;                 // tile rows (both KV and Q)
constexpr int ATOM = BT * 128;          // bytes of one [128 rows][64 x 16-bit] swizzle atom
constexpr int NUM_THREADS = 640;
constexpr int D_WARP0 = 8;
constexpr int MMA_WARP = 16;
constexpr int TMA_WARP = 17;
constexpr int Q_STAGES = 2;

// Debug timeline (only with -DFA2_TIMELINE): lane 0 of a role stamps clock64 into slot `slot` of iteration i.
#ifdef FA2_TIMELINE
#define TL(slot) do { if (p.timeline && w == 0 && lane == 0 && i < 32) p.timeline[i * 32 + (slot)] = clock64(); } while (0)
// per-work-item marks: [1024 + 8 * w + k], k = 0 item picked up, 1 first S seen, 2 dK/dV complete, 3 epilogue done, 4 SM id
#define TLC(k) do { if (p.timeline) p.timeline[1024 + 8 * w + (k)] = clock64(); } while (0)
#else
#define TL(slot) do { } while (0)
#define TLC(k) do { } while (0)
#endif

template <int DP>
struct BwdSmem {
    static constexpr int TILE = BT * DP * 2;
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = OFF_K + TILE;
    static constexpr int OFF_Q = OFF_V + TILE;                    // Q_STAGES tiles
    static constexpr int OFF_DO = OFF_Q + Q_STAGES * TILE;        // 1 tile
    static constexpr int OFF_DS = OFF_DO + TILE;                  // [128 kv][128 q] 16-bit, 2 atoms
    static constexpr int OFF_DQS = OFF_DS + 2 * ATOM;             // 2 x [128][32] fp32 staging (one per D warpgroup)
    static constexpr int OFF_LSE = OFF_DQS + 2 * BT * 128;        // Q_STAGES x 128 fp32
    static constexpr int OFF_DELTA = OFF_LSE + Q_STAGES * BT * 4;
    static constexpr int OFF_BAR = OFF_DELTA + Q_STAGES * BT * 4;
    static constexpr int NUM_BARS = 17;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int BYTES = OFF_TMEM_PTR + 16;
};

template <int DP, bool BF16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa2_bwd_kernel(const __grid_constant__ BwdParams p) {
    using L = BwdSmem<DP>;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t COL_DK = 0, COL_DV = 128, COL_DP = 256, COL_DQ = 256, COL_S = 384;
    constexpr int KSTEPS_D = DP / 16;       // contraction over the head dim
    constexpr int KSTEPS_T = BT / 16;       // contraction over a 128-row tile

    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* q_full = bars + 1;      // [2]
    uint64_t* q_empty = bars + 3;     // [2]
    uint64_t* do_full = bars + 5;
    uint64_t* do_empty = bars + 6;
    uint64_t* s_full = bars + 7;
    uint64_t* p_full = bars + 8;
    uint64_t* dp_full = bars + 9;
    uint64_t* ds_full = bars + 10;
    uint64_t* ds_empty = bars + 11;
    uint64_t* dq_full = bars + 12;
    uint64_t* dq_empty = bars + 13;
    uint64_t* dkdv_full = bars + 14;
    uint64_t* kv_empty = bars + 15;   // the item's last dQ / dP have read the K / V tiles
    uint64_t* epi_issued = bars + 16; // both compute warpgroups have issued their last dK / dV store of the item
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
    float* lse_s = reinterpret_cast<float*>(smem + L::OFF_LSE);
    float* delta_s = reinterpret_cast<float*>(smem + L::OFF_DELTA);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Persistent CTA: work item w = (bh, KV tile), taken round-robin so that the CTAs running at the same time
    // work on neighbouring KV tiles of the same few heads (Q / dO reuse in L2).  All items have n_tiles steps, so
    // every per-step barrier parity comes from the running step count G = it * n_tiles + i and every per-item
    // parity from it (items this CTA has started).
    const int n_tiles = (p.S_q + BT - 1) / BT;     // Q tiles = steps of every work item
    const int n_kvt = (p.S_kv + BT - 1) / BT;      // KV tiles of a slab = work items per slab
    const int n_work = p.BH * n_kvt;
    // Every KV tile of a (b,h) slab walks the Q tiles in a different rotation, so that at any moment the
    // concurrently running CTAs reduce-add into DIFFERENT dQ tiles (no same-address contention in L2).
    auto q_row_at = [&](int kv_tile, int i) { return ((i + kv_tile) % n_tiles) * BT; };
    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
        tma_prefetch_desc(&p.tm_do);
        tma_prefetch_desc(&p.tm_dq);
        tma_prefetch_desc(&p.tm_dk);
        tma_prefetch_desc(&p.tm_dv);
    }
    if (warp == MMA_WARP) {
        if (lane == 0) {
            if (smem_u32(smem) & 1023u) __trap();          // swizzled tiles need a 1024-B aligned base
            mbar_init(kv_full, 1);
            for (int i = 0; i < Q_STAGES; ++i) {
                mbar_init(&q_full[i], 2);     // TMA expect_tx arrive + LSE/D_i staging arrive
                mbar_init(&q_empty[i], 1);
            }
            mbar_init(do_full, 1);
            mbar_init(do_empty, 1);
            mbar_init(s_full, 1);
            mbar_init(p_full, 8);             // one arrive per compute warp
            mbar_init(dp_full, 1);
            mbar_init(ds_full, 8);            // one arrive per compute warp
            mbar_init(ds_empty, 1);
            mbar_init(dq_full, 1);
            mbar_init(dq_empty, 8);           // one arrive per drain warp
            mbar_init(dkdv_full, 1);
            mbar_init(kv_empty, 1);
            mbar_init(epi_issued, 2);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_holder, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == TMA_WARP) {
        // ------------------------------------------------------------------ producer
        setmaxnreg_dec<40>();
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int bh = w / n_kvt, kv_tile = w % n_kvt, kv_row0 = kv_tile * BT;
        auto q_row_of = [&](int i) { return q_row_at(kv_tile, i); };
        mbar_wait(kv_empty, (it & 1) ^ 1);                     // previous item's MMAs are done with K / V
        if (elect_one()) {
            mbar_expect_tx(kv_full, 2 * L::TILE);
            for (int a = 0; a < DP / 64; ++a) {
                tma_load_3d(smem + L::OFF_K + a * ATOM, &p.tm_k, kv_full, a * 64, kv_row0, bh);
                tma_load_3d(smem + L::OFF_V + a * ATOM, &p.tm_v, kv_full, a * 64, kv_row0, bh);
            }
        }
        __syncwarp();
        for (int i = 0; i < n_tiles; ++i) {
            const int G = it * n_tiles + i;
            const int s = G % Q_STAGES;
            const uint32_t ph = (G / Q_STAGES) & 1;
            // LSE (log2 domain) and D_i of this Q tile are fetched into registers BEFORE the wait for the stage, so
            // that their global-memory latency is not on the Q(i) -> S(i) critical path; rows past S: +inf / 0
            // => P = 0, dS = 0
            float r_lse[BT / 32], r_dl[BT / 32];
            const float delta_mul = p.range != nullptr ? ldg_scalar_volatile(p.range + kDeltaMul) : 1.0f;   // D_i -> units of dP' = dP s_v s_do
#pragma unroll
            for (int r = 0; r < BT / 32; ++r) {
                const int row = q_row_of(i) + r * 32 + lane;
                const bool ok = row < p.S_q;
                const size_t g = static_cast<size_t>(bh) * p.S_q + (ok ? row : 0);
                r_lse[r] = ok ? __ldg(p.lse_log2 + g) : INFINITY;
                r_dl[r] = ok ? __ldg(p.delta + g) * delta_mul : 0.0f;
            }
            mbar_wait(&q_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&q_full[s], L::TILE);
                for (int a = 0; a < DP / 64; ++a)
                    tma_load_3d(smem + L::OFF_Q + s * L::TILE + a * ATOM, &p.tm_q, &q_full[s], a * 64, q_row_of(i), bh);
            }
#pragma unroll
            for (int r = 0; r < BT / 32; ++r) {
                lse_s[s * BT + r * 32 + lane] = r_lse[r];
                delta_s[s * BT + r * 32 + lane] = r_dl[r];
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(&q_full[s]);           // LSE / D_i staged (the Q tile itself lands via TMA)
            __syncwarp();
            mbar_wait(do_empty, (G & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(do_full, L::TILE);
                for (int a = 0; a < DP / 64; ++a)
                    tma_load_3d(smem + L::OFF_DO + a * ATOM, &p.tm_do, do_full, a * 64, q_row_of(i), bh);
            }
            __syncwarp();
        }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (converged code lets the compiler keep descriptors in uniform
        // registers); one elected lane issues each group of MMAs and its commits.
        setmaxnreg_dec<40>();
        {
            const uint32_t id_ss = umma_idesc_f16(BT, BT, 0, 0, BF16 ? 1 : 0);    // S^T, dP^T
            const uint32_t id_kmn = umma_idesc_f16(BT, DP, 0, 1, BF16 ? 1 : 0);   // dV, dK : A K-major, B MN-major
            const uint32_t id_mnmn = umma_idesc_f16(BT, DP, 1, 1, BF16 ? 1 : 0);  // dQ     : A MN-major, B MN-major
            const uint32_t hi = umma_desc_hi(1024);
            // low descriptor words; "_k" = read K-major (LBO unused), "_mn" = read MN-major (LBO = next 64-wide chunk)
            const uint32_t k_k = umma_desc_lo(smem_u32(smem + L::OFF_K), 16), k_mn = umma_desc_lo(smem_u32(smem + L::OFF_K), ATOM);
            const uint32_t v_k = umma_desc_lo(smem_u32(smem + L::OFF_V), 16);
            const uint32_t q_k = umma_desc_lo(smem_u32(smem + L::OFF_Q), 16), q_mn = umma_desc_lo(smem_u32(smem + L::OFF_Q), ATOM);
            const uint32_t do_k = umma_desc_lo(smem_u32(smem + L::OFF_DO), 16), do_mn = umma_desc_lo(smem_u32(smem + L::OFF_DO), ATOM);
            const uint32_t ds_k = umma_desc_lo(smem_u32(smem + L::OFF_DS), 16), ds_mn = umma_desc_lo(smem_u32(smem + L::OFF_DS), ATOM);
            constexpr uint32_t TILE16 = L::TILE >> 4;
            const uint32_t tS = tmem_base + COL_S, tDP = tmem_base + COL_DP, tDQ = tmem_base + COL_DQ;
            const uint32_t tDK = tmem_base + COL_DK, tDV = tmem_base + COL_DV;

            auto issue_s = [&](int st) {
                static_for<KSTEPS_D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ss_off<koff_kmajor(k, ATOM), koff_kmajor(k, ATOM)>(tS, k_k, q_k + st * TILE16, hi, id_ss, k > 0);
                });
            };
            auto issue_dp = [&]() {
                static_for<KSTEPS_D>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ss_off<koff_kmajor(k, ATOM), koff_kmajor(k, ATOM)>(tDP, v_k, do_k, hi, id_ss, k > 0);
                });
            };
            auto issue_dv = [&](bool first) {      // P^T lives in two 32-column runs of the S region
                static_for<KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ts_off<(k >> 2) * 64 + (k & 3) * 8, koff_mnmajor(k)>(tDV, tS, do_mn, hi, id_kmn, (!first || k > 0) ? 1u : 0u);
                });
            };
            auto issue_dk = [&](int st, bool first) {
                static_for<KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ss_off<koff_kmajor(k, ATOM), koff_mnmajor(k)>(tDK, ds_k, q_mn + st * TILE16, hi, id_kmn, (!first || k > 0) ? 1u : 0u);
                });
            };
            auto issue_dq = [&]() {
                static_for<KSTEPS_T>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ss_off<koff_mnmajor(k), koff_mnmajor(k)>(tDQ, ds_mn, k_mn, hi, id_mnmn, k > 0);
                });
            };

            int it = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int G0 = it * n_tiles;
            mbar_wait_spin(kv_full, it & 1);
            for (int i = 0; i < n_tiles; ++i) {
                const int G = G0 + i;
                const int st = G % Q_STAGES;
                mbar_wait_spin(&q_full[st], (G / Q_STAGES) & 1);
                tc_fence_after();
                TL(0);
                if (elect_one()) {
                    issue_s(st);                               // dV(G-1) was issued before, so P^T(G-1) is consumed
                    umma_commit(s_full);
                }
                __syncwarp();
                if (i > 0) {
                    mbar_wait_spin(ds_full, (G - 1) & 1);
                    tc_fence_after();
                    TL(1);
                    if (elect_one()) {
                        issue_dq();                            // dQ(i-1) over the dead dP(i-1)
                        umma_commit(dq_full);
                        issue_dk((G - 1) % Q_STAGES, i == 1);  // dK += dS(i-1)^T Q(i-1)
                        umma_commit(&q_empty[(G - 1) % Q_STAGES]);
                        umma_commit(ds_empty);
                    }
                    __syncwarp();
                }
                if (G > 0) {
                    mbar_wait_spin(dq_empty, (G - 1) & 1);          // dQ(G-1) drained out of TMEM (i == 0: previous item's last)
                    TL(2);
                }
                mbar_wait_spin(do_full, G & 1);
                tc_fence_after();
                TL(3);
                if (elect_one()) {
                    issue_dp();
                    umma_commit(dp_full);
                }
                __syncwarp();
                mbar_wait_spin(p_full, G & 1);
                tc_fence_after();
                TL(4);
                if (elect_one()) {
                    issue_dv(i == 0);
                    umma_commit(do_empty);
                }
                __syncwarp();
            }
            {
                const int i = n_tiles - 1, G = G0 + i;
                mbar_wait_spin(ds_full, G & 1);
                tc_fence_after();
                if (elect_one()) {
                    issue_dk(G % Q_STAGES, i == 0);
                    issue_dq();
                    umma_commit(dq_full);
                    umma_commit(dkdv_full);
                    umma_commit(&q_empty[G % Q_STAGES]);       // the next item reuses the Q stage, the dS tile
                    umma_commit(ds_empty);                     // and the K / V tiles
                    umma_commit(kv_empty);
                }
                __syncwarp();
            }
            }
        }
    } else if (warp < D_WARP0) {
        // ------------------------------------------------------------------ compute warps: P^T and dS^T
        setmaxnreg_inc<128>();
        const int h = warp >> 2;                               // which 64 Q-columns of the tile
        const int n = (warp & 3) * 32 + lane;                  // kv row within the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + COL_S + h * 64;
        const uint32_t tDP = tmem_base + lane_addr + COL_DP + h * 64;
        const float c2 = p.range != nullptr ? __ldg(p.range + kC2) : p.scale_log2;
        const float2 c2v = make_float2(c2, c2);
        uint8_t* ds_atom = smem + L::OFF_DS + h * ATOM;        // Q columns [64h, 64h+64) = swizzle atom h
        const bool issuer = ((warp & 3) == 0) && lane == 0;    // owns this warpgroup's dK / dV store groups
        const uint32_t ep_bar = 5 + 2 * h;                      // named barriers private to this warpgroup
        const int n_chunk = (p.D + 31) / 32;                   // real columns only (D = 32 under DP = 64)

        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int bh = w / n_kvt, kv_row0 = (w % n_kvt) * BT;
        const int G0 = it * n_tiles;
        const bool ragged_kv = kv_row0 + BT > p.S_kv;                          // CTA-uniform: only the last KV tile of a slab
        const uint32_t kv_keep = (kv_row0 + n < p.S_kv) ? 0xffffffffu : 0u;   // padded KV lane: P = dS = 0
#ifdef FA2_TIMELINE
        if (threadIdx.x == 0 && p.timeline) {
            uint32_t smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            p.timeline[1024 + 8 * w + 4] = smid;
            TLC(0);
        }
#endif
        for (int i = 0; i < n_tiles; ++i) {
            const int G = G0 + i;
            const int st = G % Q_STAGES;
            const float4* lse_t = reinterpret_cast<const float4*>(lse_s + st * BT + h * 64);
            const float4* dl_t = reinterpret_cast<const float4*>(delta_s + st * BT + h * 64);
            mbar_wait(&q_full[st], (G / Q_STAGES) & 1);        // LSE / D_i staging visible
            mbar_wait(s_full, G & 1);
            tc_fence_after();
            if (warp == 0) TL(8);
            if (i == 0 && threadIdx.x == 0) TLC(1);
            uint32_t pk[32];                                    // P^T row (64 values) rounded to 16 bit
            {
                uint32_t sr[2][32];
                tmem_ld32(tS, sr[0]);
                tmem_ld32(tS + 32, sr[1]);
                tmem_wait_ld();
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
                            const float e0 = ex2_approx(x0.x), e1 = ex2_approx(x0.y), e2 = ex2_approx(x1.x), e3 = ex2_approx(x1.y);
                            pk[sub * 16 + 2 * cc] = BF16 ? pack_bf16x2(e0, e1) : pack_half2(e0, e1);
                            pk[sub * 16 + 2 * cc + 1] = BF16 ? pack_bf16x2(e2, e3) : pack_half2(e2, e3);
                        }
                    }
                }
            }
            if (ragged_kv) {
#pragma unroll
                for (int x = 0; x < 32; ++x) pk[x] &= kv_keep;
            }
            tmem_st32(tS, pk);                                  // over the S columns this thread already consumed
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (warp == 0) TL(9);

            mbar_wait(dp_full, G & 1);
            tc_fence_after();
            if (warp == 0) TL(10);
            uint32_t dr[2][32];
            tmem_ld32(tDP, dr[0]);
            tmem_ld32(tDP + 32, dr[1]);
            tmem_wait_ld();
            tc_fence_before();              // dP reads are complete before dQ may overwrite the columns
            if (G > 0) mbar_wait(ds_empty, (G - 1) & 1);        // dK(G-1), dQ(G-1) finished reading dS smem
            if (i == 0 && it > 0) {
                // ... and so has the previous item's last dK / dV store, which was staged in this atom
                if (issuer) tma_store_wait_read<0>();
                named_bar_sync(ep_bar, 128);
            }
            // dS^T = P^T o (dP^T - D_i): the difference in fp32 (the staged D_i already carries the operand scales of
            // dP^T), the product in packed 16-bit (it is rounded to 16 bit for the tensor core anyway)
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float4 d4[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) d4[c] = dl_t[sub * 8 + hf * 4 + c];
#pragma unroll
                    for (int c8l = 0; c8l < 2; ++c8l) {        // 8 columns -> one 16-byte chunk of dS^T
                        const int c8 = hf * 2 + c8l;
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float4 dd = d4[c8l * 2 + u];
                            const int c = c8 * 8 + u * 4;
                            const float2 t0 = __fadd2_rn(make_float2(__uint_as_float(dr[sub][c]), __uint_as_float(dr[sub][c + 1])), make_float2(-dd.x, -dd.y));
                            const float2 t1 = __fadd2_rn(make_float2(__uint_as_float(dr[sub][c + 2]), __uint_as_float(dr[sub][c + 3])), make_float2(-dd.z, -dd.w));
                            const uint32_t p0 = pk[sub * 16 + c8 * 4 + u * 2], p1 = pk[sub * 16 + c8 * 4 + u * 2 + 1];
                            w[u * 2] = BF16 ? mul_bf16x2(p0, pack_bf16x2(t0.x, t0.y)) : mul_half2(p0, pack_half2(t0.x, t0.y));
                            w[u * 2 + 1] = BF16 ? mul_bf16x2(p1, pack_bf16x2(t1.x, t1.y)) : mul_half2(p1, pack_half2(t1.x, t1.y));
                        }
                        if (ragged_kv) { w[0] &= kv_keep; w[1] &= kv_keep; w[2] &= kv_keep; w[3] &= kv_keep; }
                        *reinterpret_cast<uint4*>(ds_atom + swz128(n, sub * 4 + c8)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            fence_proxy_async_smem();       // dS smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full);
            if (warp == 0) TL(11);
        }

        // epilogue: warpgroup 0 stores dK (scaled by 1/sqrt(D)), warpgroup 1 stores dV, 32 columns at a time
        // through this warpgroup's (now dead) dS atom as a 128B-swizzled fp32 box -> TMA store (rows past S are
        // clipped by the tensor map).  The MMA warp is already computing S / dP of the next item.
        mbar_wait(dkdv_full, it & 1);
        tc_fence_after();
        if (threadIdx.x == 0) TLC(2);
        const float mul = (h == 0) ? (p.range != nullptr ? __ldg(p.range + kDkMul) : p.scale)
                                   : (p.range != nullptr ? __ldg(p.range + kDvMul) : 1.0f);
        const uint32_t tsrc = tmem_base + lane_addr + (h == 0 ? COL_DK : COL_DV);
        const CUtensorMap* tm_out = (h == 0) ? &p.tm_dk : &p.tm_dv;
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
            if (c < n_chunk) {
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
        }
        if (issuer) mbar_arrive(epi_issued);
        tc_fence_before();                  // dK / dV reads are complete before the next item's MMAs overwrite them
        if (threadIdx.x == 0) TLC(3);
        }
        if (issuer) tma_store_wait<0>();    // global writes done before the CTA retires
    } else if (warp < MMA_WARP) {
        // ------------------------------------------------------------------ dQ drain warps
        setmaxnreg_dec<88>();
        const int h = (warp - D_WARP0) >> 2;                   // head-dim half of dQ this warpgroup drains
        const int n = (warp & 3) * 32 + lane;                  // TMEM lane == q row of the tile
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        constexpr int CPW = DP / 64;                            // 32-column chunks per warpgroup (2 at D=128, else 1)
        const uint32_t tDQ = tmem_base + lane_addr + COL_DQ + h * CPW * 32;
        uint8_t* stage = smem + L::OFF_DQS + h * (BT * 128);   // dedicated 16 KB buffer
        const bool issuer = ((warp & 3) == 0) && lane == 0;
        const int n_chunk = (p.D + 31) / 32;                   // real columns only (D = 32 under DP = 64)
        const uint32_t bar_id = 1 + 2 * h;                      // named barriers private to this warpgroup

        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int bh = w / n_kvt, kv_tile = w % n_kvt;
        auto q_row_of = [&](int i) { return q_row_at(kv_tile, i); };
        auto put_chunk = [&](const uint32_t (&rc)[32], int chunk, int i) {
            const float dq_mul = p.range != nullptr ? ldg_scalar_volatile(p.range + kDqMul) : p.scale;   // 1 / sqrt(D), inverse scales
            if (issuer) tma_store_wait_read<0>();               // previous reduce out of the buffer has been read
            named_bar_sync(bar_id, 128);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
                float4 v4;
                v4.x = __uint_as_float(rc[q4 * 4]) * dq_mul;
                v4.y = __uint_as_float(rc[q4 * 4 + 1]) * dq_mul;
                v4.z = __uint_as_float(rc[q4 * 4 + 2]) * dq_mul;
                v4.w = __uint_as_float(rc[q4 * 4 + 3]) * dq_mul;
                *reinterpret_cast<float4*>(stage + swz128(n, q4)) = v4;
            }
            fence_proxy_async_smem();
            named_bar_sync(bar_id + 1, 128);
            if (issuer) {
                tma_reduce_add_3d(&p.tm_dq, stage, chunk * 32, q_row_of(i), bh);
                tma_store_commit();
            }
        };

        for (int i = 0; i < n_tiles; ++i) {
            mbar_wait(dq_full, (it * n_tiles + i) & 1);
            tc_fence_after();
            if (warp == D_WARP0) TL(15);
            uint32_t r[CPW][32];
#pragma unroll
            for (int c = 0; c < CPW; ++c) tmem_ld32(tDQ + c * 32, r[c]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);               // dP(i+1) may overwrite the columns now
            if (warp == D_WARP0) TL(16);

            // The SM's write path (~24 B/clk) is shared with the compute warps' dK / dV stores at the end of an item,
            // and those gate the next item: the last tile's reduce (which gates nothing; these warps are idle for
            // the first ~5K cycles of the next item anyway) waits until they have all been issued.
            if (i == n_tiles - 1) mbar_wait(epi_issued, it & 1);
            // Only the drain warps ever wait for the reduce engine (they hold the tile in registers).
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const int chunk = h * CPW + c;
                if (chunk < n_chunk) put_chunk(r[c], chunk, i);   // 16 KB box: [128 rows][32 fp32], 128B-swizzled
                if (warp == D_WARP0) TL(18 + c);
            }
            if (warp == D_WARP0) TL(17);
        }
        }
        if (issuer) tma_store_wait<0>();
    } else {
        setmaxnreg_dec<40>();              // warps 18, 19: register donors only
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

cudaError_t launch_bwd(const BwdParams& p, cudaStream_t st) {
    if (bwd_uses_pair(p.D)) return launch_bwd2(p, st);
    const int DP = padded_head_dim(p.D);
    const int n_tiles = (p.S_kv + BT - 1) / BT;
    // persistent: one CTA per SM (or fewer when there is less work), each walks its share of the work items
    static int sm_count[64] = {0};
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (dev < 64 && sm_count[dev] == 0 &&
        (e0 = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
        return e0;
    const long long n_work = static_cast<long long>(p.BH) * n_tiles;
    const int n_sm = dev < 64 ? sm_count[dev] : 148;
    const dim3 grid(static_cast<unsigned>(n_work < n_sm ? n_work : n_sm));
    cudaError_t e;
    auto go = [&](auto kern, int smem) -> cudaError_t {
        cudaError_t err = ensure_smem_optin(reinterpret_cast<const void*>(kern), smem);
        if (err != cudaSuccess) return err;
        kern<<<grid, NUM_THREADS, smem, st>>>(p);
        return cudaSuccess;
    };
    if (DP == 64) e = p.bf16 ? go(fa2_bwd_kernel<64, true>, BwdSmem<64>::BYTES) : go(fa2_bwd_kernel<64, false>, BwdSmem<64>::BYTES);
    else          e = p.bf16 ? go(fa2_bwd_kernel<128, true>, BwdSmem<128>::BYTES) : go(fa2_bwd_kernel<128, false>, BwdSmem<128>::BYTES);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}


cudaError_t warm_bwd() {
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd_kernel<64, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd_kernel<128, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd_kernel<64, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, fa2_bwd_kernel<128, true>)) != cudaSuccess) return e;
    return warm_bwd2();
}

}  // namespace fa2
