// fa2_fwd_sm100.cu -- FlashAttention-2 forward for sm_100a (replaces the reference's
// flash_attention2_forward_kernel, kernels/kernel_fa2_optimized.cu:19-347).
//
// Persistent kernel: one CTA per SM; work item = 256 query rows of one (batch, head) slab as two
// 128-row tiles, walked over the 128-row KV tiles; items are taken round-robin so that the CTAs
// running at the same time share the K/V of a few heads in L2.  Roles (12 warps):
//   warps 0-3  softmax + epilogue for Q tile 0   (thread == one query row, TMEM lane == row)
//   warps 4-7  softmax + epilogue for Q tile 1
//   warp  8    MMA issuer (one elected thread issues tcgen05.mma; the warp also owns TMEM alloc/free)
//   warp  9    TMA producer (Q per item, K/V double-buffered through shared memory)
//   warps 10-11 register donors (setmaxnreg); in the fused forward+backward call they also cast dO to
//              16 bit and zero-fill dQ for the backward while the tensor-core loop runs
// S_t = Q_t K_j^T and O_t += P_t V_j run as tcgen05.mma (kind::f16, fp32 accumulate) with
// S and O in TMEM; P is written back to TMEM over S as 16-bit and fed to the second MMA as
// the A operand straight from TMEM.  The online softmax keeps max / sum per thread in
// registers, works in the exp2 domain and only rescales O when the running max moved by
// more than 2^8 (warp-uniform decision).  Epilogue: O / l goes through a 128B-swizzled fp32
// staging box and a TMA store, LSE = ln(l) + m (reference epilogue: kernel_fa2_optimized.cu:327-346);
// meanwhile the MMA warp already computes S(0) of the next item.  Every mbarrier parity is derived
// from running counters, so the pipeline never drains between items.
#include "fa2_common.h"
#include "fa2_range.cuh"
#include "ptx.cuh"

namespace fa2 {

namespace {

constexpr int BM = 128;         // query rows per tile
constexpr int BN = 128;         // kv rows per tile
constexpr int KV_STAGES = 2;
constexpr int NUM_THREADS = 384;        // 3 warpgroups; warps 10, 11 only donate registers
constexpr int MMA_WARP = 8;
constexpr int TMA_WARP = 9;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units
// Which exponentials go through the FMA-pipe polynomial (ex2_poly2) instead of MUFU.EX2: one bit per group of 4
// columns inside a 32-column chunk, FA2_POLY_A for the first pair of the group, FA2_POLY_B for the second.
// MUFU (16 exps/clk/SM, shared by both softmax warpgroups) is a co-bottleneck of the forward; with the persistent
// kernel a 25 % offload (B = 0x55) measures 6-8 % faster at D=64 and D=128 (tools/fwd_variants.py), 50 % is slower
// again (FMA-pipe issue slots).
#ifndef FA2_POLY_A
#define FA2_POLY_A 0x00
#endif
#ifndef FA2_POLY_B
#define FA2_POLY_B 0x55
#endif
constexpr unsigned POLY_A = FA2_POLY_A, POLY_B = FA2_POLY_B;

#ifdef FA2_TIMELINE
#define TLF(slot) do { if (p.timeline && w == 0 && lane == 0 && j < 32) p.timeline[j * 32 + (slot)] = clock64(); } while (0)
// per-work-item marks: [1024 + 8 * w + k], k = 0 item picked up, 1 first S seen, 2 last O seen, 3 epilogue done, 4 SM id
#define TLC(k) do { if (p.timeline) p.timeline[1024 + 8 * w + (k)] = clock64(); } while (0)
#else
#define TLF(slot) do { } while (0)
#define TLC(k) do { } while (0)
#endif

template <int DP>
struct FwdSmem {
    static constexpr int TILE_BYTES = BM * DP * 2;          // one 128-row 16-bit tile
    static constexpr int ATOM_BYTES = BM * 128;              // one [128][64] swizzle-atom column
    static constexpr int OFF_Q = 0;                           // 2 tiles
    static constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;      // KV_STAGES tiles
    static constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
    static constexpr int STAGE_BYTES = BM * 128;             // fp32 [128 rows][32 cols] O staging box, one per Q tile
    static constexpr int OFF_STAGE = OFF_V + KV_STAGES * TILE_BYTES;
    static constexpr int OFF_BAR = OFF_STAGE + 2 * STAGE_BYTES;
    static constexpr int NUM_BARS = 2 + 4 * KV_STAGES + 2 + 4 + 2 + 1;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int BYTES = OFF_TMEM_PTR + 16;
    static constexpr int ALLOC = BYTES + 1024;                // slack for manual 1024-B alignment
};

template <int DP, bool BF16, bool FUSED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa2_fwd_kernel(const __grid_constant__ FwdParams p) {
    using L = FwdSmem<DP>;
    constexpr int KSTEPS_QK = DP / 16;          // UMMA K = 16 for 16-bit operands
    constexpr int KSTEPS_PV = BN / 16;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 256 + DP;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* q_full = bars;                       // [2]
    uint64_t* k_full = bars + 2;                   // [KV_STAGES]
    uint64_t* k_empty = k_full + KV_STAGES;
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* s_full = v_empty + KV_STAGES;        // [2]
    uint64_t* p_full = s_full + 2;                 // [2 tiles][2 halves of the KV columns]
    uint64_t* o_full = p_full + 4;                 // [2]
    uint64_t* q_empty = o_full + 2;                // [1] every S of this work item has been computed
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Persistent CTA: work item w = (bh, pair of Q tiles), taken round-robin so that the CTAs running at the same
    // time walk the K/V of the same few heads (L2 reuse).  Every barrier parity below comes from running counters:
    //   it        work items this CTA has started           (q_empty, Q tile 0, o_full[0])
    //   n1        of those, the ones whose second Q tile exists (Q tile 1, o_full[1])
    //   it * n_kv + j   KV ring position / tile 0 step;   n1 * n_kv + j   tile 1 step
    const int q_blocks = (p.S_q + 2 * BM - 1) / (2 * BM);
    const int n_work = p.BH * q_blocks;
    const int n_kv = (p.S_kv + BN - 1) / BN;
    auto tiles_of = [&](int w) { return ((w % q_blocks) * (2 * BM) + BM < p.S_q) ? 2 : 1; };

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
        tma_prefetch_desc(&p.tm_o);
    }
    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&q_full[i], 1);
                mbar_init(&s_full[i], 1);
                mbar_init(&p_full[2 * i], 4);       // one arrive per softmax warp
                mbar_init(&p_full[2 * i + 1], 4);
                mbar_init(&o_full[i], 1);
            }
            for (int i = 0; i < KV_STAGES; ++i) {
                mbar_init(&k_full[i], 1);
                mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1);
                mbar_init(&v_empty[i], 1);
            }
            mbar_init(q_empty, 1);
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
        // ------------------------------------------------------------------ TMA producer
        setmaxnreg_dec<56>();
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int bh = w / q_blocks, q_row0 = (w % q_blocks) * (2 * BM), n_qt = tiles_of(w);
            mbar_wait(q_empty, (it & 1) ^ 1);          // the previous work item's last S has read the Q tiles
            if (elect_one()) {
                for (int t = 0; t < n_qt; ++t) {
                    mbar_expect_tx(&q_full[t], L::TILE_BYTES);
                    for (int a = 0; a < DP / 64; ++a)
                        tma_load_3d(smem + L::OFF_Q + t * L::TILE_BYTES + a * L::ATOM_BYTES, &p.tm_q, &q_full[t],
                                    a * 64, q_row0 + t * BM, bh);
                }
            }
            __syncwarp();
            for (int j = 0; j < n_kv; ++j) {
                const int g = it * n_kv + j;
                const int s = g % KV_STAGES;
                const uint32_t ph = (g / KV_STAGES) & 1;
                mbar_wait(&k_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&k_full[s], L::TILE_BYTES);
                    for (int a = 0; a < DP / 64; ++a)
                        tma_load_3d(smem + L::OFF_K + s * L::TILE_BYTES + a * L::ATOM_BYTES, &p.tm_k, &k_full[s],
                                    a * 64, j * BN, bh);
                }
                __syncwarp();
                mbar_wait(&v_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&v_full[s], L::TILE_BYTES);
                    for (int a = 0; a < DP / 64; ++a)
                        tma_load_3d(smem + L::OFF_V + s * L::TILE_BYTES + a * L::ATOM_BYTES, &p.tm_v, &v_full[s],
                                    a * 64, j * BN, bh);
                }
                __syncwarp();
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        // whole warp runs the loop (converged); one elected lane issues each MMA group and its commits
        setmaxnreg_dec<56>();
        {
            const uint32_t idesc_qk = umma_idesc_f16(BM, BN, 0, 0, BF16 ? 1 : 0);
            const uint32_t idesc_pv = umma_idesc_f16(BM, DP, 0, 1, BF16 ? 1 : 0);
            const uint32_t hi = umma_desc_hi(1024);
            const uint32_t q_lo = umma_desc_lo(smem_u32(smem + L::OFF_Q), 16);
            const uint32_t k_lo = umma_desc_lo(smem_u32(smem + L::OFF_K), 16);
            const uint32_t v_lo = umma_desc_lo(smem_u32(smem + L::OFF_V), L::ATOM_BYTES);   // LBO = next 64-col chunk
            constexpr uint32_t TILE16 = L::TILE_BYTES >> 4;

            auto issue_qk = [&](int t, int s) {
                // S_t = Q_t K^T : A, B both K-major, DP/64 swizzle atoms, 4 K-steps of 32 B per atom
                const uint32_t d = tmem_base + (t ? COL_S1 : COL_S0);
                const uint32_t a = q_lo + t * TILE16, b = k_lo + s * TILE16;
                static_for<KSTEPS_QK>([&](auto kk) {
                    constexpr int k = decltype(kk)::value;
                    umma_ss_off<koff_kmajor(k, L::ATOM_BYTES), koff_kmajor(k, L::ATOM_BYTES)>(d, a, b, hi, idesc_qk, k > 0);
                });
            };
            auto issue_pv = [&](int t, int s, bool first, auto half) {
                // O_t += P_t V : A = P from TMEM (8 columns per K-step), B = V MN-major; one half = 64 KV rows
                constexpr int h = decltype(half)::value;
                const uint32_t d = tmem_base + (t ? COL_O1 : COL_O0);
                const uint32_t a = tmem_base + (t ? COL_S1 : COL_S0), b = v_lo + s * TILE16;
                static_for<KSTEPS_PV / 2>([&](auto kk) {
                    constexpr int k = h * (KSTEPS_PV / 2) + decltype(kk)::value;
                    umma_ts_off<k * 8, koff_mnmajor(k)>(d, a, b, hi, idesc_pv, (!first || k > 0) ? 1u : 0u);
                });
            };

            int it = 0, n1 = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                const int n_qt = tiles_of(w);
                const int g0 = it * n_kv;                             // KV ring position of this item's step 0
                // S(0) of both tiles.  The S / P columns are free (the last P V of the previous item was issued
                // before this), so this overlaps the softmax warps' epilogue of the previous item.
                mbar_wait_spin(&k_full[g0 % KV_STAGES], (g0 / KV_STAGES) & 1);
                for (int t = 0; t < n_qt; ++t) {
                    mbar_wait_spin(&q_full[t], (t == 0 ? it : n1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_qk(t, g0 % KV_STAGES);
                        umma_commit(&s_full[t]);
                        if (t == n_qt - 1) {
                            umma_commit(&k_empty[g0 % KV_STAGES]);   // K(0) is free once every S(0) has been computed
                            if (n_kv == 1) umma_commit(q_empty);
                        }
                    }
                    __syncwarp();
                }
                for (int j = 0; j < n_kv; ++j) {
                    const int g = g0 + j;
                    const int s = g % KV_STAGES;
                    const uint32_t ph = (g / KV_STAGES) & 1;
                    const int s1 = (g + 1) % KV_STAGES;
                    const uint32_t ph1 = ((g + 1) / KV_STAGES) & 1;
                    mbar_wait_spin(&v_full[s], ph);
                    for (int t = 0; t < n_qt; ++t) {
                        const uint32_t pp = ((t == 0 ? it : n1) * n_kv + j) & 1;
                        mbar_wait_spin(&p_full[2 * t], pp);
                        tc_fence_after();
                        TLF(t == 0 ? 0 : 2);
                        if (elect_one()) issue_pv(t, s, j == 0, std::integral_constant<int, 0>{});
                        __syncwarp();
                        mbar_wait_spin(&p_full[2 * t + 1], pp);
                        if (j + 1 < n_kv && t == 0) mbar_wait_spin(&k_full[s1], ph1);
                        tc_fence_after();
                        TLF(t == 0 ? 1 : 3);
                        if (elect_one()) {
                            issue_pv(t, s, j == 0, std::integral_constant<int, 1>{});
                            if (t == n_qt - 1) umma_commit(&v_empty[s]);
                            if (j + 1 < n_kv) {
                                issue_qk(t, s1);
                                umma_commit(&s_full[t]);
                                if (t == n_qt - 1) {
                                    umma_commit(&k_empty[s1]);
                                    if (j + 2 == n_kv) umma_commit(q_empty);   // that was the item's last S
                                }
                            } else {
                                umma_commit(&o_full[t]);
                            }
                        }
                        __syncwarp();
                    }
                }
                n1 += (n_qt == 2);
            }
        }
    } else if (warp >= 8) {
        // warps 10, 11: register donors.  In the fused forward+backward call they also do the backward's
        // HBM-bound preparation in the shadow of the tensor-core loop: dO -> 16 bit (zero padded to DP) and the
        // dQ zero-fill (the reference's cudaMemset before its backward launch, f-attn2-backward.cu:427).
        setmaxnreg_dec<56>();
        if (FUSED && p.dOh != nullptr) {
            constexpr int LPR = DP / 8, RPW = 32 / LPR, U = 2;     // lanes per row, rows per warp pass, passes in flight
            const int sub = lane / LPR, l = lane % LPR, col = l * 8;
            const size_t rows = static_cast<size_t>(p.donor_rows);
            const size_t wid = static_cast<size_t>(blockIdx.x) * 2 + (warp - 10), nw = static_cast<size_t>(gridDim.x) * 2;
            uint4* dOh = static_cast<uint4*>(p.dOh);
            float4* dq = reinterpret_cast<float4*>(p.dQ_zero);
            const bool col_ok = col < p.D;
            float amax = 0.0f;                                       // max |dO| seen by this thread (range fix-up input)
            for (size_t base = wid * (RPW * U); base < rows; base += nw * (RPW * U)) {
                float4 a[U], b[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t row = base + u * RPW + sub;
                    if (row < rows && col_ok) {
                        a[u] = __ldcs(reinterpret_cast<const float4*>(p.dO + row * p.D + col));
                        b[u] = __ldcs(reinterpret_cast<const float4*>(p.dO + row * p.D + col + 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t row = base + u * RPW + sub;
                    if (row < rows) {
                        uint4 out = make_uint4(0u, 0u, 0u, 0u);
                        if (col_ok) {
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a[u].x), fabsf(a[u].y)), fmaxf(fabsf(a[u].z), fabsf(a[u].w))));
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(b[u].x), fabsf(b[u].y)), fmaxf(fabsf(b[u].z), fabsf(b[u].w))));
                            out.x = BF16 ? pack_bf16x2(a[u].x, a[u].y) : pack_half2(a[u].x, a[u].y);
                            out.y = BF16 ? pack_bf16x2(a[u].z, a[u].w) : pack_half2(a[u].z, a[u].w);
                            out.z = BF16 ? pack_bf16x2(b[u].x, b[u].y) : pack_half2(b[u].x, b[u].y);
                            out.w = BF16 ? pack_bf16x2(b[u].z, b[u].w) : pack_half2(b[u].z, b[u].w);
                            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                            __stcs(dq + ((row * p.D + col) >> 2), z);
                            __stcs(dq + ((row * p.D + col) >> 2) + 1, z);
                        }
                        __stcs(dOh + row * LPR + l, out);
                    }
                }
            }
            if (p.rb != nullptr) {
                // publish max|dO|; the last donor warp of the grid to finish decides dO's scale for the backward
                const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
                if (lane == 0) {
                    if (wmax != 0u) atomicMax(&p.rb->amax[3][wid % kAmaxLanes], wmax);
                    if (range_last_arrival(&p.rb->ticket[1], 2u * gridDim.x)) decide_do(p.rb, p.D, BF16 ? 1 : 0, p.scale);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax + epilogue
        setmaxnreg_inc<224>();
        const int t = warp >> 2;
        const int row_in_tile = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_base + lane_addr + (t == 0 ? COL_S0 : COL_S1);
        const uint32_t t_o = tmem_base + lane_addr + (t == 0 ? COL_O0 : COL_O1);
        // softmax exponent scale log2(e) / sqrt(D), divided by the power-of-two scales of the 16-bit Q / K copies
        const float c2 = p.range != nullptr ? __ldg(p.range + kC2) : p.scale_log2;
        uint8_t* stage = smem + L::OFF_STAGE + t * L::STAGE_BYTES;    // this warpgroup's 16 KB O staging buffer
        const bool issuer = (warp & 3) == 0 && lane == 0;              // owns the warpgroup's TMA store groups
        const uint32_t bar_id = 1 + 2 * t;
        const int n_chunk = (p.D + 31) / 32;                           // real columns only (D = 32 under DP = 64)

        int mine = 0;                                                   // work items this warpgroup has processed
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            if (t >= tiles_of(w)) continue;
            const int bh = w / q_blocks, q_row0 = (w % q_blocks) * (2 * BM);
            const int q_row = q_row0 + t * BM + row_in_tile;
            const int step0 = mine * n_kv;
#ifdef FA2_TIMELINE
            if (threadIdx.x == 0 && p.timeline) {
                uint32_t smid;
                asm("mov.u32 %0, %%smid;" : "=r"(smid));
                p.timeline[1024 + 8 * w + 4] = smid;
                TLC(0);
            }
#endif

            float m_ref = -INFINITY;   // running reference max (raw score units)
            float l_run = 0.0f;

            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&s_full[t], (step0 + j) & 1);
                tc_fence_after();
                if ((warp & 3) == 0) TLF(t == 0 ? 8 : 12);
                if (j == 0 && threadIdx.x == 0) TLC(1);

                uint32_t sr[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(t_s + c * 32, sr[c]);
                tmem_wait_ld();

                const int valid = p.S_kv - j * BN;  // columns >= valid are padding in the last tile
                if (valid < BN) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i >= valid) sr[c][i] = __float_as_uint(-INFINITY);
                }

                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
                    mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
                    mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
                    mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
                }
                const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_ref);

                if (j == 0) {
                    m_ref = m_new;
                } else {
                    const bool need = (m_new - m_ref) * c2 > RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, need)) {
                        // S(j) complete implies O += P V (j-1) complete (commits are ordered), so O is quiescent.
                        const float alpha = ex2_approx((m_ref - m_new) * c2);
                        m_ref = m_new;
                        l_run *= alpha;
#pragma unroll
                        for (int c = 0; c < DP / 32; ++c) {
                            uint32_t orr[32];
                            tmem_ld32(t_o + c * 32, orr);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
                            tmem_st32(t_o + c * 32, orr);
                        }
                    }
                }

                // P = 2^(S*c2 - m*c2) with packed fp32x2 math, rounded to 16 bit and written over S (all of S is
                // already in registers); handed to the MMA warp in two halves so P V can start early.
                const float neg_m = -m_ref * c2;
                const float2 c2v = make_float2(c2, c2), nmv = make_float2(neg_m, neg_m);
                float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), c2v, nmv);
                        const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sr[c][i + 2]), __uint_as_float(sr[c][i + 3])), c2v, nmv);
                        const float2 e0 = ((POLY_A >> (i >> 2)) & 1u) ? ex2_poly2(x0) : make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
                        const float2 e1 = ((POLY_B >> (i >> 2)) & 1u) ? ex2_poly2(x1) : make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                        ls0 = __fadd2_rn(ls0, e0);
                        ls1 = __fadd2_rn(ls1, e1);
                        pk[i >> 1] = BF16 ? pack_bf16x2(e0.x, e0.y) : pack_half2(e0.x, e0.y);
                        pk[(i >> 1) + 1] = BF16 ? pack_bf16x2(e1.x, e1.y) : pack_half2(e1.x, e1.y);
                    }
                    tmem_st16(t_s + c * 16, pk);
                    if (c & 1) {
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&p_full[2 * t + (c >> 1)]);
                        if ((warp & 3) == 0) TLF((t == 0 ? 9 : 13) + (c >> 1));
                    }
                }
                l_run += (ls0.x + ls0.y) + (ls1.x + ls1.y);
            }

            // epilogue: O / l -> 128B-swizzled fp32 staging tile -> TMA store (rows past S are clipped by the
            // tensor map); LSE = ln(l) + m / sqrt(D).  The MMA warp is already computing S(0) of the next item.
            // fused forward+backward: this thread also forms D_i = rowsum(dO o O) for its row (the reference's
            // D_computation_reduction_kernel, f-attn2-backward.cu:342-380) while O is in registers; the dO row is
            // fetched before the wait for the last P V so that its latency hides there.
            const bool row_ok = q_row < p.S_q;
            // (rows past S read row 0 of the slab: a valid address whose result is never stored)
            // Row-per-thread global loads cost the LSU one wavefront per lane per instruction whatever their width
            // (measured: with 128-bit loads the four chunks added 5.7K cycles to every epilogue, with 256-bit loads
            // 3.5K; tools/timeline_fwd.py ... fused), hence LDG.256.  Variants that read the staged O tile with
            // coalesced dO loads, or request dO further ahead, either stall on the proxy fence before the TMA store
            // (it waits for the thread's outstanding loads, ~2500 cycles each) or spill.
            const float* do_row = (FUSED ? p.dO : p.O) + (static_cast<size_t>(bh) * p.q_pitch + (row_ok ? q_row : 0)) * p.D;
            float dov[4][8];                                            // dO columns of the chunk being reduced
            if constexpr (FUSED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256_stream(do_row + i * 8, dov[i]);
            }
            float dsum = 0.0f;
            mbar_wait(&o_full[t], mine & 1);
            tc_fence_after();
            if (threadIdx.x == 0) TLC(2);
            // 1 / l, times the inverse of V's power-of-two scale
            const float inv_l = (p.range != nullptr ? __ldg(p.range + kInvV) : 1.0f) / l_run;
#pragma unroll
            for (int c = 0; c < DP / 32; ++c) {
                if (c < n_chunk) {
                    uint32_t orr[32];
                    tmem_ld32(t_o + c * 32, orr);
                    tmem_wait_ld();
                    if constexpr (FUSED) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) dsum = fmaf(__uint_as_float(orr[i]), dov[i >> 3][i & 7], dsum);
                        if (c + 1 < DP / 32) {
                            // next chunk's dO, in flight during the store wait (unconditional loads keep dov in
                            // registers; past the last real chunk the last one is simply fetched again)
                            const int nc = (c + 1 < n_chunk) ? c + 1 : n_chunk - 1;
#pragma unroll
                            for (int i = 0; i < 4; ++i) ldg256_stream(do_row + nc * 32 + i * 8, dov[i]);
                        }
                    }
                    if (issuer) tma_store_wait_read<0>();       // the previous store out of the buffer has been read
                    named_bar_sync(bar_id, 128);
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        float4 v4;
                        v4.x = __uint_as_float(orr[q4 * 4]) * inv_l;
                        v4.y = __uint_as_float(orr[q4 * 4 + 1]) * inv_l;
                        v4.z = __uint_as_float(orr[q4 * 4 + 2]) * inv_l;
                        v4.w = __uint_as_float(orr[q4 * 4 + 3]) * inv_l;
                        *reinterpret_cast<float4*>(stage + swz128(row_in_tile, q4)) = v4;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id + 1, 128);
                    if (issuer) {
                        tma_store_3d(&p.tm_o, stage, c * 32, q_row0 + t * BM, bh);
                        tma_store_commit();
                    }
                    if (threadIdx.x == 0 && c < 3) TLC(5 + c);
                }
            }
            if (row_ok) {
                const size_t g = static_cast<size_t>(bh) * p.q_pitch + q_row;
                const float lse = m_ref * (p.range != nullptr ? __ldg(p.range + kScaleLse) : p.scale) + logf(l_run);
                p.LSE[g] = lse;
                if (FUSED && p.delta != nullptr) {
                    p.delta[g] = dsum * inv_l;
                    p.lse_log2[g] = lse * 1.4426950408889634f;      // same rounding as the stand-alone pre-pass
                }
            }
            if (threadIdx.x == 0) TLC(3);
            ++mine;
        }
        if (issuer) tma_store_wait<0>();                        // global writes done before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

cudaError_t launch_fwd(const FwdParams& p, cudaStream_t st) {
    const int DP = padded_head_dim(p.D);
    const int q_blocks = (p.S_q + 2 * BM - 1) / (2 * BM);
    // persistent: one CTA per SM (or fewer when there is less work), each walks its share of the work items
    static int sm_count[64] = {0};
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (dev < 64 && sm_count[dev] == 0 &&
        (e0 = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
        return e0;
    const long long n_work = static_cast<long long>(p.BH) * q_blocks;
    const int n_sm = dev < 64 ? sm_count[dev] : 148;
    const dim3 grid(static_cast<unsigned>(n_work < n_sm ? n_work : n_sm));
    cudaError_t e;
    auto go = [&](auto kern, int smem) -> cudaError_t {
        cudaError_t err = ensure_smem_optin(reinterpret_cast<const void*>(kern), smem);
        if (err != cudaSuccess) return err;
        kern<<<grid, NUM_THREADS, smem, st>>>(p);
        return cudaSuccess;
    };
    const bool fused = p.dO != nullptr;
#define FA2_FWD_GO(DPV) (p.bf16 ? (fused ? go(fa2_fwd_kernel<DPV, true, true>, FwdSmem<DPV>::ALLOC)    \
                                         : go(fa2_fwd_kernel<DPV, true, false>, FwdSmem<DPV>::ALLOC))  \
                                : (fused ? go(fa2_fwd_kernel<DPV, false, true>, FwdSmem<DPV>::ALLOC)   \
                                         : go(fa2_fwd_kernel<DPV, false, false>, FwdSmem<DPV>::ALLOC)))
    e = (DP == 64) ? FA2_FWD_GO(64) : FA2_FWD_GO(128);
#undef FA2_FWD_GO
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}


// Forces the module/kernels to be loaded on the current device (lazy loading would otherwise land inside
// the first timed launch).
cudaError_t warm_fwd() {
    cudaFuncAttributes a;
    cudaError_t e;
    const void* kernels[] = {
        reinterpret_cast<const void*>(fa2_fwd_kernel<64, false, false>), reinterpret_cast<const void*>(fa2_fwd_kernel<64, false, true>),
        reinterpret_cast<const void*>(fa2_fwd_kernel<64, true, false>), reinterpret_cast<const void*>(fa2_fwd_kernel<64, true, true>),
        reinterpret_cast<const void*>(fa2_fwd_kernel<128, false, false>), reinterpret_cast<const void*>(fa2_fwd_kernel<128, false, true>),
        reinterpret_cast<const void*>(fa2_fwd_kernel<128, true, false>), reinterpret_cast<const void*>(fa2_fwd_kernel<128, true, true>)};
    for (const void* k : kernels)
        if ((e = cudaFuncGetAttributes(&a, k)) != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace fa2
