// fa2_bwd_sm100.cu -- FlashAttention-2 backward for sm_100a (replaces the reference's
// flash_attention2_backward_kernel, kernels/f-attn2-backward.cu:33-339).
//
// Like the reference, one CTA owns one KV tile of a (batch, head) slab and walks the Q tiles,
// keeping dK / dV on chip and adding its dQ contribution into global memory
// (reference: atomicAdd, f-attn2-backward.cu:298; here: TMA reduce-add in L2).  Tiles are
// 128x128 and the five products run on the tensor cores, transposed so that the KV row is
// the TMEM lane:
//     S^T  = K Q^T          (SS, both K-major)            -> TMEM [384,512)
//     dP^T = V dO^T         (SS, both K-major)            -> TMEM [256,384)
//     dV  += P^T dO         (A = P^T from TMEM, B = dO tile as MN-major)   -> TMEM [128,128+DP)
//     dK  += dS^T Q         (A = dS^T smem K-major,  B = Q tile MN-major)  -> TMEM [0,DP)
//     dQ   = dS K           (A = dS^T smem read MN-major, B = K tile MN-major) -> TMEM [256,256+DP)
// dQ shares its TMEM columns with dP^T (dP is dead once dS has been formed).
// Warp roles (14 warps):  0-3 / 4-7 compute (each warpgroup owns 64 of the 128 Q columns),
// 8-11 dQ drain (TMEM -> swizzled smem -> cp.reduce.async.bulk.tensor add), 12 MMA issuer,
// 13 TMA producer (also stages LSE and D_i per Q tile).
// Padding needs no masks: TMA zero-fills out-of-range Q/K/V/dO rows, out-of-range LSE is
// staged as +inf (P = 0), and out-of-range dK/dV rows are not stored.
#include "fa2_common.h"
#include "ptx.cuh"

namespace fa2 {
namespace {

constexpr int BT = 128;                 // tile rows (both KV and Q)
constexpr int ATOM = BT * 128;          // bytes of one [128 rows][64 x 16-bit] swizzle atom
constexpr int NUM_THREADS = 448;
constexpr int DRAIN_WARP0 = 8;
constexpr int MMA_WARP = 12;
constexpr int TMA_WARP = 13;
constexpr int Q_STAGES = 2;

template <int DP>
struct BwdSmem {
    static constexpr int TILE = BT * DP * 2;
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = OFF_K + TILE;
    static constexpr int OFF_Q = OFF_V + TILE;                    // Q_STAGES tiles
    static constexpr int OFF_DO = OFF_Q + Q_STAGES * TILE;        // 1 tile
    static constexpr int OFF_DS = OFF_DO + TILE;                  // [128 kv][128 q] 16-bit, 2 atoms
    static constexpr int OFF_DQS = OFF_DS + 2 * ATOM;             // 2 x [128][32] fp32 staging
    static constexpr int OFF_LSE = OFF_DQS + 2 * BT * 128;        // Q_STAGES x 128 fp32
    static constexpr int OFF_DELTA = OFF_LSE + Q_STAGES * BT * 4;
    static constexpr int OFF_BAR = OFF_DELTA + Q_STAGES * BT * 4;
    static constexpr int NUM_BARS = 16;
    static constexpr int OFF_TMEM_PTR = OFF_BAR + NUM_BARS * 8;
    static constexpr int BYTES = OFF_TMEM_PTR + 16;
};

template <int DP>
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
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
    float* lse_s = reinterpret_cast<float*>(smem + L::OFF_LSE);
    float* delta_s = reinterpret_cast<float*>(smem + L::OFF_DELTA);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tiles = (p.S + BT - 1) / BT;       // same count for KV and Q tiles
    const int bh = blockIdx.x / n_tiles;
    const int kv_row0 = (blockIdx.x % n_tiles) * BT;

    if (warp == TMA_WARP && lane == 0) {
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
        tma_prefetch_desc(&p.tm_do);
        tma_prefetch_desc(&p.tm_dq);
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
            mbar_init(ds_full, 8);
            mbar_init(ds_empty, 1);
            mbar_init(dq_full, 1);
            mbar_init(dq_empty, 4);           // one arrive per drain warp
            mbar_init(dkdv_full, 1);
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
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * L::TILE);
            for (int a = 0; a < DP / 64; ++a) {
                tma_load_3d(smem + L::OFF_K + a * ATOM, &p.tm_k, kv_full, a * 64, kv_row0, bh);
                tma_load_3d(smem + L::OFF_V + a * ATOM, &p.tm_v, kv_full, a * 64, kv_row0, bh);
            }
        }
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % Q_STAGES;
            const uint32_t ph = (i / Q_STAGES) & 1;
            mbar_wait(&q_empty[s], ph ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&q_full[s], L::TILE);
                for (int a = 0; a < DP / 64; ++a)
                    tma_load_3d(smem + L::OFF_Q + s * L::TILE + a * ATOM, &p.tm_q, &q_full[s], a * 64, i * BT, bh);
            }
            // stage LSE (log2 domain) and D_i of this Q tile; rows past S: +inf / 0  => P = 0, dS = 0
#pragma unroll
            for (int r = 0; r < BT / 32; ++r) {
                const int m = r * 32 + lane;
                const int row = i * BT + m;
                const bool ok = row < p.S;
                const size_t g = static_cast<size_t>(bh) * p.S + (ok ? row : 0);
                lse_s[s * BT + m] = ok ? __ldg(p.lse_log2 + g) : INFINITY;
                delta_s[s * BT + m] = ok ? __ldg(p.delta + g) : 0.0f;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&q_full[s]);
                mbar_wait(do_empty, (i & 1) ^ 1);
                mbar_expect_tx(do_full, L::TILE);
                for (int a = 0; a < DP / 64; ++a)
                    tma_load_3d(smem + L::OFF_DO + a * ATOM, &p.tm_do, do_full, a * 64, i * BT, bh);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t id_ss = umma_idesc_f16(BT, BT, 0, 0, p.bf16);    // S^T, dP^T
            const uint32_t id_kmn = umma_idesc_f16(BT, DP, 0, 1, p.bf16);   // dV, dK : A K-major, B MN-major
            const uint32_t id_mnmn = umma_idesc_f16(BT, DP, 1, 1, p.bf16);  // dQ     : A MN-major, B MN-major
            const uint32_t k_addr = smem_u32(smem + L::OFF_K);
            const uint32_t v_addr = smem_u32(smem + L::OFF_V);
            const uint32_t q_addr = smem_u32(smem + L::OFF_Q);
            const uint32_t do_addr = smem_u32(smem + L::OFF_DO);
            const uint32_t ds_addr = smem_u32(smem + L::OFF_DS);
            const uint32_t tS = tmem_base + COL_S, tDP = tmem_base + COL_DP, tDQ = tmem_base + COL_DQ;
            const uint32_t tDK = tmem_base + COL_DK, tDV = tmem_base + COL_DV;

            auto kmaj = [](uint32_t base, int k) {       // K-major operand, K-step k
                return umma_smem_desc(base + (k >> 2) * ATOM + (k & 3) * 32, 16, 1024);
            };
            auto mnmaj = [](uint32_t base, int k) {      // MN-major operand, K-step k (16 rows of 128 B)
                return umma_smem_desc(base + k * 2048, ATOM, 1024);
            };
            auto issue_s = [&](int st) {
#pragma unroll
                for (int k = 0; k < KSTEPS_D; ++k)
                    umma_ss(tS, kmaj(k_addr, k), kmaj(q_addr + st * L::TILE, k), id_ss, k > 0);
            };
            auto issue_dp = [&]() {
#pragma unroll
                for (int k = 0; k < KSTEPS_D; ++k) umma_ss(tDP, kmaj(v_addr, k), kmaj(do_addr, k), id_ss, k > 0);
            };
            auto issue_dv = [&](bool first) {
#pragma unroll
                for (int k = 0; k < KSTEPS_T; ++k)     // P^T lives in two 32-column runs of the S region
                    umma_ts(tDV, tS + (k >> 2) * 64 + (k & 3) * 8, mnmaj(do_addr, k), id_kmn,
                            (!first || k > 0) ? 1u : 0u);
            };
            auto issue_dk = [&](int st, bool first) {
#pragma unroll
                for (int k = 0; k < KSTEPS_T; ++k)
                    umma_ss(tDK, kmaj(ds_addr, k), mnmaj(q_addr + st * L::TILE, k), id_kmn,
                            (!first || k > 0) ? 1u : 0u);
            };
            auto issue_dq = [&]() {
#pragma unroll
                for (int k = 0; k < KSTEPS_T; ++k) umma_ss(tDQ, mnmaj(ds_addr, k), mnmaj(k_addr, k), id_mnmn, k > 0);
            };

            mbar_wait(kv_full, 0);
            for (int i = 0; i < n_tiles; ++i) {
                const int st = i % Q_STAGES;
                mbar_wait(&q_full[st], (i / Q_STAGES) & 1);
                tc_fence_after();
                issue_s(st);                                   // S region is free: dV(i-1) was issued before
                umma_commit(s_full);
                if (i > 0) {
                    mbar_wait(ds_full, (i - 1) & 1);
                    tc_fence_after();
                    issue_dq();                                // dQ(i-1) over the dead dP(i-1)
                    umma_commit(dq_full);
                    issue_dk((i - 1) % Q_STAGES, i == 1);      // dK += dS(i-1)^T Q(i-1)
                    umma_commit(&q_empty[(i - 1) % Q_STAGES]);
                    umma_commit(ds_empty);
                    mbar_wait(dq_empty, (i - 1) & 1);          // dQ(i-1) drained out of TMEM
                }
                mbar_wait(do_full, i & 1);
                tc_fence_after();
                issue_dp();
                umma_commit(dp_full);
                mbar_wait(p_full, i & 1);
                tc_fence_after();
                issue_dv(i == 0);
                umma_commit(do_empty);
            }
            {
                const int i = n_tiles - 1;
                mbar_wait(ds_full, i & 1);
                tc_fence_after();
                issue_dk(i % Q_STAGES, i == 0);
                issue_dq();
                umma_commit(dq_full);
                umma_commit(dkdv_full);
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ compute: P^T and dS^T
        const int h = warp >> 2;                               // which 64 Q-columns of the tile
        const int n = (warp & 3) * 32 + lane;                  // kv row within the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + COL_S + h * 64;
        const uint32_t tDP = tmem_base + lane_addr + COL_DP + h * 64;
        const float c2 = p.scale_log2;
        uint8_t* ds_atom = smem + L::OFF_DS + h * ATOM;        // Q columns [64h, 64h+64) = swizzle atom h

        for (int i = 0; i < n_tiles; ++i) {
            const int st = i % Q_STAGES;
            const float* lse_t = lse_s + st * BT + h * 64;
            const float* dl_t = delta_s + st * BT + h * 64;
            mbar_wait(&q_full[st], (i / Q_STAGES) & 1);        // LSE / D_i staging visible
            mbar_wait(s_full, i & 1);
            tc_fence_after();
            uint32_t pk[32];                                    // P^T row, 64 values packed 2 x 16 bit
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint32_t sr[32];
                tmem_ld32(tS + sub * 32, sr);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(lse_t + sub * 32 + c);
                    const float e0 = ex2_approx(fmaf(__uint_as_float(sr[c]), c2, -l4.x));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(sr[c + 1]), c2, -l4.y));
                    const float e2 = ex2_approx(fmaf(__uint_as_float(sr[c + 2]), c2, -l4.z));
                    const float e3 = ex2_approx(fmaf(__uint_as_float(sr[c + 3]), c2, -l4.w));
                    pk[sub * 16 + (c >> 1)] = p.bf16 ? pack_bf16x2(e0, e1) : pack_half2(e0, e1);
                    pk[sub * 16 + (c >> 1) + 1] = p.bf16 ? pack_bf16x2(e2, e3) : pack_half2(e2, e3);
                }
            }
            tmem_st32(tS, pk);                                  // over the S columns this thread already consumed
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);

            mbar_wait(dp_full, i & 1);
            tc_fence_after();
            if (i > 0) mbar_wait(ds_empty, (i - 1) & 1);        // dK(i-1), dQ(i-1) finished reading dS smem
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint32_t dr[32];
                tmem_ld32(tDP + sub * 32, dr);
                tmem_wait_ld();
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {               // 8 columns -> one 16-byte chunk of dS^T
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c8 * 8 + u * 2;
                        const uint32_t pp = pk[sub * 16 + (c >> 1)];
                        float p0, p1;
                        if (p.bf16) {
                            p0 = __uint_as_float(pp << 16);
                            p1 = __uint_as_float(pp & 0xffff0000u);
                        } else {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pp));
                            p0 = f.x; p1 = f.y;
                        }
                        const float2 d2 = *reinterpret_cast<const float2*>(dl_t + sub * 32 + c);
                        const float s0 = p0 * (__uint_as_float(dr[c]) - d2.x);
                        const float s1 = p1 * (__uint_as_float(dr[c + 1]) - d2.y);
                        w[u] = p.bf16 ? pack_bf16x2(s0, s1) : pack_half2(s0, s1);
                    }
                    *reinterpret_cast<uint4*>(ds_atom + swz128(n, sub * 4 + c8)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            tc_fence_before();              // dP reads are complete before dQ may overwrite the columns
            fence_proxy_async_smem();       // dS smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_full);
        }

        // epilogue: warpgroup 0 stores dK (scaled by 1/sqrt(D)), warpgroup 1 stores dV
        mbar_wait(dkdv_full, 0);
        tc_fence_after();
        const int row = kv_row0 + n;
        const bool row_ok = row < p.S;
        float* dst = (h == 0 ? p.dK : p.dV) + (static_cast<size_t>(bh) * p.S + (row_ok ? row : 0)) * p.D;
        const float mul = (h == 0) ? p.scale : 1.0f;
        const uint32_t tsrc = tmem_base + lane_addr + (h == 0 ? COL_DK : COL_DV);
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tsrc + c * 32, r);
            tmem_wait_ld();
            if (row_ok && c * 32 < p.D) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 v4;
                    v4.x = __uint_as_float(r[i]) * mul;
                    v4.y = __uint_as_float(r[i + 1]) * mul;
                    v4.z = __uint_as_float(r[i + 2]) * mul;
                    v4.w = __uint_as_float(r[i + 3]) * mul;
                    *reinterpret_cast<float4*>(dst + c * 32 + i) = v4;
                }
            }
        }
    } else if (warp < 12) {
        // ------------------------------------------------------------------ dQ drain
        const int wq = warp - DRAIN_WARP0;
        const int m = wq * 32 + lane;                          // Q row within the tile == TMEM lane
        const uint32_t tDQ = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + COL_DQ;
        const bool issuer = (warp == DRAIN_WARP0 && lane == 0);
        constexpr int NCHUNK = DP / 32;
        const int n_chunk = (p.D + 31) / 32;                   // real columns only (D = 32 under DP = 64)
        uint32_t g = 0;                                         // staging buffers handed to TMA so far
        for (int i = 0; i < n_tiles; ++i) {
            mbar_wait(dq_full, i & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                uint32_t r[32];
                tmem_ld32(tDQ + c * 32, r);
                tmem_wait_ld();
                if (c == NCHUNK - 1) {                          // last TMEM read of this tile: release dP/dQ columns
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dq_empty);
                }
                if (c < n_chunk) {
                    uint8_t* stage = smem + L::OFF_DQS + (g & 1) * (BT * 128);
                    if (g >= 2) {                               // buffer was handed to TMA two chunks ago
                        if (issuer) tma_store_wait_read<1>();
                        named_bar_sync(1, 128);
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        float4 v4;
                        v4.x = __uint_as_float(r[q4 * 4]) * p.scale;
                        v4.y = __uint_as_float(r[q4 * 4 + 1]) * p.scale;
                        v4.z = __uint_as_float(r[q4 * 4 + 2]) * p.scale;
                        v4.w = __uint_as_float(r[q4 * 4 + 3]) * p.scale;
                        *reinterpret_cast<float4*>(stage + swz128(m, q4)) = v4;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (issuer) {
                        tma_reduce_add_3d(&p.tm_dq, stage, c * 32, i * BT, bh);
                        tma_store_commit();
                    }
                    ++g;
                }
            }
        }
        if (issuer) tma_store_wait<0>();
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
    const int DP = padded_head_dim(p.D);
    const int n_tiles = (p.S + BT - 1) / BT;
    const dim3 grid(static_cast<unsigned>(p.BH) * n_tiles);
    cudaError_t e;
    if (DP == 64) {
        e = cudaFuncSetAttribute(fa2_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<64>::BYTES);
        if (e != cudaSuccess) return e;
        fa2_bwd_kernel<64><<<grid, NUM_THREADS, BwdSmem<64>::BYTES, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(fa2_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 BwdSmem<128>::BYTES);
        if (e != cudaSuccess) return e;
        fa2_bwd_kernel<128><<<grid, NUM_THREADS, BwdSmem<128>::BYTES, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace fa2
