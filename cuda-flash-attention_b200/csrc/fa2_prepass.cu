// fa2_prepass.cu -- HBM-bound pre-passes around the tcgen05 kernels.
//  * cast_qkv:     fp32 [rows][D] -> 16-bit [rows][DP] (zero padded) for Q, K, V in one launch; also collects
//                  max |x| per tensor (RangeBlock::amax) for the range fix-up.
//  * bwd_prepass:  replaces the reference's D_computation_reduction_kernel
//                  (kernels/f-attn2-backward.cu:342-380: one *block* per row) with one warp
//                  per 8 rows, fused with the 16-bit cast of dO (+ its max |x|), LSE -> log2 domain and the
//                  dQ zero-fill the reference does with cudaMemset (f-attn2-backward.cu:427).
//  * range_fix_*:  the reference's fp32 kernels take any fp32 input (kernel_fa2_optimized.cu:19-347,
//                  f-attn2-backward.cu:243-266); fp16 operands do not: max 65504, precision loss below 6e-5.
//                  These kernels read the maxima, choose power-of-two scales so that every 16-bit copy and the
//                  16-bit (dP - D_i) sit well inside the fp16 range, re-cast the tensors that need a scale other than
//                  1 (rare; otherwise they return after a few loads) and publish the factors the main kernels fold
//                  into the softmax scale and their epilogues.  Powers of two: the result is what exact scaling gives.
#include "fa2_common.h"
#include "fa2_range.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fa2 {
namespace {

__device__ __forceinline__ uint32_t pack16(float lo, float hi, int bf16) {
    uint32_t r;
    if (bf16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// max |x| of a block -> one atomicMax on one of the tensor's kAmaxLanes words (float bits of non-negative floats
// order like unsigned integers; NaNs are ignored by fmaxf, an inf makes the fix-up leave the tensor alone).
__device__ __forceinline__ void block_amax(float m, unsigned* lanes) {
    __shared__ unsigned warp_max[32];
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned b = 0;
        for (unsigned i = 0; i < (blockDim.x + 31) / 32; ++i) b = warp_max[i] > b ? warp_max[i] : b;
        unsigned* dst = lanes + (blockIdx.x % kAmaxLanes);
        if (b > *reinterpret_cast<volatile unsigned*>(dst)) atomicMax(dst, b);
    }
}

// Each thread converts 8 consecutive output elements (one uint4 store).  grid.y = tensor.
// SCALED = the fix-up's re-cast (per-tensor scale, tensors with scale 1 are skipped), else the first pass (amax).
template <bool SCALED>
__device__ __forceinline__ void cast_tensor(const float* __restrict__ src, uint4* __restrict__ dst, size_t rows, int D,
                                            int DP, int bf16, float s, unsigned* amax_lanes) {
    const int vec_per_row = DP >> 3;
    const size_t total = rows * vec_per_row;
    float m = 0.0f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t row = i / vec_per_row;
        const int col = static_cast<int>(i % vec_per_row) * 8;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (col < D) {
            float4 a = __ldg(reinterpret_cast<const float4*>(src + row * D + col));
            float4 b = __ldg(reinterpret_cast<const float4*>(src + row * D + col + 4));
            if (SCALED) {
                a.x *= s; a.y *= s; a.z *= s; a.w *= s; b.x *= s; b.y *= s; b.z *= s; b.w *= s;
            } else {
                m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
                m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
            }
            out.x = pack16(a.x, a.y, bf16);
            out.y = pack16(a.z, a.w, bf16);
            out.z = pack16(b.x, b.y, bf16);
            out.w = pack16(b.z, b.w, bf16);
        }
        dst[i] = out;
    }
    if (!SCALED && amax_lanes != nullptr) block_amax(m, amax_lanes);
}

__global__ void __launch_bounds__(256)
cast_qkv_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                uint4* __restrict__ Qh, uint4* __restrict__ Kh, uint4* __restrict__ Vh,
                size_t rows, int D, int DP, int bf16, RangeBlock* rb, float scale, float scale_log2) {
    const float* src = blockIdx.y == 0 ? Q : (blockIdx.y == 1 ? K : V);
    uint4* dst = blockIdx.y == 0 ? Qh : (blockIdx.y == 1 ? Kh : Vh);
    cast_tensor<false>(src, dst, rows, D, DP, bf16, 1.0f, rb ? rb->amax[blockIdx.y] : nullptr);
    // the last block to finish turns the maxima into this launch's scales
    if (rb != nullptr && threadIdx.x == 0 && range_last_arrival(&rb->ticket[0], gridDim.x * gridDim.y))
        decide_qkv(rb, bf16, scale, scale_log2);
}

__global__ void __launch_bounds__(256)
range_fix_qkv_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                     uint4* __restrict__ Qh, uint4* __restrict__ Kh, uint4* __restrict__ Vh,
                     size_t rows, int D, int DP, int bf16, const RangeBlock* rb) {
    const float s = __ldcg(rb->sc + kSq + blockIdx.y);           // kSq, kSk, kSv are consecutive
    if (s == 1.0f) return;                                       // the usual case: nothing to redo
    const float* src = blockIdx.y == 0 ? Q : (blockIdx.y == 1 ? K : V);
    uint4* dst = blockIdx.y == 0 ? Qh : (blockIdx.y == 1 ? Kh : Vh);
    cast_tensor<true>(src, dst, rows, D, DP, bf16, s, nullptr);
}

__global__ void __launch_bounds__(256)
range_fix_do_kernel(const float* __restrict__ dO, uint4* __restrict__ dOh, size_t rows, int D, int DP, int bf16,
                    const RangeBlock* rb) {
    const float sdo = __ldcg(rb->sc + kSdo);
    if (sdo == 1.0f) return;
    cast_tensor<true>(dO, dOh, rows, D, DP, bf16, sdo, nullptr);
}

// Small problems (at most one 8-element vector per thread and tensor on a full grid): amax, scale decision and cast
// in ONE cooperative launch that reads every input exactly once.  Each thread fetches its vector of all (3 or 4)
// tensors up front -- one DRAM latency for everything -- and keeps the fp32 values in registers across the grid
// barrier (arrive counter; block 0 decides the scales, clears the maxima and releases a generation); then it scales,
// rounds and stores, and with dO given zero-fills its piece of dQ.  Saves the two always-launched re-cast kernels of
// the large path, which would cost more than the cast itself here.
struct SmallCastArgs {
    const float* src[4];
    uint4* dst[4];
    float4* dq_zero;
    size_t rows;
    int D, DP, bf16, nt;
    RangeBlock* rb;
    float scale, scale_log2;
};
constexpr int kSmallThreads = 1024;

__global__ void __launch_bounds__(kSmallThreads, 1)
cast_small_kernel(const SmallCastArgs a) {
    __shared__ unsigned s_max[4][kSmallThreads / 32];
    __shared__ float s_scale[4];
    const unsigned gen0 = *reinterpret_cast<volatile unsigned*>(&a.rb->coop_release);   // read before arriving
    const int vec_per_row = a.DP >> 3;
    const size_t total = a.rows * vec_per_row;
    const size_t i = blockIdx.x * static_cast<size_t>(kSmallThreads) + threadIdx.x;
    const bool in = i < total;
    const size_t row = in ? i / vec_per_row : 0;
    const int col = in ? static_cast<int>(i % vec_per_row) * 8 : 0;
    const bool real = in && col < a.D;                       // (padding columns of D = 32 are written as zeros)
    float4 x[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        x[t][0] = x[t][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < a.nt && real) {
            x[t][0] = __ldg(reinterpret_cast<const float4*>(a.src[t] + row * a.D + col));
            x[t][1] = __ldg(reinterpret_cast<const float4*>(a.src[t] + row * a.D + col + 4));
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float4 p = x[t][0], q = x[t][1];
        float m = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fmaxf(fabsf(p.z), fabsf(p.w)));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))));
        const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if ((threadIdx.x & 31) == 0) s_max[t][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 0; t < a.nt; ++t) {
            unsigned b = 0;
            for (int w = 0; w < kSmallThreads / 32; ++w) b = s_max[t][w] > b ? s_max[t][w] : b;
            if (b != 0u) atomicMax(&a.rb->amax[t][blockIdx.x % kAmaxLanes], b);
        }
        __threadfence();
        atomicAdd(&a.rb->coop_arrive, 1u);
        if (blockIdx.x == 0) {
            while (*reinterpret_cast<volatile unsigned*>(&a.rb->coop_arrive) != gridDim.x) { }
            a.rb->coop_arrive = 0u;
            __threadfence();
            decide_qkv(a.rb, a.bf16, a.scale, a.scale_log2);
            if (a.nt == 4) decide_do(a.rb, a.D, a.bf16, a.scale);
            __threadfence();
            atomicAdd(&a.rb->coop_release, 1u);
        } else {
            while (*reinterpret_cast<volatile unsigned*>(&a.rb->coop_release) == gen0) { }
        }
        __threadfence();
        for (int t = 0; t < 4; ++t) s_scale[t] = __ldcg(a.rb->sc + kSq + t);       // kSq, kSk, kSv, kSdo
    }
    __syncthreads();
    if (!in) return;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t >= a.nt) break;
        const float s = s_scale[t];
        const float4 p = x[t][0], q = x[t][1];
        uint4 out;                                            // (zeros stay zeros: the padding columns)
        out.x = pack16(p.x * s, p.y * s, a.bf16);
        out.y = pack16(p.z * s, p.w * s, a.bf16);
        out.z = pack16(q.x * s, q.y * s, a.bf16);
        out.w = pack16(q.z * s, q.w * s, a.bf16);
        a.dst[t][i] = out;
    }
    if (a.nt == 4 && real) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        a.dq_zero[(row * a.D + col) >> 2] = z;
        a.dq_zero[((row * a.D + col) >> 2) + 1] = z;
    }
}

// One row per group of (D/8 <= 16) lanes: each lane handles 8 elements of the row.
// PARTS bit 0: cast dO to 16 bit and zero-fill dQ (independent of the forward's outputs);
// PARTS bit 1: D_i = rowsum(dO o O) and LSE -> log2 domain (need O / LSE).
template <int LANES_PER_ROW, int PARTS>
__global__ void __launch_bounds__(256)
bwd_prepass_kernel(const float* __restrict__ O, const float* __restrict__ dO, const float* __restrict__ LSE,
                   uint4* __restrict__ dOh, float* __restrict__ delta, float* __restrict__ lse_log2,
                   float4* __restrict__ dQ, size_t rows, int D, int DP, int bf16, RangeBlock* rb, float scale,
                   unsigned range_rows, unsigned pitch_rows) {
    // rows are counted over a row range of every slab: range_rows consecutive rows, slabs pitch_rows apart (the
    // sequence-split path; pointers come pre-offset to the first row of the range).  Normally range_rows == pitch_rows.
    constexpr int ROWS_PER_WARP = 32 / LANES_PER_ROW;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES_PER_ROW;          // which row of the warp's group
    const int l = lane % LANES_PER_ROW;            // position within the row
    const size_t warp_global = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    const size_t n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const int vec_per_row = DP >> 3;
    float m = 0.0f;
    for (size_t base = warp_global * ROWS_PER_WARP; base < rows; base += n_warps * ROWS_PER_WARP) {
        const size_t lrow = base + sub;
        const size_t row = (range_rows == pitch_rows) ? lrow : (lrow / range_rows) * pitch_rows + lrow % range_rows;
        float acc = 0.f;
        if (lrow < rows) {
            const int col = l * 8;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (col < D) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(dO + row * D + col));
                const float4 b = __ldg(reinterpret_cast<const float4*>(dO + row * D + col + 4));
                if (PARTS & 2) {
                    const float4 oa = __ldg(reinterpret_cast<const float4*>(O + row * D + col));
                    const float4 ob = __ldg(reinterpret_cast<const float4*>(O + row * D + col + 4));
                    acc = a.x * oa.x + a.y * oa.y + a.z * oa.z + a.w * oa.w + b.x * ob.x + b.y * ob.y + b.z * ob.z +
                          b.w * ob.w;
                }
                if (PARTS & 1) {
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
                    out.x = pack16(a.x, a.y, bf16);
                    out.y = pack16(a.z, a.w, bf16);
                    out.z = pack16(b.x, b.y, bf16);
                    out.w = pack16(b.z, b.w, bf16);
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    dQ[(row * D + col) >> 2] = z;
                    dQ[((row * D + col) >> 2) + 1] = z;
                }
            }
            if ((PARTS & 1) && l < vec_per_row) dOh[row * vec_per_row + l] = out;
        }
        if (PARTS & 2) {
#pragma unroll
            for (int off = LANES_PER_ROW / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lrow < rows && l == 0) {
                delta[row] = acc;
                lse_log2[row] = LSE[row] * 1.4426950408889634f;
            }
        }
    }
    if ((PARTS & 1) && rb != nullptr) {
        block_amax(m, rb->amax[3]);
        if (threadIdx.x == 0 && range_last_arrival(&rb->ticket[1], gridDim.x)) decide_do(rb, D, bf16, scale);
    }
}

// Sequence-split backward: every GPU of a group holds a partial dQ (its KV range's contribution to ALL query rows); the
// owner of a query-row range adds the other partials' rows to its own, reading them straight out of the peers' memory
// (NVLink / NVSwitch P2P loads; the owner's range is read by nobody else, so the sum is formed in place).
struct PeerPtrs { const float4* p[8]; };
__global__ void __launch_bounds__(256)
dq_peer_reduce_kernel(float4* __restrict__ own, PeerPtrs peers, int n_peers, size_t seg_vec, size_t pitch_vec, int cnt) {
    const size_t total = seg_vec * cnt;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t idx = (i / seg_vec) * pitch_vec + i % seg_vec;
        float4 acc = own[idx];
        for (int q = 0; q < n_peers; ++q) {
            const float4 v = __ldcg(peers.p[q] + idx);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        own[idx] = acc;
    }
}

unsigned cast_blocks(size_t rows, int DP) {
    const size_t total = rows * (DP >> 3);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;     // grid-stride, a multiple of the SM count
    if (blocks == 0) blocks = 1;
    return static_cast<unsigned>(blocks);
}

// the re-cast kernels usually return at once: a modest grid keeps the always-paid launch cheap
unsigned fix_blocks(size_t rows, int DP) {
    const unsigned b = cast_blocks(rows, DP);
    return b > 148 * 4 ? 148 * 4 : b;
}

}  // namespace

cudaError_t launch_cast_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                            size_t rows, int D, int DP, int bf16, RangeBlock* rb, float scale,
                            float scale_log2, cudaStream_t st) {
    cast_qkv_kernel<<<dim3(cast_blocks(rows, DP), 3), 256, 0, st>>>(
        Q, K, V, static_cast<uint4*>(Qh), static_cast<uint4*>(Kh), static_cast<uint4*>(Vh), rows, D, DP, bf16, rb,
        scale, scale_log2);
    return cudaGetLastError();
}

cudaError_t launch_cast_small(const float* Q, const float* K, const float* V, const float* dO, void* Qh, void* Kh,
                              void* Vh, void* dOh, float* dQ_zero, size_t rows, int D, int DP, int bf16, RangeBlock* rb,
                              float scale, float scale_log2, int n_sm, cudaStream_t st) {
    SmallCastArgs a{};
    a.src[0] = Q; a.src[1] = K; a.src[2] = V; a.src[3] = dO;
    a.dst[0] = static_cast<uint4*>(Qh); a.dst[1] = static_cast<uint4*>(Kh); a.dst[2] = static_cast<uint4*>(Vh);
    a.dst[3] = static_cast<uint4*>(dOh);
    a.dq_zero = reinterpret_cast<float4*>(dQ_zero);
    a.rows = rows; a.D = D; a.DP = DP; a.bf16 = bf16; a.nt = dO ? 4 : 3; a.rb = rb; a.scale = scale; a.scale_log2 = scale_log2;
    const size_t total = rows * (DP >> 3);
    size_t blocks = (total + kSmallThreads - 1) / kSmallThreads;   // one vector per thread and tensor
    if (blocks > static_cast<size_t>(n_sm)) return cudaErrorInvalidValue;   // (the caller tests cast_small_fits)
    if (blocks == 0) blocks = 1;
    void* args[] = {&a};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(cast_small_kernel), dim3(static_cast<unsigned>(blocks)),
                                       dim3(kSmallThreads), args, 0, st);
}

// one block per SM at most (co-resident for the grid barrier), one 8-element vector per thread and tensor
bool cast_small_fits(size_t rows, int DP, int n_sm) {
    return rows * static_cast<size_t>(DP >> 3) <= static_cast<size_t>(n_sm) * kSmallThreads;
}

cudaError_t launch_dq_peer_reduce(float* own, const float* const* peers, int n_peers, size_t seg_floats,
                                  size_t pitch_floats, int cnt, cudaStream_t st) {
    if (n_peers <= 0 || cnt <= 0 || seg_floats == 0) return cudaSuccess;
    if (n_peers > 8 || (seg_floats & 3) || (pitch_floats & 3)) return cudaErrorInvalidValue;
    PeerPtrs pp{};
    for (int i = 0; i < n_peers; ++i) pp.p[i] = reinterpret_cast<const float4*>(peers[i]);
    const size_t total = (seg_floats >> 2) * cnt;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    dq_peer_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<float4*>(own), pp, n_peers,
                                                                         seg_floats >> 2, pitch_floats >> 2, cnt);
    return cudaGetLastError();
}

cudaError_t launch_range_fix_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                                 size_t rows, int D, int DP, int bf16, const RangeBlock* rb, cudaStream_t st) {
    range_fix_qkv_kernel<<<dim3(fix_blocks(rows, DP), 3), 256, 0, st>>>(
        Q, K, V, static_cast<uint4*>(Qh), static_cast<uint4*>(Kh), static_cast<uint4*>(Vh), rows, D, DP, bf16, rb);
    return cudaGetLastError();
}

cudaError_t launch_range_fix_do(const float* dO, void* dOh, size_t rows, int D, int DP, int bf16, const RangeBlock* rb,
                                cudaStream_t st) {
    range_fix_do_kernel<<<fix_blocks(rows, DP), 256, 0, st>>>(dO, static_cast<uint4*>(dOh), rows, D, DP, bf16, rb);
    return cudaGetLastError();
}

cudaError_t launch_bwd_prepass(const float* O, const float* dO, const float* LSE, void* dOh, float* delta,
                               float* lse_log2, float* dQ_zero, size_t rows, int D, int DP, int bf16, int parts,
                               RangeBlock* rb, float scale, cudaStream_t st, unsigned range_rows, unsigned pitch_rows) {
    const int lanes = DP >> 3;                    // 8 (DP = 64) or 16 (DP = 128)
    const size_t rows_per_block = (256 / 32) * (32 / lanes);
    size_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks == 0) blocks = 1;
    const unsigned g = static_cast<unsigned>(blocks);
    uint4* dh = static_cast<uint4*>(dOh);
    float4* dq = reinterpret_cast<float4*>(dQ_zero);
#define FA2_PRE(L, P) bwd_prepass_kernel<L, P><<<g, 256, 0, st>>>(O, dO, LSE, dh, delta, lse_log2, dq, rows, D, DP, bf16, rb, scale, range_rows, pitch_rows)
    if (lanes == 8) { if (parts == 1) FA2_PRE(8, 1); else if (parts == 2) FA2_PRE(8, 2); else FA2_PRE(8, 3); }
    else            { if (parts == 1) FA2_PRE(16, 1); else if (parts == 2) FA2_PRE(16, 2); else FA2_PRE(16, 3); }
#undef FA2_PRE
    return cudaGetLastError();
}

}  // namespace fa2
