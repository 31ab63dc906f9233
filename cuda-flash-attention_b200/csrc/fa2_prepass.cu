// fa2_prepass.cu -- HBM-bound pre-passes around the tcgen05 kernels.
//  * cast_qkv:     fp32 [rows][D] -> 16-bit [rows][DP] (zero padded) for Q, K, V in one launch.
//  * bwd_prepass:  replaces the reference's D_computation_reduction_kernel
//                  (kernels/f-attn2-backward.cu:342-380: one *block* per row) with one warp
//                  per 8 rows, fused with the 16-bit cast of dO, LSE -> log2 domain and the
//                  dQ zero-fill the reference does with cudaMemset (f-attn2-backward.cu:427).
#include "fa2_common.h"
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

// Each thread converts 8 consecutive output elements (one uint4 store).  grid.y = tensor.
__global__ void __launch_bounds__(256)
cast_qkv_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                uint4* __restrict__ Qh, uint4* __restrict__ Kh, uint4* __restrict__ Vh,
                size_t rows, int D, int DP, int bf16) {
    const float* src = blockIdx.y == 0 ? Q : (blockIdx.y == 1 ? K : V);
    uint4* dst = blockIdx.y == 0 ? Qh : (blockIdx.y == 1 ? Kh : Vh);
    const int vec_per_row = DP >> 3;
    const size_t total = rows * vec_per_row;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t row = i / vec_per_row;
        const int col = static_cast<int>(i % vec_per_row) * 8;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (col < D) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + row * D + col));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + row * D + col + 4));
            out.x = pack16(a.x, a.y, bf16);
            out.y = pack16(a.z, a.w, bf16);
            out.z = pack16(b.x, b.y, bf16);
            out.w = pack16(b.z, b.w, bf16);
        }
        dst[i] = out;
    }
}

// One row per group of (D/8 <= 16) lanes: each lane handles 8 elements of the row.
// PARTS bit 0: cast dO to 16 bit and zero-fill dQ (independent of the forward's outputs);
// PARTS bit 1: D_i = rowsum(dO o O) and LSE -> log2 domain (need O / LSE).
template <int LANES_PER_ROW, int PARTS>
__global__ void __launch_bounds__(256)
bwd_prepass_kernel(const float* __restrict__ O, const float* __restrict__ dO, const float* __restrict__ LSE,
                   uint4* __restrict__ dOh, float* __restrict__ delta, float* __restrict__ lse_log2,
                   float4* __restrict__ dQ, size_t rows, int D, int DP, int bf16) {
    constexpr int ROWS_PER_WARP = 32 / LANES_PER_ROW;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES_PER_ROW;          // which row of the warp's group
    const int l = lane % LANES_PER_ROW;            // position within the row
    const size_t warp_global = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    const size_t n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const int vec_per_row = DP >> 3;
    for (size_t base = warp_global * ROWS_PER_WARP; base < rows; base += n_warps * ROWS_PER_WARP) {
        const size_t row = base + sub;
        float acc = 0.f;
        if (row < rows) {
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
            if (row < rows && l == 0) {
                delta[row] = acc;
                lse_log2[row] = LSE[row] * 1.4426950408889634f;
            }
        }
    }
}

}  // namespace

cudaError_t launch_cast_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                            size_t rows, int D, int DP, int bf16, cudaStream_t st) {
    const size_t total = rows * (DP >> 3);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;     // grid-stride, a multiple of the SM count
    if (blocks == 0) blocks = 1;
    cast_qkv_kernel<<<dim3(static_cast<unsigned>(blocks), 3), 256, 0, st>>>(
        Q, K, V, static_cast<uint4*>(Qh), static_cast<uint4*>(Kh), static_cast<uint4*>(Vh), rows, D, DP, bf16);
    return cudaGetLastError();
}

cudaError_t launch_bwd_prepass(const float* O, const float* dO, const float* LSE, void* dOh, float* delta,
                               float* lse_log2, float* dQ_zero, size_t rows, int D, int DP, int bf16, int parts,
                               cudaStream_t st) {
    const int lanes = DP >> 3;                    // 8 (DP = 64) or 16 (DP = 128)
    const size_t rows_per_block = (256 / 32) * (32 / lanes);
    size_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks == 0) blocks = 1;
    const unsigned g = static_cast<unsigned>(blocks);
    uint4* dh = static_cast<uint4*>(dOh);
    float4* dq = reinterpret_cast<float4*>(dQ_zero);
#define FA2_PRE(L, P) bwd_prepass_kernel<L, P><<<g, 256, 0, st>>>(O, dO, LSE, dh, delta, lse_log2, dq, rows, D, DP, bf16)
    if (lanes == 8) { if (parts == 1) FA2_PRE(8, 1); else if (parts == 2) FA2_PRE(8, 2); else FA2_PRE(8, 3); }
    else            { if (parts == 1) FA2_PRE(16, 1); else if (parts == 2) FA2_PRE(16, 2); else FA2_PRE(16, 3); }
#undef FA2_PRE
    return cudaGetLastError();
}

}  // namespace fa2
