// fa2_common.h -- internal declarations shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <mutex>

namespace fa2 {

enum Precision { kOperandF16 = 0, kOperandBF16 = 1 };

// 16-bit operand copies of the fp32 API tensors live in a per-device workspace as
// [BH][S][DP] row-major, DP = padded head dim (64 or 128), zero padded when D < DP.
inline int padded_head_dim(int D) { return D <= 64 ? 64 : 128; }

struct FwdParams {
    CUtensorMap tm_q;   // 16-bit [BH][S][DP], box {64, 128, 1}, 128B swizzle
    CUtensorMap tm_k;
    CUtensorMap tm_v;
    CUtensorMap tm_o;   // fp32 [BH][S][D], box {32, 128, 1}, 128B swizzle (store target)
    float* O;           // fp32 [BH][S][D]
    float* LSE;         // fp32 [BH][S] natural log
    int BH, S, D;
    float scale_log2;   // log2(e) / sqrt(D)
    float scale;        // 1 / sqrt(D)
    int bf16;
    // fused forward+backward only (all null otherwise): the forward also prepares the backward's side inputs
    const float* dO;    // fp32 [BH][S][D]
    void* dOh;          // 16-bit [BH][S][DP] copy of dO        (register-donor warps)
    float* dQ_zero;     // fp32 [BH][S][D], zero-filled          (register-donor warps)
    float* delta;       // fp32 [BH][S]  rowsum(dO * O)          (epilogue)
    float* lse_log2;    // fp32 [BH][S]  LSE * log2(e)           (epilogue)
    unsigned long long* timeline;   // debug builds (-DFA2_TIMELINE) only
};

struct BwdParams {
    CUtensorMap tm_q;    // 16-bit [BH][S][DP]
    CUtensorMap tm_k;
    CUtensorMap tm_v;
    CUtensorMap tm_do;
    CUtensorMap tm_dq;   // fp32 [BH][S][D], box {32, 128, 1}, 128B swizzle (reduce-add target)
    CUtensorMap tm_dk;   // same geometry, plain store targets
    CUtensorMap tm_dv;
    const float* lse_log2;  // fp32 [BH][S]  LSE * log2(e)
    const float* delta;     // fp32 [BH][S]  rowsum(dO * O)
    float* dQ;              // fp32 [BH][S][D], zeroed by the pre-pass, reduce-added here
    float* dK;
    float* dV;
    int BH, S, D;
    float scale_log2;
    float scale;
    int bf16;
    unsigned long long* timeline;   // debug builds (-DFA2_TIMELINE) only: per-role clock64 stamps of CTA 0
};

// Opt a kernel in to > 48 KB of dynamic shared memory once per (kernel, device) instead of on every launch.
inline cudaError_t ensure_smem_optin(const void* kern, int smem_bytes) {
    static std::mutex mu;
    static const void* kerns[16];
    static bool done[16][64];
    static int n = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    int k = 0;
    while (k < n && kerns[k] != kern) ++k;
    if (k == n) {
        if (n == 16 || dev < 0 || dev >= 64) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        kerns[n++] = kern;
    }
    if (dev < 0 || dev >= 64 || !done[k][dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[k][dev] = true;
    }
    return cudaSuccess;
}

// launchers (each returns a cudaError_t from the launch)
cudaError_t launch_cast_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                            size_t rows, int D, int DP, int bf16, cudaStream_t st);
cudaError_t launch_bwd_prepass(const float* O, const float* dO, const float* LSE, void* dOh, float* delta,
                               float* lse_log2, float* dQ_zero, size_t rows, int D, int DP, int bf16,
                               int parts /* 1: dO cast + dQ zero, 2: D_i + LSE, 3: both */, cudaStream_t st);
cudaError_t launch_fwd(const FwdParams& p, cudaStream_t st);
cudaError_t launch_bwd(const BwdParams& p, cudaStream_t st);
cudaError_t warm_fwd();
cudaError_t warm_bwd();

}  // namespace fa2
