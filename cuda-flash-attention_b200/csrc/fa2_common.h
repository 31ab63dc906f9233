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

// Range block at offset 0 of every per-device workspace: running |x| maxima collected by the cast passes and the
// per-launch scale factors the fix-up kernels derive from them (the `fp32` API must keep fp32 RANGE although the
// tensor cores see fp16 operands: a tensor whose largest magnitude does not fit fp16 comfortably is re-cast with a
// power-of-two scale whose inverse is folded into the softmax scale / the epilogues; exact, since powers of two).
// The last block of a cast pass turns the maxima into scales and clears them again for the next call.
constexpr int kAmaxLanes = 16;              // atomics to one address serialise in L2: spread each tensor over 16 words
enum RangeIdx {
    kSq = 0, kSk, kSv, kSdo,                // scales applied to the 16-bit copies of Q, K, V, dO (s_do includes the factor
                                            // that keeps the 16-bit dP - D_i in range)
    kC2,                                    // log2(e) / sqrt(D) / (s_q s_k)     softmax exponent scale
    kScaleLse,                              // 1 / sqrt(D) / (s_q s_k)           raw score -> natural-log units
    kInvV,                                  // 1 / s_v                           folded into O's 1 / l
    kAmaxV,                                 // max |V| s_v (kept for the dO fix-up)
    kDeltaMul,                              // s_v s_do                          D_i -> units of dP' = V' dO'^T
    kDkMul, kDqMul, kDvMul,                 // epilogue factors of dK, dQ, dV (1 / sqrt(D) and the inverse scales)
    kRangeCount = 16
};
struct RangeBlock {
    unsigned amax[4][kAmaxLanes];           // [Q, K, V, dO][lane]: float bits of max |x| (order-preserving), zero between calls
    float sc[kRangeCount];
    unsigned ticket[2];                     // arrivals of the Q/K/V cast blocks / of the dO cast blocks (or donor warps)
    unsigned coop_arrive, coop_release;     // grid barrier of the small-problem cast kernel (cooperative launch)
};
constexpr size_t kRangeBytes = 1024;        // the block's share of the workspace (keeps the tensors 1024-B aligned)
static_assert(sizeof(RangeBlock) <= kRangeBytes, "RangeBlock must fit its slot");

struct FwdParams {
    CUtensorMap tm_q;   // 16-bit [BH][S][DP], box {64, 128, 1}, 128B swizzle
    CUtensorMap tm_k;
    CUtensorMap tm_v;
    CUtensorMap tm_o;   // fp32 [BH][S][D], box {32, 128, 1}, 128B swizzle (store target)
    float* O;           // fp32 [BH][S][D]
    float* LSE;         // fp32 [BH][S] natural log
    // Geometry.  One launch covers S_q query rows of every slab against S_kv key/value rows.  Normally
    // S_q = S_kv = q_pitch = S; the sequence-split path launches a row range [r0, r1) of every slab: the tensor maps
    // and the row-indexed pointers (LSE, dO, delta, lse_log2) are pre-offset by r0 and q_pitch stays the full S.
    int BH, S_q, S_kv, q_pitch, D;
    long long donor_rows;   // rows of dO / dQ the donor warps cast / zero-fill (all BH * S rows, whatever S_q is)
    float scale_log2;   // log2(e) / sqrt(D)
    float scale;        // 1 / sqrt(D)
    int bf16;
    const float* range; // RangeBlock::sc of this launch (null: the unscaled defaults above)
    RangeBlock* rb;     // fused call: the donor warps publish max|dO| and the last of them decides dO's scale
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
    CUtensorMap tm_q64;  // CTA-pair kernel: Q / dO with box {64, 64, 1} (the 64 query rows one CTA feeds to S^T / dP^T)
    CUtensorMap tm_do64;
    CUtensorMap tm_dq64; // CTA-pair kernel: dQ with box {32, 64, 1} (each CTA reduce-adds 64 of the 128 query rows)
    CUtensorMap tm_dq;   // fp32 [BH][S][D], box {32, 128, 1}, 128B swizzle (reduce-add target)
    CUtensorMap tm_dk;   // same geometry, plain store targets
    CUtensorMap tm_dv;
    const float* lse_log2;  // fp32 [BH][S]  LSE * log2(e)
    const float* delta;     // fp32 [BH][S]  rowsum(dO * O)
    float* dQ;              // fp32 [BH][S][D], zeroed by the pre-pass, reduce-added here
    float* dK;
    float* dV;
    // One launch covers S_kv key/value rows of every slab (work items = their 128-row tiles; the sequence-split path
    // passes a row range through pre-offset tensor maps) against all S_q query rows.
    int BH, S_q, S_kv, D;
    float scale_log2;
    float scale;
    int bf16;
    const float* range;             // RangeBlock::sc of this launch (null: the unscaled defaults above)
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

// SM count of the current device (cached per device).
inline int sm_count_current() {
    static int count[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (count[dev] == 0 && cudaDeviceGetAttribute(&count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    return count[dev];
}

// launchers (each returns a cudaError_t from the launch)
cudaError_t launch_cast_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                            size_t rows, int D, int DP, int bf16, RangeBlock* rb, float scale,
                            float scale_log2, cudaStream_t st);
cudaError_t launch_bwd_prepass(const float* O, const float* dO, const float* LSE, void* dOh, float* delta,
                               float* lse_log2, float* dQ_zero, size_t rows, int D, int DP, int bf16,
                               int parts /* 1: dO cast + dQ zero, 2: D_i + LSE, 3: both */, RangeBlock* rb,
                               float scale, cudaStream_t st, unsigned range_rows = 1, unsigned pitch_rows = 1);
// dQ reduce-scatter of the sequence-split path: own[i] += sum over peers of peer[i] on `cnt` row ranges of
// `seg_floats` floats, `pitch_floats` apart; the peer pointers are other GPUs' memory read over NVLink (P2P loads).
cudaError_t launch_dq_peer_reduce(float* own, const float* const* peers, int n_peers, size_t seg_floats,
                                  size_t pitch_floats, int cnt, cudaStream_t st);
// Small problems (cast_small_fits: at most one 8-element vector per thread and tensor on one 1024-thread block per
// SM): ONE cooperative launch reads Q, K, V (and dO) once, measures max|x|, decides the scales behind a grid barrier
// and casts from registers; with dO it also zero-fills dQ.  No re-cast kernels needed.
bool cast_small_fits(size_t rows, int DP, int n_sm);
cudaError_t launch_cast_small(const float* Q, const float* K, const float* V, const float* dO, void* Qh, void* Kh,
                              void* Vh, void* dOh, float* dQ_zero, size_t rows, int D, int DP, int bf16, RangeBlock* rb,
                              float scale, float scale_log2, int n_sm, cudaStream_t st);
// Re-cast kernels (always launched; they return after one load unless a scale other than 1 was decided).
cudaError_t launch_range_fix_qkv(const float* Q, const float* K, const float* V, void* Qh, void* Kh, void* Vh,
                                 size_t rows, int D, int DP, int bf16, const RangeBlock* rb, cudaStream_t st);
cudaError_t launch_range_fix_do(const float* dO, void* dOh, size_t rows, int D, int DP, int bf16, const RangeBlock* rb,
                                cudaStream_t st);
cudaError_t launch_fwd(const FwdParams& p, cudaStream_t st);
cudaError_t launch_bwd(const BwdParams& p, cudaStream_t st);      // picks the CTA-pair kernel when bwd_uses_pair(D)
cudaError_t launch_bwd2(const BwdParams& p, cudaStream_t st);     // fa2_bwd2_sm100.cu: cluster of two CTAs per pair of KV tiles (D = 128)
bool bwd_uses_pair(int D);
cudaError_t warm_bwd2();
cudaError_t warm_fwd();
cudaError_t warm_bwd();

}  // namespace fa2
