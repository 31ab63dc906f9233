// fa2_range.cuh -- device side of the fp32-range handling (see RangeBlock in fa2_common.h).
//
// The reference's fp32 kernels take any fp32 input (kernel_fa2_optimized.cu:19-347, f-attn2-backward.cu:243-266);
// fp16 tensor-core operands do not: they overflow at 65504 and lose precision below 6e-5.  The cast passes therefore
// collect max|x| per tensor, the LAST block (or warp) of a cast pass to finish turns the maxima into power-of-two
// scales ("decide"), and a re-cast kernel redoes the 16-bit copy of whatever tensor got a scale other than 1 (rare:
// for randn / rand / ones data every scale is 1 and the re-cast kernels return after one load).  The inverse scales
// are folded into the softmax scale and the epilogue factors of the main kernels; powers of two, so the result is
// what exact scaling gives.  Scaling is per tensor and launch: a single tensor whose (b,h) slabs differ in magnitude
// by more than ~2^20 still loses its small slabs, as any 16-bit copy with one scale must.
#pragma once
#include "fa2_common.h"

namespace fa2 {

__device__ __forceinline__ float amax_of(const unsigned* lanes) {
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < kAmaxLanes; ++i) { const unsigned v = __ldcg(lanes + i); b = v > b ? v : b; }
    return __uint_as_float(b);
}

// Scale for the 16-bit copy of a tensor whose largest magnitude is amax: 1 while amax sits in [2^-6, hi) (every
// usual input), otherwise the power of two that brings it to [2, 4).  bf16 copies keep the fp32 exponent range.
__device__ __forceinline__ float pick_scale(float amax, int bf16, float hi = 32768.0f) {
    if (bf16 || !(amax > 0.0f) || amax > 3.0e38f) return 1.0f;
    if (amax >= 0.015625f && amax < hi) return 1.0f;
    int e;
    frexpf(amax, &e);                       // amax = m 2^e, m in [0.5, 1)  =>  floor(log2 amax) = e - 1
    int k = 1 - (e - 1);
    k = k > 40 ? 40 : (k < -40 ? -40 : k);
    return ldexpf(1.0f, k);
}

// One thread, after every block of the Q/K/V cast pass has published its maxima.  scale = 1/sqrt(D) and
// scale_log2 = scale * log2(e) as the host computed them: with all scales 1 the factors are bit-identical to them.
__device__ __forceinline__ void decide_qkv(RangeBlock* rb, int bf16, float scale, float scale_log2) {
    const float am_q = amax_of(rb->amax[0]), am_k = amax_of(rb->amax[1]), am_v = amax_of(rb->amax[2]);
    // (V's window ends at 2^7: max|V'| bounds the factor the dO decision may have to take out of dO')
    const float sq = pick_scale(am_q, bf16), sk = pick_scale(am_k, bf16), sv = pick_scale(am_v, bf16, 128.0f);
    rb->sc[kSq] = sq; rb->sc[kSk] = sk; rb->sc[kSv] = sv;
    rb->sc[kC2] = scale_log2 / sq / sk;
    rb->sc[kScaleLse] = scale / sq / sk;
    rb->sc[kInvV] = 1.0f / sv;
    rb->sc[kAmaxV] = am_v * sv;
    // every block has published and nobody else reads the maxima: clear them for the next call on this device
    for (int i = 0; i < 3 * kAmaxLanes; ++i) (&rb->amax[0][0])[i] = 0u;
}

// One thread, after the dO cast pass (pre-pass blocks or the forward's donor warps) has published its maximum.
// |dP' - D_i'| <= 2 D max|dO'| max|V'|, and that difference is rounded to 16 bit before it multiplies P.  While the
// bound sits in [2^9, 2^15] (randn / rand / ones data) nothing is done; otherwise dO' takes the power of two that
// brings it to (2^11, 2^12]: no overflow, and dS = P (dP - D_i) stays clear of the fp16 subnormals.  With max|V'|
// in [2^-6, 2^7) that leaves max|dO'| between 2^-3 and 2^12.
__device__ __forceinline__ void decide_do(RangeBlock* rb, int D, int bf16, float scale) {
    const float am_do = amax_of(rb->amax[3]);
    float sdo = pick_scale(am_do, bf16);
    const float bound = am_do * sdo * rb->sc[kAmaxV] * 2.0f * static_cast<float>(D);
    if (!bf16 && bound > 0.0f && bound < 3.0e38f && !(bound >= 512.0f && bound <= 32768.0f)) {
        int e;
        const float m = frexpf(bound, &e);                   // bound = m 2^e, m in [0.5, 1): ceil(log2) = e (e - 1 at m = 0.5)
        int k = 12 - (m == 0.5f ? e - 1 : e);
        k = k > 60 ? 60 : (k < -60 ? -60 : k);
        sdo *= ldexpf(1.0f, k);
    }
    const float sq = rb->sc[kSq], sk = rb->sc[kSk], sv = rb->sc[kSv];
    rb->sc[kSdo] = sdo;
    rb->sc[kDeltaMul] = sv * sdo;
    rb->sc[kDkMul] = scale / sv / sdo / sq;                  // (stepwise: every division by a power of two is exact)
    rb->sc[kDqMul] = scale / sv / sdo / sk;
    rb->sc[kDvMul] = 1.0f / sdo;
    for (int i = 0; i < kAmaxLanes; ++i) rb->amax[3][i] = 0u;       // cleared for the next call
}

// "Last one out decides": called by one thread per block / warp after its atomicMax.  Returns true for the caller
// that finished last (its reads see every other caller's maximum: fence + ticket), after re-arming the ticket.
__device__ __forceinline__ bool range_last_arrival(unsigned* ticket, unsigned n_arrivals) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    if (t != n_arrivals - 1u) return false;
    __threadfence();
    *ticket = 0u;
    return true;
}

}  // namespace fa2
