// TEST INFRASTRUCTURE (oracle).  Same idea as fwd_tu.cu for the reference's D_computation_reduction_kernel and
// flash_attention2_backward_kernel (/root/reference/kernels/f-attn2-backward.cu), launched as in :439-466.
#include "f-attn2-backward.cu"

template <int D>
static int run_bwd(const float* q, const float* k, const float* v, const float* o, const float* go, const float* lse,
                   float* dvec, float* dq, float* dk, float* dv, int B, int H, int S) {
    D_computation_reduction_kernel<D><<<B * H * S, D, sizeof(float) * D>>>(go, o, B, H, S, dvec);
    if (cudaGetLastError() != cudaSuccess) return 1;
    auto kern = flash_attention2_backward_kernel<32, 32, D>;
    const int smem = sizeof(shm_t<32, 32, D>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    const int T_c = (S + 31) / 32;
    kern<<<B * H * T_c, 256, smem>>>(q, k, v, o, go, lse, dvec, dq, dk, dv, B, H, S);
    return cudaGetLastError() != cudaSuccess;
}

extern "C" int ref_any_d_backward(const float* q, const float* k, const float* v, const float* o, const float* go,
                                  const float* lse, float* dvec, float* dq, float* dk, float* dv, int B, int H, int S,
                                  int D) {
    switch (D) {
        case 32: return run_bwd<32>(q, k, v, o, go, lse, dvec, dq, dk, dv, B, H, S);
        case 64: return run_bwd<64>(q, k, v, o, go, lse, dvec, dq, dk, dv, B, H, S);
        case 128: return run_bwd<128>(q, k, v, o, go, lse, dvec, dq, dk, dv, B, H, S);
        default: return 2;
    }
}

// The reference's own backward host launcher (f-attn2-backward.cu:384-485) at any head dim; see fwd_tu.cu.
template <int D>
static int run_host_bwd(const float* q, const float* k, const float* v, const float* o, const float* go,
                        const float* lse, float* dq, float* dk, float* dv, int B, int S, int H, TimerManager* tm) {
    auto kern = flash_attention2_backward_kernel<32, 32, D>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(shm_t<32, 32, D>)) != cudaSuccess) return 1;
    host_flash_attention2_backward<D>(q, k, v, o, go, lse, dq, dk, dv, B, S, H, tm);
    return 0;
}
int ref_any_d_host_backward(const float* q, const float* k, const float* v, const float* o, const float* go,
                            const float* lse, float* dq, float* dk, float* dv, int B, int H, int S, int D,
                            TimerManager* tm) {
    switch (D) {
        case 32: return run_host_bwd<32>(q, k, v, o, go, lse, dq, dk, dv, B, S, H, tm);
        case 64: return run_host_bwd<64>(q, k, v, o, go, lse, dq, dk, dv, B, S, H, tm);
        case 128: return run_host_bwd<128>(q, k, v, o, go, lse, dq, dk, dv, B, S, H, tm);
        default: return 2;
    }
}
