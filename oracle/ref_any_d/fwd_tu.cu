// TEST INFRASTRUCTURE (oracle).  Compiles the reference's forward kernel template as it lies in
// /root/reference/kernels/kernel_fa2_optimized.cu and instantiates it for head dims the reference's own
// dispatcher refuses (include/dispatcher.h:226-227: only 32 and 64), opting in to > 48 KB dynamic shared memory.
// No reference source is copied: the file is #included from the reference tree at build time (-I).
#include "kernel_fa2_optimized.cu"

template <int D>
static int run_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int H, int S) {
    auto kern = flash_attention2_forward_kernel<32, 32, D, 4, 4, 4>;
    const int smem = sizeof(shm_t<32, 32, D>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    const int T_r = (S + 31) / 32;
    kern<<<B * H * T_r, 256, smem>>>(q, k, v, o, lse, B, H, S);      // launch shape of kernel_fa2_optimized.cu:405-409
    return cudaGetLastError() != cudaSuccess;
}

extern "C" int ref_any_d_forward(const float* q, const float* k, const float* v, float* o, float* lse, int B, int H,
                                 int S, int D) {
    switch (D) {
        case 32: return run_fwd<32>(q, k, v, o, lse, B, H, S);
        case 64: return run_fwd<64>(q, k, v, o, lse, B, H, S);
        case 128: return run_fwd<128>(q, k, v, o, lse, B, H, S);
        default: return 2;
    }
}

// The reference's own host launcher (kernel_fa2_optimized.cu:351-423: cudaMalloc, H2D, timed launch, D2H, cudaFree)
// at any head dim.  It launches with sizeof(shm_t) and no opt-in, which only works up to 48 KB; the attribute is
// per function and sticky, so it is set here first.  Argument order (B, S, H) is the reference's.
template <int D>
static int run_host_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int S, int H,
                        TimerManager* tm) {
    auto kern = flash_attention2_forward_kernel<32, 32, D, 4, 4, 4>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(shm_t<32, 32, D>)) != cudaSuccess) return 1;
    host_flash_attention2_forward<D>(q, k, v, o, lse, B, S, H, tm);
    return 0;
}
int ref_any_d_host_forward(const float* q, const float* k, const float* v, float* o, float* lse, int B, int H, int S,
                           int D, TimerManager* tm) {
    switch (D) {
        case 32: return run_host_fwd<32>(q, k, v, o, lse, B, S, H, tm);
        case 64: return run_host_fwd<64>(q, k, v, o, lse, B, S, H, tm);
        case 128: return run_host_fwd<128>(q, k, v, o, lse, B, S, H, tm);
        default: return 2;
    }
}
