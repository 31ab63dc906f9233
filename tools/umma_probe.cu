// umma_probe.cu -- GPU-side unit probe for the descriptor / layout conventions in ptx.cuh.
// Runs four 128x128x128 fp16 GEMMs through TMA + tcgen05.mma and checks them on the host:
//   mode 0: D = A * B^T        A [m][k] K-major smem,  B [n][k] K-major smem      (S = Q K^T)
//   mode 1: D = A * B          A K-major smem,         B [k][n] MN-major smem     (dQ = dS K)
//   mode 2: D = A * B^T        A from TMEM (tcgen05.st of packed fp16), B K-major
//   mode 3: D = A * B          A from TMEM,            B [k][n] MN-major smem     (O = P V)
//   mode 4: D = A^T * B        A [k][m] MN-major smem, B [k][n] MN-major smem     (dK = dS^T Q)
// Diagnostic tool only (not part of the library).  Build: see tools/Makefile target in
// cuda-flash-attention_b200/Makefile (`make probe`).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_fp16.h>
#include "../cuda-flash-attention_b200/csrc/ptx.cuh"

using namespace fa2;

struct ProbeParams {
    CUtensorMap tm_a;   // fp16 [1][128][128], box {64,128,1}, SW128
    CUtensorMap tm_b;
    const __half* a_gmem;
    float* out;         // [128][128]
    int mode;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;               // 32 KB: two [128][64] atoms
    uint8_t* sB = smem + 32768;       // 32 KB
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 65536);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* holder = reinterpret_cast<uint32_t*>(bar_load + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) { mbar_init(bar_load, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc(holder, 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    const bool a_tmem = (p.mode == 2 || p.mode == 3);
    const bool b_mn = (p.mode == 1 || p.mode == 3 || p.mode == 4);
    const bool a_mn = (p.mode == 4);

    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_load, a_tmem ? 32768 : 65536);
        if (!a_tmem) {
            tma_load_3d(sA, &p.tm_a, bar_load, 0, 0, 0);
            tma_load_3d(sA + 16384, &p.tm_a, bar_load, 64, 0, 0);
        }
        tma_load_3d(sB, &p.tm_b, bar_load, 0, 0, 0);
        tma_load_3d(sB + 16384, &p.tm_b, bar_load, 64, 0, 0);
    }
    if (a_tmem) {
        // thread == row; pack 128 fp16 of the row into 64 columns at TMEM column 128
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(p.a_gmem + (size_t)threadIdx.x * 128);
        uint32_t r[32];
        for (int c = 0; c < 2; ++c) {
            for (int i = 0; i < 32; ++i) r[i] = row[c * 32 + i];
            tmem_st32(taddr + c * 32, r);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
        mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, 128, a_mn ? 1 : 0, b_mn ? 1 : 0, 0);
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
        for (int k = 0; k < 8; ++k) {
            const uint32_t koff = (k >> 2) * 16384 + (k & 3) * 32;     // K-major: atom, then 32 B per K-step
            const uint32_t mnoff = k * 2048;                             // MN-major: 16 k-rows of 128 B
            const uint64_t bd = b_mn ? umma_smem_desc(b_addr + mnoff, 16384, 1024)
                                     : umma_smem_desc(b_addr + koff, 16, 1024);
            if (a_tmem) {
                umma_ts(tmem, tmem + 128 + k * 8, bd, idesc, k > 0);
            } else {
                const uint64_t ad = a_mn ? umma_smem_desc(a_addr + mnoff, 16384, 1024)
                                         : umma_smem_desc(a_addr + koff, 16, 1024);
                umma_ss(tmem, ad, bd, idesc, k > 0);
            }
        }
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    {
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) p.out[(size_t)threadIdx.x * 128 + c * 32 + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    const int N = 128;
    std::vector<__half> hA(N * N), hB(N * N);
    std::vector<float> fA(N * N), fB(N * N);
    srand(1);
    for (int i = 0; i < N * N; ++i) {
        fA[i] = (float)(rand() % 17 - 8) / 8.0f; fB[i] = (float)(rand() % 13 - 6) / 4.0f;
        hA[i] = __float2half(fA[i]); hB[i] = __float2half(fB[i]);
    }
    __half *dA, *dB; float* dOut;
    CK(cudaMalloc(&dA, N * N * 2)); CK(cudaMalloc(&dB, N * N * 2)); CK(cudaMalloc(&dOut, N * N * 4));
    CK(cudaMemcpy(dA, hA.data(), N * N * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), N * N * 2, cudaMemcpyHostToDevice));
    ProbeParams p{};
    cuuint64_t dims[3] = {128, 128, 1}; cuuint64_t strides[2] = {256, 256 * 128};
    cuuint32_t box[3] = {64, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
    for (CUtensorMap* tm : {&p.tm_a, &p.tm_b}) {
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, tm == &p.tm_a ? (void*)dA : (void*)dB, dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    }
    p.a_gmem = dA; p.out = dOut;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));
    int bad = 0;
    for (int mode = 0; mode < 5; ++mode) {
        p.mode = mode;
        CK(cudaMemset(dOut, 0, N * N * 4));
        probe_kernel<<<1, 128, 70000>>>(p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> out(N * N);
        CK(cudaMemcpy(out.data(), dOut, N * N * 4, cudaMemcpyDeviceToHost));
        const bool b_mn = (mode == 1 || mode == 3 || mode == 4), a_mn = (mode == 4);
        double maxerr = 0;
        for (int m = 0; m < N; ++m)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < N; ++k) {
                    const float a = a_mn ? fA[k * N + m] : fA[m * N + k];
                    const float b = b_mn ? fB[k * N + n] : fB[n * N + k];
                    acc += (double)a * b;
                }
                maxerr = fmax(maxerr, fabs(acc - out[m * N + n]));
            }
        printf("umma_probe mode %d: max_abs_err = %.6f  %s\n", mode, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
        if (!(maxerr < 1e-3)) bad++;
    }
    return bad ? 1 : 0;
}
