// mma2_probe.cu -- validates and times tcgen05.mma.cta_group::2 (a CTA pair computing one M=256 tile: each SM holds
// its own 128 rows of A and HALF of B) against the cta_group::1 numbers of mma_bw_probe.cu.
//   D[256 x 128] = A[256 x 128] * B[128 x 128]^T, fp16 in, fp32 out; CTA r owns A rows [128r, 128r+128) and
//   B rows (= N columns of D) [64r, 64r+64).   A[m][k] = rank+1, B[n][k] = n%4+1+4(n/64)  =>  D[m][n] = 128 (rank+1) B[n].
//   mode 0: SS (A from shared memory)     mode 1: TS (A from TMEM)
// One cluster of 2 per SM pair; the leader CTA issues every MMA; tcgen05.commit multicasts to both CTAs' barriers.
// Diagnostic tool only.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../cuda-flash-attention_b200/csrc/ptx.cuh"
using namespace fa2;

struct Params { int mode; int iters; long long* cycles; int* errors; };

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit all prior MMAs of this thread; arrive on the barrier at the same smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
template <uint32_t A_OFF, uint32_t B_OFF>
__device__ __forceinline__ void umma2_ss_off(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 al, bl;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\tadd.u32 al, %1, %6;\n\tadd.u32 bl, %2, %7;\n\t"
        "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc), "n"(A_OFF), "n"(B_OFF) : "memory");
}
template <uint32_t A_COL_OFF, uint32_t B_OFF>
__device__ __forceinline__ void umma2_ts_off(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 at, bl;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\tadd.u32 at, %1, %6;\n\tadd.u32 bl, %2, %7;\n\t"
        "mov.b64 db, {bl, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [at], db, %4, p;\n\t}\n"
        ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc), "n"(A_COL_OFF), "n"(B_OFF) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                   // 2 atoms [128 rows][128 B] = 32 KB
    uint8_t* sB = smem + 32768;           // 2 atoms [64 rows][128 B]  = 16 KB (this CTA's half of B)
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    {
        const __half av = __float2half(float(rank + 1));
        __half* a = reinterpret_cast<__half*>(sA);
        for (int i = threadIdx.x; i < 16384; i += blockDim.x) a[i] = av;
        __half* b = reinterpret_cast<__half*>(sB);
        for (int i = threadIdx.x; i < 8192; i += blockDim.x) {
            const int atom = i / 4096, row = (i % 4096) / 64;      // [atom][row][64 halves]
            (void)atom;
            b[i] = __float2half(float((row % 4) + 1 + 4 * rank));
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc2(&holder, 512); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (p.mode == 1) {       // A into TMEM columns [256, 320): 128 fp16 per lane
        uint32_t r[32];
        const __half2 h2 = __float2half2_rn(float(rank + 1));
        for (int i = 0; i < 32; ++i) r[i] = *reinterpret_cast<const uint32_t*>(&h2);
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256;
        tmem_st32(taddr, r);
        tmem_st32(taddr + 32, r);
        tmem_wait_st();
        tc_fence_before();
    }
    cluster_sync();          // both CTAs: smem filled, TMEM allocated, barriers initialised
    tc_fence_after();
    long long t0 = clock64();
    if (rank == 0 && warp == 0) {
        const uint32_t a_lo = umma_desc_lo(smem_u32(sA), 16), b_lo = umma_desc_lo(smem_u32(sB), 16);
        constexpr uint32_t idesc = umma_idesc_f16(256, 128, 0, 0, 0);
        constexpr uint32_t hi = umma_desc_hi(1024);
        for (int it = 0; it < p.iters; ++it) {
            if (elect_one()) {
                if (p.mode == 0) {
                    static_for<8>([&](auto kc) {
                        constexpr int kk = decltype(kc)::value;
                        umma2_ss_off<koff_kmajor(kk, 16384), koff_kmajor(kk, 8192)>(tmem, a_lo, b_lo, hi, idesc, kk > 0);
                    });
                } else {
                    static_for<8>([&](auto kc) {
                        constexpr int kk = decltype(kc)::value;
                        umma2_ts_off<kk * 8, koff_kmajor(kk, 8192)>(tmem, tmem + 256, b_lo, hi, idesc, kk > 0);
                    });
                }
                if ((it & 7) == 7 || it == p.iters - 1) umma_commit2(&bar, 3);
            }
            __syncwarp();
            if ((it & 7) == 7 || it == p.iters - 1) { mbar_wait(&bar, (it >> 3) & 1); tc_fence_after(); }
        }
    } else if (rank == 1 && warp == 0) {
        const int n_commit = (p.iters + 7) / 8;
        for (int c = 0; c < n_commit; ++c) mbar_wait(&bar, c & 1);
        tc_fence_after();
    }
    long long t1 = clock64();
    __syncthreads();
    tc_fence_after();
    {   // verify this CTA's 128 x 128 accumulator
        int bad = 0;
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) {
                const int n = c * 32 + i;
                const float want = 128.0f * float(rank + 1) * float((n % 4) + 1 + 4 * (n / 64));
                if (__uint_as_float(r[i]) != want) ++bad;
            }
        }
        if (bad) atomicAdd(p.errors, bad);
    }
    if (threadIdx.x == 0) p.cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    cluster_sync();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main() {
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc; int* err;
    CK(cudaMalloc(&cyc, 256 * 8)); CK(cudaMalloc(&err, 4));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nsm); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 49152 + 1024;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, k, &cfg));
        printf("SMs %d, max co-resident clusters of 2 (48 KB smem): %d\n", nsm, ncl);
    }
    for (int grid : {2, nsm & ~1}) {
        for (int mode = 0; mode < 2; ++mode) {
            Params p{mode, 512, cyc, err};
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(err, 0, 4));
                k<<<grid, 128, 49152 + 1024>>>(p);
                CK(cudaDeviceSynchronize());
            }
            long long h[256]; int e;
            CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost));
            double c = 0; for (int i = 0; i < grid; i += 2) c += h[i]; c /= (grid / 2);
            printf("grid %3d cta_group::2 %s M=256 N=128 K=128: %7.1f cycles per GEMM (per-SM share 128x128x128; cta_group::1: SS 715, TS 594), wrong elements %d\n",
                   grid, mode ? "TS" : "SS", c / p.iters, e);
        }
    }
    return 0;
}
