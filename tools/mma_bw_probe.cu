// mma_bw_probe.cu -- microbenchmark: do tcgen05.mma operand reads and LSU traffic share the SM's shared-memory
// bandwidth?  One CTA per SM.  Warp 0 issues back-to-back 128 x N x 128 fp16 GEMMs (8 K-steps of 16); warps 1-4
// optionally hammer shared memory with conflict-free 16-byte loads or stores until the MMA warp is done.
//   mma kind:  0 none | 1 SS N=128 | 2 TS N=128 (A from TMEM) | 3 SS N=256 | 4 TS N=256 | 5-7 N=128 with MN-major operands
//              (5: B MN-major = dK shape, 6: TS with B MN-major = PV / dV shape, 7: A and B MN-major = dQ shape)
//   traffic:   0 none | 1 LDS.128 | 2 STS.128
// Prints cycles per GEMM (floor: 512 at N=128, 1024 at N=256) and the LSU bytes/clk/SM achieved next to it.
// Diagnostic tool only (feeds the shared-memory traffic model in DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>
#include "../cuda-flash-attention_b200/csrc/ptx.cuh"
using namespace fa2;

struct Params { int mma; int traffic; int iters; long long* cycles; unsigned long long* bytes; };

template <int N, bool TS, bool A_MN = false, bool B_MN = false>
__device__ __forceinline__ void gemm_once(uint32_t tmem_d, uint32_t tmem_a, uint32_t a_lo, uint32_t b_lo) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, A_MN ? 1 : 0, B_MN ? 1 : 0, 0);
    constexpr uint32_t hi = umma_desc_hi(1024);
    static_for<8>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        constexpr uint32_t a_off = A_MN ? koff_mnmajor(k) : koff_kmajor(k, 16384);
        constexpr uint32_t b_off = B_MN ? koff_mnmajor(k) : koff_kmajor(k, N * 128);
        if constexpr (TS) umma_ts_off<k * 8, b_off>(tmem_d, tmem_a, b_lo, hi, idesc, k > 0);
        else              umma_ss_off<a_off, b_off>(tmem_d, a_lo, b_lo, hi, idesc, k > 0);
    });
}

__global__ void __launch_bounds__(160, 1) k(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                   // 32 KB
    uint8_t* sB = smem + 32768;           // 64 KB (N up to 256)
    uint8_t* sX = smem + 98304;           // 64 KB scratch for the LSU traffic
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 163840 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
    if (warp == 0) { __syncwarp(); tmem_alloc(&holder, 512); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (warp == 0) {
        // MN-major operands: LBO = distance to the next 64-wide chunk (16 KB), K-major: unused
        const bool mn = p.mma >= 5;
        const uint32_t a_lo = umma_desc_lo(smem_u32(sA), p.mma == 7 ? 16384 : 16), b_lo = umma_desc_lo(smem_u32(sB), mn ? 16384 : 16);
        const long long t0 = clock64();
        if (p.mma) {
            for (int it = 0; it < p.iters; ++it) {
                if (elect_one()) {
                    const uint32_t d = tmem + (it & 1) * 0;     // same accumulator: the pipe is in order anyway
                    switch (p.mma) {
                        case 1: gemm_once<128, false>(d, tmem + 256, a_lo, b_lo); break;
                        case 2: gemm_once<128, true>(d, tmem + 256, a_lo, b_lo); break;
                        case 3: gemm_once<256, false>(d, tmem + 256, a_lo, b_lo); break;
                        case 4: gemm_once<256, true>(d, tmem + 256, a_lo, b_lo); break;
                        case 5: gemm_once<128, false, false, true>(d, tmem + 256, a_lo, b_lo); break;   // dK shape
                        case 6: gemm_once<128, true, false, true>(d, tmem + 256, a_lo, b_lo); break;    // PV / dV shape
                        default: gemm_once<128, false, true, true>(d, tmem + 256, a_lo, b_lo); break;   // dQ shape
                    }
                    if ((it & 7) == 7 || it == p.iters - 1) umma_commit(&bar);
                }
                __syncwarp();
                if ((it & 7) == 7 || it == p.iters - 1) { mbar_wait(&bar, (it >> 3) & 1); tc_fence_after(); }
            }
        } else {
            while (clock64() - t0 < 400000) {}
        }
        const long long t1 = clock64();
        done = 1;
        if (lane == 0) p.cycles[blockIdx.x] = t1 - t0;
    } else if (p.traffic) {
        const uint32_t base = smem_u32(sX) + (threadIdx.x - 32) * 16;       // 128 threads x 16 B = 2 KB per sweep row
        unsigned long long n = 0;
        uint32_t acc = 0;
        while (!done) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (p.traffic == 1) {
                    uint32_t a, b, c, d;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + j * 2048));
                    acc ^= a ^ b ^ c ^ d;
                } else {
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + j * 2048), "r"(acc) : "memory");
                }
            }
            n += 32 * 16;
        }
        if (acc == 0x12345u) p.bytes[0] = 1;
        atomicAdd(p.bytes + 1 + blockIdx.x, n);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main() {
    const int nsm = 148;
    long long* cyc; unsigned long long* bytes;
    CK(cudaMalloc(&cyc, nsm * 8)); CK(cudaMalloc(&bytes, (nsm + 1) * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840));
    const char* mma_name[] = {"none", "SS N=128", "TS N=128", "SS N=256", "TS N=256", "SS A-K B-MN", "TS B-MN", "SS A-MN B-MN"};
    const char* tr_name[] = {"none", "LDS.128", "STS.128"};
    for (int mma = 0; mma < 8; ++mma) {
        for (int tr = 0; tr < 3; ++tr) {
            if (mma == 0 && tr == 0) continue;
            Params p{mma, tr, 512, cyc, bytes};
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(bytes, 0, (nsm + 1) * 8));
                k<<<nsm, 160, 163840>>>(p);
                CK(cudaDeviceSynchronize());
            }
            long long h[148]; unsigned long long hb[149];
            CK(cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hb, bytes, (nsm + 1) * 8, cudaMemcpyDeviceToHost));
            double c = 0, b = 0;
            for (int i = 0; i < nsm; ++i) { c += h[i]; b += hb[1 + i]; }
            c /= nsm; b /= nsm;
            const double operand = (mma == 0) ? 0 : ((mma == 1 || mma == 5 || mma == 7) ? 65536 : (mma == 2 || mma == 6) ? 32768 : (mma == 3) ? 98304 : 65536);
            printf("mma %-13s traffic %-8s: %7.1f cycles/GEMM   LSU %6.1f B/clk/SM   MMA operand reads %6.1f B/clk/SM\n",
                   mma_name[mma], tr_name[tr], mma ? c / p.iters : 0.0, b / c, mma ? operand * p.iters / c : 0.0);
        }
    }
    return 0;
}
