// reduce_probe.cu -- microbenchmark: how fast can one SM add a 16 KB fp32 tile into global memory?
//   mode 0: cp.reduce.async.bulk.tensor.3d (TMA tensor reduce, 128B-swizzled [128][32] box)   <- what fa2_bwd uses
//   mode 1: cp.reduce.async.bulk (1-D bulk reduce of 16 KB contiguous)
//   mode 2: red.global.add.v4.f32 from registers (128 threads x 32 x 16 B)
//   mode 3: red.global.add.f32 scalar, coalesced (warp covers 128 B)
//   mode 6: mode 0 while warp 1 keeps TMA-LOADING 16 KB per reduced tile (like the Q / dO stream of fa2_bwd)
//   mode 7: mode 0 while warps 1-3 load the same 16 KB per tile with cp.async (LDGSTS) instead of TMA
// One CTA per SM, ITERS tiles each, distinct destination tiles (L2 resident).  Diagnostic tool only.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../cuda-flash-attention_b200/csrc/ptx.cuh"
using namespace fa2;

struct Params { CUtensorMap tm; CUtensorMap tm512; float* dst; int iters; int mode; long long* cycles; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 2 * 4096; i += 128) s[i] = 1.0f;
    fence_proxy_async_smem();
    __syncthreads();
    const int tiles_per_cta = 32;                    // destination tiles this CTA cycles through
    float* base = p.dst + (size_t)blockIdx.x * tiles_per_cta * 4096;
    long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
        const int t = it % tiles_per_cta;
        if (p.mode == 0) {
            if (threadIdx.x == 0) {
                tma_reduce_add_3d(&p.tm, smem + (it & 1) * 16384, 0, t * 128, blockIdx.x);
                tma_store_commit();
                tma_store_wait_read<1>();
            }
        } else if (p.mode == 5) {
            // real dQ geometry: box {32 cols, 128 rows} inside rows of 128 floats (512 B pitch)
            if (threadIdx.x == 0) {
                tma_reduce_add_3d(&p.tm512, smem + (it & 1) * 16384, (it & 3) * 32, (t / 4) * 128, blockIdx.x);
                tma_store_commit();
                tma_store_wait_read<1>();
            }
        } else if (p.mode == 6 || p.mode == 7) {
            if (threadIdx.x == 0) {
                tma_reduce_add_3d(&p.tm, smem + (it & 1) * 16384, 0, t * 128, blockIdx.x);
                tma_store_commit();
                tma_store_wait_read<1>();
            } else if (p.mode == 6 && threadIdx.x == 32) {
                uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
                if (it == 0) { mbar_init(bar, 1); fence_mbar_init(); }
                mbar_expect_tx(bar, 16384);
                tma_load_3d(smem + 32768, &p.tm, bar, 0, ((t + 7) % tiles_per_cta) * 128, blockIdx.x);
                mbar_wait(bar, it & 1);
            } else if (p.mode == 7 && threadIdx.x >= 32) {
                const float* src = base + (size_t)((t + 7) % tiles_per_cta) * 4096;
                for (int q = threadIdx.x - 32; q < 1024; q += 96) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem + 32768 + q * 16)), "l"(src + q * 4) : "memory");
                }
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            }
        } else if (p.mode == 1) {
            if (threadIdx.x == 0) {
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(base + (size_t)t * 4096), "r"(smem_u32(smem + (it & 1) * 16384)), "r"(16384) : "memory");
                tma_store_commit();
                tma_store_wait_read<1>();
            }
        } else if (p.mode == 2) {
            float* row = base + (size_t)t * 4096 + threadIdx.x * 32;      // thread = row of 32 floats
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + q * 4), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
        } else if (p.mode == 3) {
            float* tile = base + (size_t)t * 4096;
#pragma unroll
            for (int q = 0; q < 32; ++q) atomicAdd(tile + q * 128 + threadIdx.x, 1.0f);
        } else {
            // mode 4: warp 0 drives the TMA reduce of one 16 KB tile while warps 1-3 add a SECOND 16 KB tile with
            // coalesced scalar atomics (are the two paths additive?)
            if (threadIdx.x == 0) {
                tma_reduce_add_3d(&p.tm, smem + (it & 1) * 16384, 0, t * 128, blockIdx.x);
                tma_store_commit();
                tma_store_wait_read<1>();
            } else if (threadIdx.x >= 32) {
                float* tile = base + (size_t)((t + 16) % tiles_per_cta) * 4096;
                for (int q = threadIdx.x - 32; q < 4096; q += 96) atomicAdd(tile + q, 1.0f);
            }
        }
    }
    if (threadIdx.x == 0 && (p.mode < 2 || p.mode >= 4)) tma_store_wait<0>();
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) p.cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    const int nsm = 148, tiles = 32;
    float* dst; long long* cyc;
    CK(cudaMalloc(&dst, (size_t)nsm * tiles * 4096 * 4)); CK(cudaMemset(dst, 0, (size_t)nsm * tiles * 4096 * 4));
    CK(cudaMalloc(&cyc, nsm * 8));
    Params p{}; p.dst = dst; p.cycles = cyc; p.iters = 256;
    // tensor view: [nsm][tiles*128 rows][32 cols] fp32, box {32,128,1}
    cuuint64_t dims[3] = {32, (cuuint64_t)tiles * 128, (cuuint64_t)nsm}; cuuint64_t strides[2] = {128, (cuuint64_t)tiles * 128 * 128};
    cuuint32_t box[3] = {32, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    {
        cuuint64_t d2[3] = {128, (cuuint64_t)tiles * 32, (cuuint64_t)nsm}; cuuint64_t s2[2] = {512, (cuuint64_t)tiles * 32 * 512};
        CUresult r2 = enc(&p.tm512, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dst, d2, s2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r2 != CUDA_SUCCESS) { printf("encode2 failed %d\n", (int)r2); return 2; }
    }
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000));
    for (int grid : {148, 16}) {
        for (int mode : {0, 5, 6, 7}) {
            p.mode = mode;
            k<<<grid, 128, 60000>>>(p); CK(cudaDeviceSynchronize());
            k<<<grid, 128, 60000>>>(p); CK(cudaDeviceSynchronize());
            long long h[148]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
            printf("grid %3d mode %d: %.0f cycles per iteration  (%.1f B/clk/SM)\n", grid, mode, avg / p.iters, (mode == 4 ? 2 : 1) * 16384.0 * p.iters / avg);
        }
    }
    return 0;
}
