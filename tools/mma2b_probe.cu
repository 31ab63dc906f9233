// mma2b_probe.cu -- checks the two hardware behaviours the CTA-pair backward (csrc/fa2_bwd2_sm100.cu) relies on beyond
// what mma2_probe.cu covers:
//  (1) tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA), A and B both MN-major from shared memory, K = 256 issued
//      as 16 K-steps whose operand tiles sit in two separate buffers ("lo" / "hi" halves of the contraction), and the
//      TMEM layout of the result: rows m -> lanes m (columns n < N/2) and lanes 64 + m (columns n >= N/2), N/2 columns.
//  (2) cp.async.bulk.shared::cluster.shared::cta: a 16 KB tile copied from one CTA's shared memory into its peer's,
//      completion signalled on an mbarrier of the DESTINATION CTA (complete_tx), data then read by the tensor core.
// D[m][n] = sum_k A[k][m] B[k][n] with small integers (exact in fp16/fp32).  Diagnostic tool only.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../cuda-flash-attention_b200/csrc/ptx.cuh"
using namespace fa2;

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
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// A tile [128 k rows][64 m] fp16, 128B-swizzled MN-major atom (16 KB); value A[k][m]
__device__ __forceinline__ void fill_atom(uint8_t* base, int which /*0 = A, 1 = B*/, int rank, int khalf) {
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {          // 16-byte chunks: row k, chunk c (8 elements)
        const int k = i >> 3, c = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) {
            const int x = c * 8 + e;                                    // m (or n) within this CTA's 64-wide slice
            const int kk = khalf * 128 + k;
            const float f = which == 0 ? float(((kk * 3 + (64 * rank + x)) % 5) - 2) : float(((kk + 2 * (64 * rank + x)) % 7) - 3);
            v[e] = __float2half(f);
        }
        *reinterpret_cast<uint4*>(base + swz128(k, c)) = *reinterpret_cast<uint4*>(v);
    }
}

struct Params { int* errors; int* info; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA_lo = smem;                 // A, contraction rows   0..127
    uint8_t* sA_hi = smem + 16384;         // A, contraction rows 128..255  (filled by the PEER through DSMEM)
    uint8_t* sB_lo = smem + 32768;
    uint8_t* sB_hi = smem + 49152;
    uint8_t* sSend = smem + 65536;         // what this CTA sends to its peer's sA_hi / sA_lo
    __shared__ uint64_t bar_mma, bar_recv;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    // CTA r computes locally the A half of ITS contraction range (khalf = r) for its own m slice, and the A half of
    // its contraction range for the PEER's m slice, which it sends over (like dS atoms in the backward).
    fill_atom(rank == 0 ? sA_lo : sA_hi, 0, rank, rank);             // own m slice, own k half
    fill_atom(sSend, 0, rank ^ 1, rank);                              // peer's m slice, own k half
    fill_atom(sB_lo, 1, rank, 0);
    fill_atom(sB_hi, 1, rank, 1);
    if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_recv, 1); fence_mbar_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc2(&holder, 64); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync();
    const uint32_t tmem = holder;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_recv, 16384);
        // my k half of the peer's m slice goes into the peer's buffer for that k half
        uint8_t* dst_local = rank == 0 ? sA_lo : sA_hi;               // same offset in the peer = its buffer for MY k half
        dsmem_bulk_copy(cluster_map(dst_local, rank ^ 1), sSend, 16384, cluster_map(&bar_recv, rank ^ 1));
    }
    mbar_wait(&bar_recv, 0);               // the peer's tile has landed in my shared memory
    cluster_sync();                        // (probe only: both CTAs have received before the leader issues)
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128, 1, 1, 0);        // M = 128 over the pair, N = 128, A and B MN-major
            for (int ks = 0; ks < 16; ++ks) {
                uint8_t* a = (ks < 8 ? sA_lo : sA_hi) + (ks & 7) * 2048;
                uint8_t* b = (ks < 8 ? sB_lo : sB_hi) + (ks & 7) * 2048;
                umma2_ss(tmem, umma_smem_desc(smem_u32(a), 16384, 1024), umma_smem_desc(smem_u32(b), 16384, 1024), idesc, ks > 0);
            }
            umma_commit2(&bar_mma, 3);
        }
        __syncwarp();
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    {   // every lane reads its 64 columns; expected layout: lane l -> row m = l % 64 (global 64 rank + m), n = 64 (l / 64) + col
        int bad = 0;
        const int lane_g = warp * 32 + (threadIdx.x & 31);
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) {
                const int m = 64 * rank + (lane_g % 64), n = 64 * (lane_g / 64) + c * 32 + i;
                float want = 0.f;
                for (int kk = 0; kk < 256; ++kk) want += float(((kk * 3 + m) % 5) - 2) * float(((kk + 2 * n) % 7) - 3);
                if (__uint_as_float(r[i]) != want) {
                    if (bad == 0 && lane_g % 37 == 0) { p.info[0] = lane_g; p.info[1] = c * 32 + i; p.info[2] = __float_as_int(__uint_as_float(r[i])); p.info[3] = __float_as_int(want); }
                    ++bad;
                }
            }
        }
        if (bad) atomicAdd(p.errors + rank, bad);
    }
    tc_fence_before();
    cluster_sync();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 64); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

int main() {
    int* err; int* info;
    CK(cudaMalloc(&err, 8)); CK(cudaMalloc(&info, 16));
    CK(cudaMemset(err, 0, 8)); CK(cudaMemset(info, 0, 16));
    const int smem = 5 * 16384 + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    Params p{err, info};
    k<<<2, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
    int he[2], inf[4];
    CK(cudaMemcpy(he, err, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(inf, info, 16, cudaMemcpyDeviceToHost));
    printf("cta_group::2 M=128 N=128 K=256, A/B MN-major, A halves exchanged by DSMEM bulk copy: wrong elements CTA0 %d, CTA1 %d (of 8192 each)\n", he[0], he[1]);
    if (he[0] || he[1]) printf("  first mismatch: lane %d col %d got %f want %f\n", inf[0], inf[1], *reinterpret_cast<float*>(&inf[2]), *reinterpret_cast<float*>(&inf[3]));
    else printf("  TMEM layout confirmed: lanes 0-63 hold columns 0..63, lanes 64-127 hold columns 64..127 of rows 64 rank + (lane %% 64)\n");
    return (he[0] || he[1]) ? 1 : 0;
}
