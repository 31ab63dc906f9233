// mufu_probe.cu -- throughput of ex2.approx.ftz.f32 vs ex2.approx.f16x2 (values per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float x0 = threadIdx.x * 1e-3f - 1.f, x1 = x0 - 0.1f, x2 = x0 - 0.2f, x3 = x0 - 0.3f;
    unsigned h0 = 0xb800b900u + threadIdx.x, h1 = h0 + 7, h2 = h0 + 13, h3 = h0 + 29;   // packed negative halves
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
            x0 -= 1.f; x1 -= 1.f; x2 -= 1.f; x3 -= 1.f;
        } else {
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
            h0 |= 0x80008000u; h1 |= 0x80008000u; h2 |= 0x80008000u; h3 |= 0x80008000u;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) {
        for (int threads : {128, 256, 512}) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters); else k<1><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters); else k<1><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const double instr = 4.0 * iters * threads;             // thread-level MUFU instructions per SM
            printf("mode %s threads %3d: %.2f thread-instr/clk/SM = %.2f values/clk/SM\n", mode ? "f16x2" : "f32  ", threads,
                   instr / avg, instr / avg * (mode ? 2 : 1));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
