// TEST INFRASTRUCTURE (oracle).  File-in / file-out driver around the reference's own fa2 fp32 kernels for any
// head dim in {32, 64, 128}:   ref_any_d <data_dir named B%d_H%d_S%d_D%d>
// reads Q.bin K.bin V.bin (+ dO.bin, else dO = 1), writes O.bin logsumexp.bin dQ.bin dK.bin dV.bin.
// Bench mode (bench.py --impl reference):   ref_any_d --bench B H S D steps warmup [budget_seconds]
// runs the reference's own host launchers (host_flash_attention2_forward / _backward: malloc, H2D from pageable
// memory, launches timed by its TimerGPU, D2H, free -- what dispatch_forward_backward does, include/dispatcher.h:91-104)
// on synthetic data and prints one line "REF_BENCH {json}" with the TimerGPU kernel time and the wall time per step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>
#include "timer.h"      // the reference's TimerGPU / TimerManager (include/timer.h)

extern "C" int ref_any_d_forward(const float*, const float*, const float*, float*, float*, int, int, int, int);
extern "C" int ref_any_d_backward(const float*, const float*, const float*, const float*, const float*, const float*,
                                  float*, float*, float*, float*, int, int, int, int);

int ref_any_d_host_forward(const float*, const float*, const float*, float*, float*, int, int, int, int, TimerManager*);
int ref_any_d_host_backward(const float*, const float*, const float*, const float*, const float*, const float*, float*,
                            float*, float*, int, int, int, int, TimerManager*);

// cheap deterministic uniform(-1.7, 1.7) fill (unit variance); the reference kernels have no data-dependent control flow
static void fill(float* p, size_t n, uint64_t seed) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    for (size_t i = 0; i < n; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        p[i] = ((float)(uint32_t)(x >> 40) * (1.0f / 8388608.0f) - 1.0f) * 1.7320508f;
    }
}

static int bench(int B, int H, int S, int D, int steps, int warmup, double budget_s) {
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const int BH_full = B * H;
    int bh = BH_full;                       // slabs per step; shrinks (same S, D) if the run would exceed the budget
    const size_t slab = (size_t)S * D;
    std::vector<float> q(slab * bh), k(slab * bh), v(slab * bh), g(slab * bh), o(slab * bh), dq(slab * bh), dk(slab * bh),
        dv(slab * bh), lse((size_t)S * bh);
    fill(q.data(), q.size(), 1); fill(k.data(), k.size(), 2); fill(v.data(), v.size(), 3); fill(g.data(), g.size(), 4);
    TimerGPU gpu_timer;
    TimerManager tm;
    tm.SetTimer(&gpu_timer);
    FILE* real_stdout = stdout;
    double kernel_ms = 0, wall_s = 0;
    int timed = 0;
    const auto t_begin = clk::now();
    for (int i = 0; i < warmup + steps; ++i) {
        tm.Reset();
        stdout = fopen("/dev/null", "w");                  // the launchers print ~15 lines per call
        const auto t0 = clk::now();
        int rc = ref_any_d_host_forward(q.data(), k.data(), v.data(), o.data(), lse.data(), 1, bh, S, D, &tm);
        if (!rc) rc = ref_any_d_host_backward(q.data(), k.data(), v.data(), o.data(), g.data(), lse.data(), dq.data(),
                                              dk.data(), dv.data(), 1, bh, S, D, &tm);
        const auto t1 = clk::now();
        fclose(stdout);
        stdout = real_stdout;
        if (rc) { fprintf(stderr, "reference host launcher failed (%d)\n", rc); return 1; }
        if (i >= warmup) { kernel_ms += tm.TotalElapsedMillis(); wall_s += secs(t0, t1); ++timed; }
        if (i == 0) {
            // bound the whole run: if (steps + warmup) steps of this size would not fit, go on with fewer slabs
            const double per_step = secs(t0, t1), want = per_step * (warmup + steps);
            if (want > budget_s && bh > 8) {
                int nb = (int)(bh * budget_s / want);
                if (nb < 8) nb = 8;
                if (nb < bh) { bh = nb; if (warmup == 0) { kernel_ms = 0; wall_s = 0; timed = 0; ++steps; } }
            }
        }
    }
    const double f = 14.0 * bh * (double)S * S * D;        // fwd 4 BHS^2D + bwd 10 BHS^2D
    printf("REF_BENCH {\"B\": %d, \"H\": %d, \"S\": %d, \"D\": %d, \"slabs_per_step\": %d, \"slabs_full\": %d, \"steps\": %d, "
           "\"kernel_ms_per_step\": %.4f, \"wall_ms_per_step\": %.3f, \"kernel_tflops\": %.4f, \"e2e_tflops\": %.4f, "
           "\"total_wall_s\": %.2f}\n",
           B, H, S, D, bh, BH_full, timed, kernel_ms / timed, wall_s * 1e3 / timed, f / (kernel_ms / timed * 1e-3) / 1e12,
           f / (wall_s / timed) / 1e12, secs(t_begin, clk::now()));
    return 0;
}

static std::vector<float> load(const std::string& p, size_t n, bool optional = false) {
    std::vector<float> v(n, 1.0f);
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { if (optional) return v; perror(p.c_str()); exit(1); }
    if (fread(v.data(), 4, n, f) != n) { fprintf(stderr, "short file %s\n", p.c_str()); exit(1); }
    fclose(f);
    return v;
}
static void save(const std::string& p, const float* d, size_t n) {
    std::vector<float> h(n);
    cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
    FILE* f = fopen(p.c_str(), "wb");
    fwrite(h.data(), 4, n, f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <dir B_H_S_D> | --bench B H S D steps warmup [budget_s]\n", argv[0]); return 1; }
    if (strcmp(argv[1], "--bench") == 0) {
        if (argc < 8) { fprintf(stderr, "usage: %s --bench B H S D steps warmup [budget_s]\n", argv[0]); return 1; }
        return bench(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]),
                     argc > 8 ? atof(argv[8]) : 150.0);
    }
    std::string dir(argv[1]);
    while (!dir.empty() && dir.back() == '/') dir.pop_back();
    const std::string leaf = dir.substr(dir.find_last_of('/') + 1);
    int B, H, S, D;
    if (sscanf(leaf.c_str(), "B%d_H%d_S%d_D%d", &B, &H, &S, &D) != 4) { fprintf(stderr, "bad dir name\n"); return 1; }
    const size_t n = (size_t)B * H * S * D, nl = (size_t)B * H * S;
    auto hq = load(dir + "/Q.bin", n), hk = load(dir + "/K.bin", n), hv = load(dir + "/V.bin", n);
    auto hg = load(dir + "/dO.bin", n, true);
    float *q, *k, *v, *o, *g, *lse, *dvec, *dq, *dk, *dv;
    for (float** p : {&q, &k, &v, &o, &g, &dq, &dk, &dv}) cudaMalloc(p, n * 4);
    cudaMalloc(&lse, nl * 4); cudaMalloc(&dvec, nl * 4);
    cudaMemcpy(q, hq.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(k, hk.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v, hv.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(g, hg.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemset(dq, 0, n * 4); cudaMemset(dk, 0, n * 4); cudaMemset(dv, 0, n * 4);   // f-attn2-backward.cu:427-429
    int rc = ref_any_d_forward(q, k, v, o, lse, B, H, S, D);
    if (!rc) rc = ref_any_d_backward(q, k, v, o, g, lse, dvec, dq, dk, dv, B, H, S, D);
    if (rc || cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "reference kernels failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    save(dir + "/O.bin", o, n); save(dir + "/logsumexp.bin", lse, nl);
    save(dir + "/dQ.bin", dq, n); save(dir + "/dK.bin", dk, n); save(dir + "/dV.bin", dv, n);
    printf("ref_any_d: B%d H%d S%d D%d done\n", B, H, S, D);
    return 0;
}
