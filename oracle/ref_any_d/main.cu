// TEST INFRASTRUCTURE (oracle).  File-in / file-out driver around the reference's own fa2 fp32 kernels for any
// head dim in {32, 64, 128}:   ref_any_d <data_dir named B%d_H%d_S%d_D%d>
// reads Q.bin K.bin V.bin (+ dO.bin, else dO = 1), writes O.bin logsumexp.bin dQ.bin dK.bin dV.bin.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

extern "C" int ref_any_d_forward(const float*, const float*, const float*, float*, float*, int, int, int, int);
extern "C" int ref_any_d_backward(const float*, const float*, const float*, const float*, const float*, const float*,
                                  float*, float*, float*, float*, int, int, int, int);

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
    if (argc < 2) { fprintf(stderr, "usage: %s <dir B_H_S_D>\n", argv[0]); return 1; }
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
