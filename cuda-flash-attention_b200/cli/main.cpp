// FlashAttention CLI -- drop-in for the reference's src/main.cpp:14-135 on the fa2 path:
//   FlashAttention <naive|fa1|fa2> <forward|backward|forward_backward> <fp16|fp32> <data_dir> [--gpus N]
// Reads Q.bin K.bin V.bin (+ O.bin logsumexp.bin for backward, optional dO.bin else dO = 1) from
// <data_dir> = .../B{B}_H{H}_S{S}_D{D}, writes O.bin logsumexp.bin and/or dQ.bin dK.bin dV.bin there.
// The compute goes through the C ABI of libfa2_b200.so; the batch*head slabs are split over
// --gpus devices (the reference is single-GPU).  Sizes are 64-bit (the reference's qkv_size is int,
// src/main.cpp:27).
//
// File I/O is part of the pipeline: the reference loads every file, computes, then saves (src/main.cpp:74-118), and at
// 0.5-4 GiB per tensor that serial fread / fwrite is most of the wall time.  Here the (b,h) slabs -- contiguous byte
// ranges of every file -- are cut into chunks that flow through three pinned buffer sets: a reader thread preads chunk
// c+1 (one thread per file), the main thread runs chunk c through fa2_host_* (H2D / kernels / D2H overlapped inside the
// library), a writer thread pwrites chunk c-1.  FA2_CLI_STREAM=0 selects the serial load-all / compute / save-all order.
//
// Methods fa1 and naive are comparison baselines of the reference, not part of this library: when the reference's own CLI
// is available ($FA2_BASELINE_CLI, or oracle/_ref/FlashAttention_ref next to this tree) the call is handed over to it.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "cli_utils.h"
#include "fa2_b200.h"

using namespace fa2cli;

namespace {

float* host_buffer(size_t count) {
    float* p = static_cast<float*>(fa2_host_alloc(count * sizeof(float)));
    if (!p) {
        std::fprintf(stderr, "Error: host allocation of %zu bytes failed: %s\n", count * sizeof(float), fa2_last_error());
        std::exit(EXIT_FAILURE);
    }
    return p;
}

struct Bufs {                       // one set of host buffers (a whole problem, or one chunk of it)
    float *Q = nullptr, *K = nullptr, *V = nullptr, *O = nullptr, *L = nullptr, *dO = nullptr, *dQ = nullptr, *dK = nullptr,
          *dV = nullptr;
    void alloc(size_t qkv, size_t rows, bool bwd) {
        Q = host_buffer(qkv); K = host_buffer(qkv); V = host_buffer(qkv); O = host_buffer(qkv); L = host_buffer(rows);
        if (bwd) { dO = host_buffer(qkv); dQ = host_buffer(qkv); dK = host_buffer(qkv); dV = host_buffer(qkv); }
    }
    void release() {
        for (float* p : {Q, K, V, O, L, dO, dQ, dK, dV}) fa2_host_free(p);
    }
};

int run_mode(Mode mode, const Bufs& b, int B, int H, int S, int D, int prec, int n_gpus, float* ms) {
    if (mode == Mode::Forward) return fa2_host_forward(b.Q, b.K, b.V, b.O, b.L, B, H, S, D, prec, n_gpus, ms);
    if (mode == Mode::Backward) return fa2_host_backward(b.Q, b.K, b.V, b.O, b.dO, b.L, b.dQ, b.dK, b.dV, B, H, S, D, prec, n_gpus, ms);
    return fa2_host_forward_backward(b.Q, b.K, b.V, b.dO, b.O, b.L, b.dQ, b.dK, b.dV, B, H, S, D, prec, n_gpus, ms);
}

// ---- positional file I/O (64-bit offsets, loops over short transfers) ------------------------------------------------
void pread_all(int fd, void* dst, size_t bytes, off_t off) {
    char* p = static_cast<char*>(dst);
    while (bytes) {
        const ssize_t n = ::pread(fd, p, bytes, off);
        if (n <= 0) die("fread");                 // same message as a short fread in the reference (src/utils.cpp:17)
        p += n; off += n; bytes -= static_cast<size_t>(n);
    }
}
void pwrite_all(int fd, const void* src, size_t bytes, off_t off) {
    const char* p = static_cast<const char*>(src);
    while (bytes) {
        const ssize_t n = ::pwrite(fd, p, bytes, off);
        if (n <= 0) die("fwrite");
        p += n; off += n; bytes -= static_cast<size_t>(n);
    }
}
int open_in(const std::string& path, size_t need_bytes) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) die("fopen");
    struct stat st;
    if (::fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) < need_bytes) die("fread");     // short file
    return fd;
}
int open_out(const std::string& path) {
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) die("fopen");
    return fd;
}

// ---- the streamed pipeline ---------------------------------------------------------------------------------------------
struct Stage {                      // per buffer set: 0 = free, 1 = filled by the reader, 2 = computed
    std::mutex mu;
    std::condition_variable cv;
    int state = 0;
    void wait_for(int s) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state == s; }); }
    void set(int s) { { std::lock_guard<std::mutex> lk(mu); state = s; } cv.notify_all(); }
};

struct FileSet { int q = -1, k = -1, v = -1, o = -1, l = -1, g = -1, dq = -1, dk = -1, dv = -1; };

// Runs `jobs` concurrently (one thread per file: page-cache copies are CPU-bound, the files are independent).
template <class F>
void parallel(const std::vector<F>& jobs) {
    std::vector<std::thread> th;
    for (size_t i = 1; i < jobs.size(); ++i) th.emplace_back(jobs[i]);
    if (!jobs.empty()) jobs[0]();
    for (auto& t : th) t.join();
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// argv[0]-relative default for the reference's own CLI (reported baselines fa1 / naive)
std::string baseline_cli_path() {
    if (const char* env = std::getenv("FA2_BASELINE_CLI")) return env;
    char self[4096];
    const ssize_t n = ::readlink("/proc/self/exe", self, sizeof(self) - 1);
    if (n <= 0) return "";
    self[n] = 0;
    std::string dir(self);
    dir = dir.substr(0, dir.find_last_of('/'));
    return dir + "/../oracle/_ref/FlashAttention_ref";
}

}  // namespace

int main(int argc, char** argv) {
    const Args args = parse_args(argc, argv);

    int B, H, S, D;
    parse_config_string(args.data_path, &B, &H, &S, &D);
    const size_t slab = static_cast<size_t>(S) * D;                  // floats per (b,h) slab
    const int BH = B * H;
    const size_t qkv = static_cast<size_t>(BH) * slab;
    const size_t rows = static_cast<size_t>(BH) * S;

    // Dispatcher checks of include/dispatcher.h:15-89,:131-139, done before any I/O.
    if (args.method != Method::FlashAttention2) {
        const char* name = args.method == Method::FlashAttention1 ? "Flash Attention 1" : "Vanilla Attention";
        if (args.mode != Mode::Forward) {
            std::printf("Batch size:    %d\nNum heads:     %d\nSequence len:  %d\nHead dim:      %d\n", B, H, S, D);
            std::fprintf(stderr, "Error: %s backward pass not implemented\n", name);
            return EXIT_FAILURE;
        }
        const std::string exe = baseline_cli_path();
        if (!exe.empty() && ::access(exe.c_str(), X_OK) == 0) {
            std::fprintf(stderr, "note: %s is a comparison baseline of the reference; running its own CUDA-core kernel through %s\n",
                         name, exe.c_str());
            char* fwd_argv[] = {const_cast<char*>(exe.c_str()), argv[1], argv[2], argv[3], argv[4], nullptr};
            ::execv(exe.c_str(), fwd_argv);
        }
        std::printf("Batch size:    %d\nNum heads:     %d\nSequence len:  %d\nHead dim:      %d\n", B, H, S, D);
        std::fprintf(stderr, "Error: %s is a comparison baseline of the reference and is not part of this build (only fa2 is "
                             "provided; set FA2_BASELINE_CLI to the reference's own CLI to run it)\n", name);
        return EXIT_FAILURE;
    }

    std::printf("Batch size:    %d\n", B);
    std::printf("Num heads:     %d\n", H);
    std::printf("Sequence len:  %d\n", S);
    std::printf("Head dim:      %d\n", D);
    if (D != 32 && D != 64 && D != 128) {
        std::fprintf(stderr, "Error: Unsupported head dimension %d\n", D);
        return EXIT_FAILURE;
    }

    const bool fwd = args.mode != Mode::Backward, bwd = args.mode != Mode::Forward;
    const std::string dir(args.data_path);
    const std::string q_path = dir + "/Q.bin", k_path = dir + "/K.bin", v_path = dir + "/V.bin";
    const std::string o_path = dir + "/O.bin", lse_path = dir + "/logsumexp.bin", do_path = dir + "/dO.bin";
    const std::string dq_path = dir + "/dQ.bin", dk_path = dir + "/dK.bin", dv_path = dir + "/dV.bin";

    bool ok = file_exists(q_path.c_str()) && file_exists(k_path.c_str()) && file_exists(v_path.c_str());
    if (args.mode == Mode::Backward) ok = ok && file_exists(o_path.c_str()) && file_exists(lse_path.c_str());
    if (!ok) die("Data files not found.\n");
    const bool have_dO = bwd && file_exists(do_path.c_str());

    const int prec = args.precision == ShmPrecision::FP16 ? FA2_PRECISION_FP16
                   : args.precision == ShmPrecision::FP32 ? FA2_PRECISION_FP32 : FA2_PRECISION_BF16;
    const char* tag = args.precision == ShmPrecision::FP32 ? " (fp16 tensor-core operands, fp32 range and accumulation)"
                    : args.precision == ShmPrecision::FP16 ? " with 16-bit SHM precision flag" : " with bf16 operands";
    const char* what = args.mode == Mode::Forward ? "Flash Attention 2 Forward"
                     : args.mode == Mode::Backward ? "Flash Attention 2 Backward" : "Forward+Backward Pass";

    const char* stream_env = std::getenv("FA2_CLI_STREAM");
    const bool stream = !(stream_env && std::atoi(stream_env) == 0);
    float total_ms = 0.f;
    const double t_start = now_s();

    if (!stream) {
        // ---- the reference's order: load everything, compute, save everything (src/main.cpp:74-118)
        Bufs b;
        b.alloc(qkv, rows, bwd);
        std::printf("Loading data...\n");
        load_binary_file(q_path.c_str(), b.Q, qkv);
        load_binary_file(k_path.c_str(), b.K, qkv);
        load_binary_file(v_path.c_str(), b.V, qkv);
        if (args.mode == Mode::Backward) {
            load_binary_file(o_path.c_str(), b.O, qkv);
            load_binary_file(lse_path.c_str(), b.L, rows);
        }
        if (bwd) {
            if (have_dO) load_binary_file(do_path.c_str(), b.dO, qkv);
            else std::fill(b.dO, b.dO + qkv, 1.0f);                   // L = sum(O)  =>  dL/dO = 1
        }
        std::printf("Data loaded successfully.\n\n");
        std::printf("Running...\n");
        std::printf("Running %s (HEAD_DIM=%d)%s on %d B200...\n", what, D, tag, args.n_gpus);
        if (run_mode(args.mode, b, B, H, S, D, prec, args.n_gpus, &total_ms) != FA2_OK) {
            std::fprintf(stderr, "Error: %s\n", fa2_last_error());
            return EXIT_FAILURE;
        }
        std::printf("Kernel execution completed: %.4f seconds.\n\n", total_ms * 1e-3);
        std::printf("Saving output...\n");
        if (fwd) {
            save_binary_file(o_path.c_str(), b.O, qkv);
            save_binary_file(lse_path.c_str(), b.L, rows);
        }
        if (bwd) {
            save_binary_file(dq_path.c_str(), b.dQ, qkv);
            save_binary_file(dk_path.c_str(), b.dK, qkv);
            save_binary_file(dv_path.c_str(), b.dV, qkv);
        }
        std::printf("Output saved successfully.\n");
        b.release();
        std::printf("Total wall time: %.3f seconds (serial load / compute / save).\n", now_s() - t_start);
        return EXIT_SUCCESS;
    }

    // ---- streamed: chunks of slabs through three buffer sets
    // ~32 MiB per tensor per chunk (FA2_CLI_CHUNK_MB overrides): three sets of nine such buffers are pinned, and pinning
    // costs about as much per byte as reading the page cache does
    size_t chunk_mb = 32;
    if (const char* e = std::getenv("FA2_CLI_CHUNK_MB")) chunk_mb = std::max(1, std::atoi(e));
    size_t chunk_slabs = (chunk_mb << 20) / (slab * sizeof(float));
    const size_t per_gpu_min = static_cast<size_t>(args.n_gpus) * 4;           // every device still gets a few slabs
    if (chunk_slabs < per_gpu_min) chunk_slabs = per_gpu_min;
    if (chunk_slabs < 1) chunk_slabs = 1;
    if (chunk_slabs > static_cast<size_t>(BH)) chunk_slabs = BH;
    const int n_chunks = static_cast<int>((BH + chunk_slabs - 1) / chunk_slabs);
    constexpr int kSets = 3;
    const int n_sets = n_chunks < kSets ? n_chunks : kSets;

    FileSet f;
    f.q = open_in(q_path, qkv * 4); f.k = open_in(k_path, qkv * 4); f.v = open_in(v_path, qkv * 4);
    if (args.mode == Mode::Backward) { f.o = open_in(o_path, qkv * 4); f.l = open_in(lse_path, rows * 4); }
    if (have_dO) f.g = open_in(do_path, qkv * 4);

    Bufs set[kSets];
    Stage stage[kSets];
    for (int i = 0; i < n_sets; ++i) {
        set[i].alloc(chunk_slabs * slab, chunk_slabs * S, bwd);
        if (bwd && !have_dO) std::fill(set[i].dO, set[i].dO + chunk_slabs * slab, 1.0f);   // L = sum(O)  =>  dL/dO = 1
    }
    // outputs are created only once the inputs have been opened and checked (backward reads O.bin / logsumexp.bin)
    if (fwd) { f.o = open_out(o_path); f.l = open_out(lse_path); }
    if (bwd) { f.dq = open_out(dq_path); f.dk = open_out(dk_path); f.dv = open_out(dv_path); }

    std::printf("Loading data...\n");
    auto chunk_range = [&](int c, size_t* s0, size_t* cnt) {
        *s0 = static_cast<size_t>(c) * chunk_slabs;
        *cnt = std::min(chunk_slabs, static_cast<size_t>(BH) - *s0);
    };
    using Job = std::function<void()>;
    std::thread reader([&] {
        for (int c = 0; c < n_chunks; ++c) {
            Bufs& b = set[c % n_sets];
            stage[c % n_sets].wait_for(0);
            size_t s0, cnt;
            chunk_range(c, &s0, &cnt);
            const size_t nb = cnt * slab * 4, nl = cnt * S * 4;
            const off_t off = static_cast<off_t>(s0 * slab * 4), offl = static_cast<off_t>(s0 * S * 4);
            std::vector<Job> jobs = {[&] { pread_all(f.q, b.Q, nb, off); }, [&] { pread_all(f.k, b.K, nb, off); },
                                     [&] { pread_all(f.v, b.V, nb, off); }};
            if (args.mode == Mode::Backward) {
                jobs.push_back([&] { pread_all(f.o, b.O, nb, off); });
                jobs.push_back([&] { pread_all(f.l, b.L, nl, offl); });
            }
            if (have_dO) jobs.push_back([&] { pread_all(f.g, b.dO, nb, off); });
            parallel(jobs);
            stage[c % n_sets].set(1);
        }
    });
    std::thread writer([&] {
        for (int c = 0; c < n_chunks; ++c) {
            Bufs& b = set[c % n_sets];
            stage[c % n_sets].wait_for(2);
            size_t s0, cnt;
            chunk_range(c, &s0, &cnt);
            const size_t nb = cnt * slab * 4, nl = cnt * S * 4;
            const off_t off = static_cast<off_t>(s0 * slab * 4), offl = static_cast<off_t>(s0 * S * 4);
            std::vector<Job> jobs;
            if (fwd) {
                jobs.push_back([&] { pwrite_all(f.o, b.O, nb, off); });
                jobs.push_back([&] { pwrite_all(f.l, b.L, nl, offl); });
            }
            if (bwd) {
                jobs.push_back([&] { pwrite_all(f.dq, b.dQ, nb, off); });
                jobs.push_back([&] { pwrite_all(f.dk, b.dK, nb, off); });
                jobs.push_back([&] { pwrite_all(f.dv, b.dV, nb, off); });
            }
            parallel(jobs);
            stage[c % n_sets].set(0);
        }
    });
    int rc = FA2_OK;
    std::string err;
    for (int c = 0; c < n_chunks; ++c) {
        stage[c % n_sets].wait_for(1);
        size_t s0, cnt;
        chunk_range(c, &s0, &cnt);
        float ms = 0.f;
        if (rc == FA2_OK) {
            rc = run_mode(args.mode, set[c % n_sets], 1, static_cast<int>(cnt), S, D, prec, args.n_gpus, &ms);
            if (rc != FA2_OK) err = fa2_last_error();
        }
        total_ms += ms;
        stage[c % n_sets].set(2);             // (after a failure the chunks still drain so that both threads finish)
    }
    reader.join();
    writer.join();
    for (int fd : {f.q, f.k, f.v, f.o, f.l, f.g, f.dq, f.dk, f.dv}) if (fd >= 0) ::close(fd);
    if (rc != FA2_OK) {
        std::fprintf(stderr, "Error: %s\n", err.c_str());
        return EXIT_FAILURE;
    }
    std::printf("Data loaded successfully.\n\n");
    std::printf("Running...\n");
    std::printf("Running %s (HEAD_DIM=%d)%s on %d B200...\n", what, D, tag, args.n_gpus);
    std::printf("Kernel execution completed: %.4f seconds.\n\n", total_ms * 1e-3);
    std::printf("Saving output...\n");
    std::printf("Output saved successfully.\n");
    for (int i = 0; i < n_sets; ++i) set[i].release();
    std::printf("Total wall time: %.3f seconds (read / compute / write streamed in %d chunk(s) of up to %zu slabs).\n",
                now_s() - t_start, n_chunks, chunk_slabs);
    return EXIT_SUCCESS;
}
