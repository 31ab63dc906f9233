// FlashAttention CLI -- drop-in for the reference's src/main.cpp:14-135 on the fa2 path:
//   FlashAttention <naive|fa1|fa2> <forward|backward|forward_backward> <fp16|fp32> <data_dir> [--gpus N]
// Reads Q.bin K.bin V.bin (+ O.bin logsumexp.bin for backward, optional dO.bin else dO = 1) from
// <data_dir> = .../B{B}_H{H}_S{S}_D{D}, writes O.bin logsumexp.bin and/or dQ.bin dK.bin dV.bin there.
// The compute goes through the C ABI of libfa2_b200.so; the batch*head slabs are split over
// --gpus devices (the reference is single-GPU).  Sizes are 64-bit (the reference's qkv_size is int,
// src/main.cpp:27).  Host buffers are pinned so H2D/D2H run at link speed.
#include <cstdio>
#include <cstdlib>
#include <string>

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
}  // namespace

int main(int argc, char** argv) {
    const Args args = parse_args(argc, argv);

    int B, H, S, D;
    parse_config_string(args.data_path, &B, &H, &S, &D);
    const size_t qkv = static_cast<size_t>(B) * H * S * D;
    const size_t rows = static_cast<size_t>(B) * H * S;

    std::printf("Batch size:    %d\n", B);
    std::printf("Num heads:     %d\n", H);
    std::printf("Sequence len:  %d\n", S);
    std::printf("Head dim:      %d\n", D);

    // Dispatcher checks of include/dispatcher.h:15-89,:131-139, done before any I/O.
    if (args.method != Method::FlashAttention2) {
        const char* name = args.method == Method::FlashAttention1 ? "Flash Attention 1" : "Vanilla Attention";
        if (args.mode != Mode::Forward)
            std::fprintf(stderr, "Error: %s backward pass not implemented\n", name);
        else
            std::fprintf(stderr, "Error: %s is a comparison baseline of the reference and is not part of this "
                                 "build (only fa2 is provided)\n", name);
        return EXIT_FAILURE;
    }
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

    float* hQ = host_buffer(qkv);
    float* hK = host_buffer(qkv);
    float* hV = host_buffer(qkv);
    float* hO = host_buffer(qkv);
    float* hL = host_buffer(rows);
    float *hdO = nullptr, *hdQ = nullptr, *hdK = nullptr, *hdV = nullptr;
    if (bwd) {
        hdO = host_buffer(qkv);
        hdQ = host_buffer(qkv);
        hdK = host_buffer(qkv);
        hdV = host_buffer(qkv);
    }

    std::printf("Loading data...\n");
    load_binary_file(q_path.c_str(), hQ, qkv);
    load_binary_file(k_path.c_str(), hK, qkv);
    load_binary_file(v_path.c_str(), hV, qkv);
    if (args.mode == Mode::Backward) {
        load_binary_file(o_path.c_str(), hO, qkv);
        load_binary_file(lse_path.c_str(), hL, rows);
    }
    if (bwd) {
        if (file_exists(do_path.c_str())) {
            load_binary_file(do_path.c_str(), hdO, qkv);
        } else {
            for (size_t i = 0; i < qkv; ++i) hdO[i] = 1.0f;    // L = sum(O)  =>  dL/dO = 1
        }
    }
    std::printf("Data loaded successfully.\n\n");

    std::printf("Running...\n");
    const int prec = args.precision == ShmPrecision::FP16 ? FA2_PRECISION_FP16
                   : args.precision == ShmPrecision::FP32 ? FA2_PRECISION_FP32 : FA2_PRECISION_BF16;
    const char* tag = args.precision == ShmPrecision::FP32 ? "" : " with 16-bit SHM precision flag";
    float ms = 0.f;
    int rc;
    if (args.mode == Mode::Forward) {
        std::printf("Running Flash Attention 2 Forward (HEAD_DIM=%d)%s on %d B200...\n", D, tag, args.n_gpus);
        rc = fa2_host_forward(hQ, hK, hV, hO, hL, B, H, S, D, prec, args.n_gpus, &ms);
    } else if (args.mode == Mode::Backward) {
        std::printf("Running Flash Attention 2 Backward (HEAD_DIM=%d)%s on %d B200...\n", D, tag, args.n_gpus);
        rc = fa2_host_backward(hQ, hK, hV, hO, hdO, hL, hdQ, hdK, hdV, B, H, S, D, prec, args.n_gpus, &ms);
    } else {
        std::printf("Running Forward+Backward Pass (HEAD_DIM=%d)%s on %d B200...\n", D, tag, args.n_gpus);
        rc = fa2_host_forward_backward(hQ, hK, hV, hdO, hO, hL, hdQ, hdK, hdV, B, H, S, D, prec, args.n_gpus, &ms);
    }
    if (rc != FA2_OK) {
        std::fprintf(stderr, "Error: %s\n", fa2_last_error());
        return EXIT_FAILURE;
    }
    std::printf("Kernel execution completed: %.4f seconds.\n\n", ms * 1e-3);

    std::printf("Saving output...\n");
    if (fwd) {
        save_binary_file(o_path.c_str(), hO, qkv);
        save_binary_file(lse_path.c_str(), hL, rows);
    }
    if (bwd) {
        save_binary_file(dq_path.c_str(), hdQ, qkv);
        save_binary_file(dk_path.c_str(), hdK, qkv);
        save_binary_file(dv_path.c_str(), hdV, qkv);
    }
    std::printf("Output saved successfully.\n");

    for (float* p : {hQ, hK, hV, hO, hL, hdO, hdQ, hdK, hdV}) fa2_host_free(p);
    return EXIT_SUCCESS;
}
