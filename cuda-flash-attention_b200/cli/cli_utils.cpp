#include "cli_utils.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>

namespace fa2cli {

void usage(const char* prog) {
    std::fprintf(stderr,
                 "USAGE: %s <computation_method:naive|fa1|fa2> <mode:forward|backward|forward_backward> "
                 "<SHM_precision:fp16|fp32> <data_folder_path> [--gpus N]\n",
                 prog);
    std::exit(EXIT_FAILURE);
}

void die(const char* what) {
    std::perror(what);
    std::exit(EXIT_FAILURE);
}

Args parse_args(int argc, char** argv) {
    if (argc < 5) usage(argv[0]);
    Args a{};
    const std::string method = argv[1], mode = argv[2], prec = argv[3];
    if (method == "fa2") a.method = Method::FlashAttention2;
    else if (method == "fa1") a.method = Method::FlashAttention1;
    else if (method == "naive") a.method = Method::Naive;
    else usage(argv[0]);

    if (mode == "forward") a.mode = Mode::Forward;
    else if (mode == "backward") a.mode = Mode::Backward;
    else if (mode == "forward_backward" || mode == "both" || mode == "forward-backward")   // README aliases
        a.mode = Mode::ForwardBackward;
    else usage(argv[0]);

    if (prec == "fp16") a.precision = ShmPrecision::FP16;
    else if (prec == "fp32") a.precision = ShmPrecision::FP32;
    else if (prec == "bf16") a.precision = ShmPrecision::BF16;
    else usage(argv[0]);

    a.data_path = argv[4];
    a.n_gpus = 1;
    if (const char* env = std::getenv("FA2_NUM_GPUS")) a.n_gpus = std::atoi(env);
    for (int i = 5; i < argc; ++i) {
        if (std::strcmp(argv[i], "--gpus") == 0 && i + 1 < argc) a.n_gpus = std::atoi(argv[++i]);
        else usage(argv[0]);
    }
    if (a.n_gpus < 1) usage(argv[0]);
    return a;
}

void parse_config_string(const char* path, int* B, int* H, int* S, int* D) {
    std::string p(path);
    while (!p.empty() && p.back() == '/') p.pop_back();
    const size_t slash = p.find_last_of('/');
    const std::string leaf = (slash == std::string::npos) ? p : p.substr(slash + 1);
    if (std::sscanf(leaf.c_str(), "B%d_H%d_S%d_D%d", B, H, S, D) != 4) die("sscanf");
}

bool file_exists(const char* path) {
    struct stat st;
    return ::stat(path, &st) == 0;
}

void load_binary_file(const char* path, float* dst, size_t count) {
    FILE* f = std::fopen(path, "rb");
    if (!f) die("fopen");
    const size_t got = std::fread(dst, sizeof(float), count, f);
    std::fclose(f);
    if (got != count) die("fread");
}

void save_binary_file(const char* path, const float* src, size_t count) {
    FILE* f = std::fopen(path, "wb");
    if (!f) die("fopen");
    const size_t put = std::fwrite(src, sizeof(float), count, f);
    std::fclose(f);
    if (put != count) die("fwrite");
}

}  // namespace fa2cli
