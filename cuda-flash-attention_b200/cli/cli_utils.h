// cli_utils.h -- argument grammar, folder-name shape parsing and raw float32 .bin I/O of the
// FlashAttention CLI.  Mirrors the behaviour of the reference's src/utils.cpp:5-100 and
// include/error_utils.h:6-19 (same grammar, same messages, exit(EXIT_FAILURE) on error).
#pragma once
#include <cstddef>

namespace fa2cli {

enum class Method { Naive, FlashAttention1, FlashAttention2 };          // include/enum_types.h:3-7
enum class Mode { Forward, Backward, ForwardBackward };                  // :9-13
enum class ShmPrecision { FP16, FP32, BF16 };                            // :15-18 (+ bf16 extension)

struct Args {
    Method method;
    Mode mode;
    ShmPrecision precision;
    const char* data_path;
    int n_gpus;          // extension: optional "--gpus N" after the four positional arguments
};

[[noreturn]] void usage(const char* prog);
// Dies with usage() on any grammar error, like parse_args (src/utils.cpp:52-100).
Args parse_args(int argc, char** argv);
// "…/B2_H8_S512_D64[/]" -> dims; dies like the reference on a malformed name (src/utils.cpp:32-49).
void parse_config_string(const char* path, int* B, int* H, int* S, int* D);
bool file_exists(const char* path);
void load_binary_file(const char* path, float* dst, size_t count);      // raw native-endian float32, no header
void save_binary_file(const char* path, const float* src, size_t count);
[[noreturn]] void die(const char* what);                                  // perror + exit(EXIT_FAILURE)

}  // namespace fa2cli
