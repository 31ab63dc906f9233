/*
 * fa2_b200.h -- C ABI of libfa2_b200.so, the B200-native (sm_100a) drop-in for the
 * FlashAttention-2 forward/backward path of detker/CUDA-Flash-Attention.
 *
 * Plain pointers and sizes only; no C++/torch types; no exit() -- every entry point returns
 * 0 on success or an FA2_ERR_* code, with a message available from fa2_last_error().
 * Tensors are fp32, row-major contiguous [B,H,S,D]; logsumexp (natural log) is [B,H,S]
 * (reference layout: src/main.cpp:27-38, kernels/kernel_fa2_optimized.cu:56-58,:340-343).
 * Supported head dims: 32, 64, 128 (the reference dispatches 32 and 64 only,
 * include/dispatcher.h:226-227).  There is no CPU fallback: without a B200 the compute
 * entry points fail with FA2_ERR_CUDA.
 */
#ifndef FA2_B200_H
#define FA2_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA2_OK 0
#define FA2_ERR_INVALID_ARGUMENT 1   /* null pointer, non-positive dim, unsupported head dim */
#define FA2_ERR_CUDA 2               /* a CUDA runtime/driver call failed (see fa2_last_error) */
#define FA2_ERR_UNSUPPORTED 3        /* valid request this build cannot serve (e.g. method fa1/naive) */
#define FA2_ERR_IO 4                 /* CLI layer: missing / short .bin file */

/* <SHM_precision> of the reference CLI (include/enum_types.h:15-18).  Both values run fp16
 * tensor-core operands with fp32 accumulation and per-tensor power-of-two range scaling decided on
 * the device for every launch (any fp32 input magnitude is accepted, as by the reference's fp32
 * kernels); FP16 is the reference's "fp16 shared-memory mode".  Contract: O, dQ, dK, dV within 1e-2
 * and logsumexp within 1e-3 (max-abs) of the reference's fp32 kernels.
 * BF16 is an extension OUTSIDE that contract (8-bit mantissa: O 1e-2, logsumexp 5e-3, gradients
 * 3e-2 on randn data). */
#define FA2_PRECISION_FP16 0
#define FA2_PRECISION_FP32 1
#define FA2_PRECISION_BF16 2

/* mode for fa2_workspace_bytes (include/enum_types.h:9-13) */
#define FA2_MODE_FORWARD 0
#define FA2_MODE_BACKWARD 1
#define FA2_MODE_FORWARD_BACKWARD 2

int fa2_version(void);
/* Thread-local text of the last failure on the calling thread ("" if none). */
const char* fa2_last_error(void);
/* Device scratch (16-bit operand copies, D_i, log2-domain LSE) the library keeps per device. */
size_t fa2_workspace_bytes(int B, int H, int S, int D, int mode);

/* ---------------------------------------------------------------------------------------
 * Device-pointer entry points.  Replace the reference's kernel-level ABI
 *   flash_attention2_forward_kernel_wrapper   kernels/kernel_fa2_optimized.cu:428-444
 *   D_computation_reduction_kernel_wrapper    kernels/f-attn2-backward.cu:514-528
 *   flash_attention2_backward_kernel_wrapper  kernels/f-attn2-backward.cu:491-512
 * as called by the harness (test_flash_attention2.py:278-289, :504-535).  All pointers are
 * device pointers on the current device, caller-owned and 16-byte aligned.  The library keeps ONE scratch
 * workspace per device: calls on the same device must be stream-ordered with respect to each other (one
 * stream, or event dependencies); different devices are independent and may be driven from different threads.  Asynchronous on `cuda_stream`
 * (a cudaStream_t, NULL = default stream).  fa2_backward computes D_i = rowsum(dO*O) and
 * zero-fills dQ itself (the reference needs a separate launch and three fill(0)s).
 * ------------------------------------------------------------------------------------- */
int fa2_forward(const float* Q, const float* K, const float* V, float* O, float* LSE,
                int B, int H, int S, int D, int precision, void* cuda_stream);

int fa2_backward(const float* Q, const float* K, const float* V, const float* O, const float* dO,
                 const float* LSE, float* dQ, float* dK, float* dV,
                 int B, int H, int S, int D, int precision, void* cuda_stream);

/* forward then backward without the reference's device->host->device round trip of O/LSE
 * (include/dispatcher.h:91-104) and with one shared 16-bit copy of Q, K, V. */
int fa2_forward_backward(const float* Q, const float* K, const float* V, const float* dO,
                         float* O, float* LSE, float* dQ, float* dK, float* dV,
                         int B, int H, int S, int D, int precision, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Host-pointer entry points.  Replace host_flash_attention2_forward<D> / _backward<D>
 * (+ _fp16 twins; kernels/f-attn2.cuh:13-71, called from include/dispatcher.h:24-27,:66-71).
 * NOTE the argument order here is (B, H, S, D); the reference templates take (B, S, H).
 * Pointers are host memory, caller-owned; the call is synchronous.  The batch*head slabs are
 * split over devices 0..n_gpus-1 (n_gpus <= 1: the caller's current device only); each device's share is
 * independent so there is no collective.  *kernel_ms (optional) receives the max over devices
 * of the device-side time of everything between fp32 inputs and fp32 outputs in device
 * memory (what the reference's TimerGPU measures, include/timer.h:50-64, plus our pre-passes).
 * ------------------------------------------------------------------------------------- */
int fa2_host_forward(const float* Q, const float* K, const float* V, float* O, float* LSE,
                     int B, int H, int S, int D, int precision, int n_gpus, float* kernel_ms);

int fa2_host_backward(const float* Q, const float* K, const float* V, const float* O, const float* dO,
                      const float* LSE, float* dQ, float* dK, float* dV,
                      int B, int H, int S, int D, int precision, int n_gpus, float* kernel_ms);

int fa2_host_forward_backward(const float* Q, const float* K, const float* V, const float* dO,
                              float* O, float* LSE, float* dQ, float* dK, float* dV,
                              int B, int H, int S, int D, int precision, int n_gpus, float* kernel_ms);

/* The host-side partitioner: slab range [*bh0, *bh0 + *count) of part `part` out of `n_parts`
 * over BH = B*H independent (batch, head) slabs.  Pure arithmetic, no GPU needed. */
int fa2_partition(int BH, int n_parts, int part, int* bh0, int* count);

/* The host pipeline's chunking of one device's share of `count` slabs (mode = FA2_MODE_*): writes up to
 * max_chunks chunk sizes (in slabs, in processing order) and returns the number of chunks (-1 on bad
 * arguments).  The path is PCIe-bound: small chunks at both ends shorten the one-directional head and tail,
 * chunks of ~20 MiB per tensor copy in between keep launches and the D2H backlog low, with a geometric ramp
 * from one to the other; every chunk is large enough for the kernels to keep up with the copies.
 * Pure arithmetic, no GPU needed. */
int fa2_plan_chunks(int count, int S, int D, int mode, int* sizes, int max_chunks);

/* How the host-pointer entry points lay BH = B*H slabs of S rows over n_gpus devices: *g_bh groups of devices
 * split the slabs (fa2_partition), the *g_s devices of a group split the ROWS of the group's slabs
 * (*g_bh * *g_s <= n_gpus).  *g_s == 1 is the plain slab split (the usual case); a sequence split is chosen when
 * there are fewer slabs than devices or an uneven handful (override: FA2_SEQ_SPLIT=<g_s>).  With *g_s > 1 the
 * forward runs on a range of query rows per device (K/V replicated), the backward on the same range of key/value
 * rows, and the group's partial dQ are summed by a reduce-scatter over NVLink -- the only collective of the design.
 * fa2_seq_range gives part `part`'s row range [*r0, *r1) (boundaries on multiples of 256 rows).  Pure arithmetic. */
int fa2_plan_split(int BH, int S, int n_gpus, int* g_bh, int* g_s);
int fa2_seq_range(int S, int parts, int part, int* r0, int* r1);

/* Number of CUDA devices visible to the library (0 without a GPU; never fails). */
int fa2_device_count(void);

/* Page-locked host memory for the host-pointer entry points and the CLI (falls back to malloc
 * when no CUDA device is present so that argument/IO errors are still reported on a CPU box).
 * Free with fa2_host_free. */
void* fa2_host_alloc(size_t bytes);
void fa2_host_free(void* p);

/* Per-kernel device timing for bench.py's roofline leg.  When enabled, the device-pointer entry
 * points bracket each of their kernels with cudaEvents on the launch stream.  fa2_profile_read
 * waits for the recorded events, ADDS the elapsed milliseconds since the previous read into
 * ms[4] = {cast pre-pass, forward kernel, backward pre-pass, backward kernel} and the number of
 * launches into launches[4], then clears the record.  Not thread-safe; off by default. */
int fa2_profile_enable(int on);
int fa2_profile_read(float* ms, int* launches);
/* Number of KERNELS launched inside profiled spans since the previous call (a span can hold more than one:
 * the Q/K/V cast of the large path is the cast plus the early-exit re-cast launch); resets the counter. */
long long fa2_profile_kernel_launches(void);

/* Release every per-device workspace the library holds. */
int fa2_release_workspaces(void);

#ifdef __cplusplus
}
#endif
#endif /* FA2_B200_H */
