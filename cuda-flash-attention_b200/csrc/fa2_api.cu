// fa2_api.cu -- the C ABI of libfa2_b200.so (include/fa2_b200.h): argument checking, the
// per-device workspace, TMA descriptor construction, the device-pointer entry points and the
// host-pointer entry points with the batch*head partitioner across the GPUs of one box.
//
// Replaces, for the fa2 path only: host_flash_attention2_{forward,backward}{,_fp16}
// (kernels/kernel_fa2_optimized.cu:351-423, kernels/f-attn2-backward.cu:384-485 and twins)
// and the CuPy kernel wrappers.  No exit(): errors are returned.
#include "../../include/fa2_b200.h"
#include "fa2_common.h"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fa2 {
namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define FA2_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(FA2_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
// ------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Descriptors are a pure function of (base, geometry, type): a small per-thread cache saves the 4 + 7 driver calls
// per forward + backward call when the same tensors come back (training loops, the harness's timed repeats).
struct TmapKey {
    const void* base; int BH, S, pitch, D, kind, box_rows;      // kind: 0 = fp16 copy, 1 = bf16 copy, 2 = fp32 tensor
    bool operator==(const TmapKey& o) const {
        return base == o.base && BH == o.BH && S == o.S && pitch == o.pitch && D == o.D && kind == o.kind && box_rows == o.box_rows;
    }
};
struct TmapCache {
    static constexpr int N = 32;
    TmapKey key[N];
    CUtensorMap map[N];
    int used = 0, next = 0;
    const CUtensorMap* find(const TmapKey& k) const {
        for (int i = 0; i < used; ++i) if (key[i] == k) return &map[i];
        return nullptr;
    }
    void put(const TmapKey& k, const CUtensorMap& m) {
        const int i = used < N ? used++ : (next = (next + 1) % N);
        key[i] = k; map[i] = m;
    }
};
thread_local TmapCache g_tmaps;

// 16-bit [BH][S][DP] tensor, box {64 cols, box_rows, 1 slab}, 128-byte swizzle, zero OOB fill.  pitch_rows > 0: the S
// rows are a row range of slabs that are pitch_rows apart in memory (base points at the first row of the range).
int make_tmap_16(CUtensorMap* tm, void* base, int BH, int S, int DP, int box_rows, int bf16, int pitch_rows = 0) {
    if (pitch_rows <= 0) pitch_rows = S;
    const TmapKey k{base, BH, S, pitch_rows, DP, bf16 ? 1 : 0, box_rows};
    if (const CUtensorMap* hit = g_tmaps.find(k)) { *tm = *hit; return FA2_OK; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(FA2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(DP), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(BH)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(DP) * 2, static_cast<cuuint64_t>(pitch_rows) * DP * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FA2_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    g_tmaps.put(k, *tm);
    return FA2_OK;
}

// fp32 [BH][S][D] tensor addressed by the dQ reduce-add: box {32 cols, 128 rows, 1 slab}, 128-byte swizzle.
int make_tmap_f32(CUtensorMap* tm, void* base, int BH, int S, int D, int pitch_rows = 0, int box_rows = 128) {
    if (pitch_rows <= 0) pitch_rows = S;
    const TmapKey k{base, BH, S, pitch_rows, D, 2, box_rows};
    if (const CUtensorMap* hit = g_tmaps.find(k)) { *tm = *hit; return FA2_OK; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(FA2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(BH)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(D) * 4, static_cast<cuuint64_t>(pitch_rows) * D * 4};
    cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FA2_ERR_CUDA, "cuTensorMapEncodeTiled(fp32) failed with CUresult %d", (int)r);
    g_tmaps.put(k, *tm);
    return FA2_OK;
}

// ------------------------------------------------------------------------------------------
// per-device arenas (grown on demand, kept until fa2_release_workspaces)
// ------------------------------------------------------------------------------------------
struct Arena {
    void* ptr = nullptr;
    size_t bytes = 0;
};
constexpr int kMaxDevices = 64;
std::mutex g_mu;
Arena g_work[kMaxDevices];   // 16-bit operand copies, delta, lse_log2
Arena g_io[kMaxDevices];     // fp32 device mirrors used by the host-pointer entry points

// zero_head > 0: the first zero_head bytes are state that must start at zero (the work arena's RangeBlock).
int arena_reserve(Arena* arenas, int dev, size_t bytes, void** out, size_t zero_head = 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    Arena& a = arenas[dev];
    if (a.bytes < bytes) {
        if (a.ptr) {
            FA2_CUDA(cudaDeviceSynchronize());
            FA2_CUDA(cudaFree(a.ptr));
            a.ptr = nullptr;
            a.bytes = 0;
        }
        FA2_CUDA(cudaMalloc(&a.ptr, bytes));
        a.bytes = bytes;
        if (zero_head) FA2_CUDA(cudaMemset(a.ptr, 0, zero_head));     // (synchronous: ordered before any later launch)
    }
    *out = a.ptr;
    return FA2_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WorkLayout {
    size_t off_q, off_k, off_v, off_do, off_delta, off_lse2, total;
};
WorkLayout work_layout(size_t rows, int DP, bool backward) {
    WorkLayout w{};
    const size_t t16 = align_up(rows * DP * 2, 1024);
    const size_t vec = align_up(rows * 4, 1024);
    size_t off = kRangeBytes;               // the RangeBlock sits at offset 0 whatever the shape
    w.off_q = off; off += t16;
    w.off_k = off; off += t16;
    w.off_v = off; off += t16;
    w.off_do = off; if (backward) off += t16;
    w.off_delta = off; if (backward) off += vec;
    w.off_lse2 = off; if (backward) off += vec;
    w.total = off;
    return w;
}

int check_shape(int B, int H, int S, int D) {
    if (B <= 0 || H <= 0 || S <= 0 || D <= 0)
        return fail(FA2_ERR_INVALID_ARGUMENT, "dimensions must be positive (B=%d H=%d S=%d D=%d)", B, H, S, D);
    if (D != 32 && D != 64 && D != 128)
        return fail(FA2_ERR_INVALID_ARGUMENT, "Unsupported head dimension %d (supported: 32, 64, 128)", D);
    if (static_cast<long long>(B) * H > 0x7fffffffLL / 64)
        return fail(FA2_ERR_INVALID_ARGUMENT, "B*H too large");
    return FA2_OK;
}
// float4 loads/stores, TMA and the reduce-add all need 16-byte aligned tensors (any cudaMalloc / torch / CuPy
// allocation is; an odd sub-view is not).
bool aligned16(std::initializer_list<const void*> ptrs) {
    for (const void* q : ptrs)
        if (reinterpret_cast<uintptr_t>(q) & 15u) return false;
    return true;
}

int check_precision(int precision) {
    if (precision != FA2_PRECISION_FP16 && precision != FA2_PRECISION_FP32 && precision != FA2_PRECISION_BF16)
        return fail(FA2_ERR_INVALID_ARGUMENT, "unknown precision %d", precision);
    return FA2_OK;
}

struct Prepared {
    int dev = 0;
    int BH = 0, S = 0, D = 0, DP = 0, bf16 = 0;
    size_t rows = 0;
    uint8_t* work = nullptr;
    WorkLayout wl{};
    float scale = 0.f, scale_log2 = 0.f;
    RangeBlock* range() const { return reinterpret_cast<RangeBlock*>(work); }
};

int prepare(Prepared* pr, int B, int H, int S, int D, int precision, bool backward) {
    int rc = check_shape(B, H, S, D);
    if (rc) return rc;
    rc = check_precision(precision);
    if (rc) return rc;
    FA2_CUDA(cudaGetDevice(&pr->dev));
    if (pr->dev >= kMaxDevices) return fail(FA2_ERR_UNSUPPORTED, "device ordinal %d too large", pr->dev);
    pr->BH = B * H; pr->S = S; pr->D = D; pr->DP = padded_head_dim(D);
    pr->bf16 = (precision == FA2_PRECISION_BF16) ? 1 : 0;
    pr->rows = static_cast<size_t>(pr->BH) * S;
    pr->wl = work_layout(pr->rows, pr->DP, backward);
    void* w = nullptr;
    rc = arena_reserve(g_work, pr->dev, pr->wl.total, &w, kRangeBytes);
    if (rc) return rc;
    pr->work = static_cast<uint8_t*>(w);
    pr->scale = 1.0f / sqrtf(static_cast<float>(D));
    pr->scale_log2 = pr->scale * 1.4426950408889634f;
    return FA2_OK;
}

// optional per-kernel timing (fa2_profile_enable / fa2_profile_read)
std::atomic<bool> g_profile{false};
struct ProfSpan { int kind; cudaEvent_t a, b; };
std::atomic<long long> g_kernel_launches{0};   // kernels launched inside profiled spans since the last read
std::mutex g_prof_mu;                      // guards g_spans (the host entry points launch from one thread per device)
std::vector<ProfSpan> g_spans;
struct ProfScope {
    int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool on;
    // n_kernels = kernels launched inside the span (the Q/K/V cast of the large path is two: cast + re-cast)
    ProfScope(int k, cudaStream_t s, int n_kernels = 1) : kind(k), st(s), on(g_profile.load(std::memory_order_relaxed)) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); g_kernel_launches += n_kernels; }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(g_prof_mu);
            g_spans.push_back({kind, a, b});
        }
    }
};

unsigned long long* g_timeline = nullptr;   // set by fa2_debug_set_timeline (debug builds)

int run_cast(const Prepared& pr, const float* Q, const float* K, const float* V, cudaStream_t st) {
    ProfScope prof(0, st, 2);
    FA2_CUDA(launch_cast_qkv(Q, K, V, pr.work + pr.wl.off_q, pr.work + pr.wl.off_k, pr.work + pr.wl.off_v, pr.rows,
                             pr.D, pr.DP, pr.bf16, pr.range(), pr.scale, pr.scale_log2, st));
    // the cast's last block has decided the scales; re-cast only what does not fit fp16 as it is
    FA2_CUDA(launch_range_fix_qkv(Q, K, V, pr.work + pr.wl.off_q, pr.work + pr.wl.off_k, pr.work + pr.wl.off_v, pr.rows,
                                  pr.D, pr.DP, pr.bf16, pr.range(), st));
    return FA2_OK;
}

// Problems of up to ~1.2 M elements per tensor: one cooperative launch does amax + scale decision + cast (+ dO cast
// and dQ zero-fill when dO is given) -- the always-launched re-cast kernels would cost more than the cast itself there.
bool is_small(const Prepared& pr) { return cast_small_fits(pr.rows, pr.DP, sm_count_current()); }

int run_cast_small(const Prepared& pr, const float* Q, const float* K, const float* V, const float* dO, float* dQ,
                   cudaStream_t st) {
    ProfScope prof(0, st);
    FA2_CUDA(launch_cast_small(Q, K, V, dO, pr.work + pr.wl.off_q, pr.work + pr.wl.off_k, pr.work + pr.wl.off_v,
                               dO ? pr.work + pr.wl.off_do : nullptr, dQ, pr.rows, pr.D, pr.DP, pr.bf16, pr.range(),
                               pr.scale, pr.scale_log2, sm_count_current(), st));
    return FA2_OK;
}

// after the dO cast (pre-pass or the forward's donor warps), before the backward kernel
int run_fix_do(const Prepared& pr, const float* dO, cudaStream_t st) {
    ProfScope prof(2, st);
    FA2_CUDA(launch_range_fix_do(dO, pr.work + pr.wl.off_do, pr.rows, pr.D, pr.DP, pr.bf16, pr.range(), st));
    return FA2_OK;
}

// Tuning knob FA2_FUSE (default 3): bit 0 = the forward's donor warps cast dO and zero-fill dQ, bit 1 = the forward's
// epilogue forms D_i and LSE*log2(e); whatever is switched off is done by the stand-alone backward pre-pass instead.
int fuse_mask() {
    static const int mask = [] { const char* ev = getenv("FA2_FUSE"); return ev ? (atoi(ev) & 3) : 3; }();
    return mask;
}

// dO / dQ non-null = fused forward+backward: the forward kernel also writes the 16-bit dO copy, D_i, LSE*log2(e)
// and zero-fills dQ, so no separate backward pre-pass is launched.
// Row range of every slab a launch covers (sequence-split path); {0, S} = everything.
struct RowRange { int r0, r1; };

int run_fwd_main(const Prepared& pr, float* O, float* LSE, cudaStream_t st, const float* dO = nullptr,
                 float* dQ = nullptr, int mask = 0, const RowRange* qr = nullptr) {
    const int r0 = qr ? qr->r0 : 0, Sq = qr ? qr->r1 - qr->r0 : pr.S;
    const size_t e0 = static_cast<size_t>(r0) * pr.D, h0 = static_cast<size_t>(r0) * pr.DP * 2;
    FwdParams p{};
    int rc;
    if ((rc = make_tmap_16(&p.tm_q, pr.work + pr.wl.off_q + h0, pr.BH, Sq, pr.DP, 128, pr.bf16, pr.S))) return rc;
    if ((rc = make_tmap_16(&p.tm_k, pr.work + pr.wl.off_k, pr.BH, pr.S, pr.DP, 128, pr.bf16))) return rc;
    if ((rc = make_tmap_16(&p.tm_v, pr.work + pr.wl.off_v, pr.BH, pr.S, pr.DP, 128, pr.bf16))) return rc;
    if ((rc = make_tmap_f32(&p.tm_o, O + e0, pr.BH, Sq, pr.D, pr.S))) return rc;
    p.O = O + e0; p.LSE = LSE + r0; p.BH = pr.BH; p.S_q = Sq; p.S_kv = pr.S; p.q_pitch = pr.S; p.D = pr.D;
    p.donor_rows = static_cast<long long>(pr.rows);
    p.scale = pr.scale; p.scale_log2 = pr.scale_log2; p.bf16 = pr.bf16;
    p.range = pr.range()->sc;
    p.timeline = g_timeline;
    if (dO && dQ && mask) {
        p.dO = dO + e0;
        if (mask & 1) { p.dOh = pr.work + pr.wl.off_do; p.dQ_zero = dQ; p.rb = pr.range(); }
        if (mask & 2) {
            p.delta = reinterpret_cast<float*>(pr.work + pr.wl.off_delta) + r0;
            p.lse_log2 = reinterpret_cast<float*>(pr.work + pr.wl.off_lse2) + r0;
        }
    }
    ProfScope prof(1, st);
    FA2_CUDA(launch_fwd(p, st));
    return FA2_OK;
}

int run_bwd_prepass(const Prepared& pr, const float* O, const float* dO, const float* LSE, float* dQ, int parts,
                    cudaStream_t st, const RowRange* qr = nullptr) {
    const int r0 = qr ? qr->r0 : 0, n = qr ? qr->r1 - qr->r0 : pr.S;
    const size_t e0 = static_cast<size_t>(r0) * pr.D;
    float* delta = reinterpret_cast<float*>(pr.work + pr.wl.off_delta) + r0;
    float* lse2 = reinterpret_cast<float*>(pr.work + pr.wl.off_lse2) + r0;
    ProfScope prof(2, st);
    FA2_CUDA(launch_bwd_prepass(O + e0, dO + e0, LSE + r0, pr.work + pr.wl.off_do + static_cast<size_t>(r0) * pr.DP * 2, delta,
                                lse2, dQ + e0, static_cast<size_t>(pr.BH) * n, pr.D, pr.DP, pr.bf16, parts, pr.range(),
                                pr.scale, st, static_cast<unsigned>(n), static_cast<unsigned>(pr.S)));
    return FA2_OK;
}

int run_bwd_main(const Prepared& pr, float* dQ, float* dK, float* dV, cudaStream_t st, const RowRange* kr = nullptr) {
    const int r0 = kr ? kr->r0 : 0, Skv = kr ? kr->r1 - kr->r0 : pr.S;
    const size_t e0 = static_cast<size_t>(r0) * pr.D, h0 = static_cast<size_t>(r0) * pr.DP * 2;
    float* delta = reinterpret_cast<float*>(pr.work + pr.wl.off_delta);
    float* lse2 = reinterpret_cast<float*>(pr.work + pr.wl.off_lse2);
    BwdParams p{};
    int rc;
    if ((rc = make_tmap_16(&p.tm_q, pr.work + pr.wl.off_q, pr.BH, pr.S, pr.DP, 128, pr.bf16))) return rc;
    if ((rc = make_tmap_16(&p.tm_k, pr.work + pr.wl.off_k + h0, pr.BH, Skv, pr.DP, 128, pr.bf16, pr.S))) return rc;
    if ((rc = make_tmap_16(&p.tm_v, pr.work + pr.wl.off_v + h0, pr.BH, Skv, pr.DP, 128, pr.bf16, pr.S))) return rc;
    if ((rc = make_tmap_16(&p.tm_do, pr.work + pr.wl.off_do, pr.BH, pr.S, pr.DP, 128, pr.bf16))) return rc;
    if ((rc = make_tmap_f32(&p.tm_dq, dQ, pr.BH, pr.S, pr.D))) return rc;
    if (bwd_uses_pair(pr.D)) {
        if ((rc = make_tmap_16(&p.tm_q64, pr.work + pr.wl.off_q, pr.BH, pr.S, pr.DP, 64, pr.bf16))) return rc;
        if ((rc = make_tmap_16(&p.tm_do64, pr.work + pr.wl.off_do, pr.BH, pr.S, pr.DP, 64, pr.bf16))) return rc;
        if ((rc = make_tmap_f32(&p.tm_dq64, dQ, pr.BH, pr.S, pr.D, 0, 64))) return rc;
    }
    if ((rc = make_tmap_f32(&p.tm_dk, dK + e0, pr.BH, Skv, pr.D, pr.S))) return rc;
    if ((rc = make_tmap_f32(&p.tm_dv, dV + e0, pr.BH, Skv, pr.D, pr.S))) return rc;
    p.lse_log2 = lse2; p.delta = delta; p.dQ = dQ; p.dK = dK + e0; p.dV = dV + e0;
    p.BH = pr.BH; p.S_q = pr.S; p.S_kv = Skv; p.D = pr.D; p.scale = pr.scale; p.scale_log2 = pr.scale_log2; p.bf16 = pr.bf16;
    p.range = pr.range()->sc;
    p.timeline = getenv("FA2_TL_FWD_ONLY") ? nullptr : g_timeline;     // (timeline builds: both kernels share the buffer)
    ProfScope prof(3, st);
    FA2_CUDA(launch_bwd(p, st));
    return FA2_OK;
}

// ------------------------------------------------------------------------------------------
// host-pointer path: one worker thread per device, each on its own contiguous slab range
// ------------------------------------------------------------------------------------------
struct HostJob {
    const float *Q, *K, *V, *O_in, *dO, *LSE_in;
    float *O, *LSE, *dQ, *dK, *dV;
    int B, H, S, D, precision;
    int mode;   // FA2_MODE_*
};

// One device's share [bh0, bh0+count) of the slabs, processed in chunks through kSets device buffer sets and
// three streams so that the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1
// overlap (PCIe is full duplex; the reference does malloc -> H2D -> kernel -> D2H -> free serially,
// kernels/kernel_fa2_optimized.cu:371-421).
// Streams and events of the host pipeline, kept per device between calls (creating ~60 of them costs ~0.7 ms per
// call).  PipeSet::mu is THE per-device lock of the host-pointer entry points: it is held for a whole call and covers
// the streams, the events and both arenas (g_io / g_work) of the device, so two concurrent fa2_host_* calls that
// land on the same device run one after the other.  Lock order everywhere: PipeSet::mu, then g_mu.
constexpr int kSets = 3;     // device buffer sets of the host pipeline (H2D of c+1, kernels of c, D2H of c-1 and its tail)
struct PipeSet {
    std::mutex mu;
    bool ready = false;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSets] = {}, ev_comp[kSets] = {}, ev_out[kSets] = {};
    std::vector<cudaEvent_t> k0, k1;          // timed events around each chunk's kernels

    cudaError_t init() {
        cudaError_t e;
        if (!ready) {
            if ((e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
            for (int i = 0; i < kSets; ++i) {
                if ((e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
            ready = true;
        }
        return cudaSuccess;
    }
    cudaError_t grow(size_t n) {
        cudaError_t e;
        while (k0.size() < n) {
            cudaEvent_t a, b;
            if ((e = cudaEventCreate(&a)) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&b)) != cudaSuccess) return e;
            k0.push_back(a);
            k1.push_back(b);
        }
        return cudaSuccess;
    }
    void destroy() {
        if (!ready) return;
        cudaStreamDestroy(s_in); cudaStreamDestroy(s_comp); cudaStreamDestroy(s_out);
        for (int i = 0; i < kSets; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_comp[i]); cudaEventDestroy(ev_out[i]); }
        for (cudaEvent_t e : k0) cudaEventDestroy(e);
        for (cudaEvent_t e : k1) cudaEventDestroy(e);
        k0.clear(); k1.clear();
        ready = false;
    }
};
PipeSet g_pipes[kMaxDevices];

// Chunk sizes (in slabs) of one device's share.  The host path is PCIe-bound (pinned Gen5 x16: ~47 GB/s each way in
// duplex; the kernels of a chunk take a fraction of its transfer time), so the job takes (bytes of the busier
// direction) / (duplex rate) plus whatever time only one direction is busy: the H2D of the first chunk and, at the end,
// the backlog of the D2H stream, which runs one chunk (+ its kernels) behind the H2D stream.  Small chunks shorten
// head and tail but cost launches and synchronisation (3-slab chunks at config C: 35 GB/s per direction, 59 ms);
// large ones leave a long tail (16 slabs: 2.2 ms of one-directional D2H).  Measured at config C (tools/e2e_trace.py
// with FA2_CHUNK_BIG / FA2_CHUNK_MIN, profiles/r02/experiments.md): ~20 MiB per tensor copy (10 slabs) with a
// geometric ramp from a ~0.4 ms first chunk is the best and the most repeatable (47.1-47.8 ms on a 47 GB/s box,
// floor 45.7 ms).  Every chunk is grown until the persistent kernels -- a chunk costs ceil(items / SMs) rounds of one
// work item each -- keep up with the copies.
std::vector<int> plan_chunks(int count, int S, int D, bool fwd, bool bwd, int n_sm = 148) {
    std::vector<int> sizes;
    if (count <= 0) return sizes;
    const double slab_bytes = static_cast<double>(S) * D * 4.0;
    const double in_b = slab_bytes * (3 + (bwd ? 1 : 0) + (!fwd ? 1 : 0)), out_b = slab_bytes * ((fwd ? 1 : 0) + (bwd ? 3 : 0));
    const double t_x = (in_b > out_b ? in_b : out_b) / 45e9;                      // seconds of PCIe per slab
    const int items_f = (S + 255) / 256, items_b = (S + 127) / 128, n_steps = (S + 127) / 128;
    const double step_f = D > 64 ? 1.65e-6 : 0.95e-6, step_b = D > 64 ? 2.15e-6 : 1.25e-6;   // measured per 128x128 tile step
    auto kernels = [&](int c) {
        double t = 10e-6 + c * in_b * 1.5 / 6.5e12;                               // launches + the fp32 -> 16-bit pre-passes
        if (fwd) t += static_cast<double>((static_cast<long long>(c) * items_f + n_sm - 1) / n_sm) * n_steps * step_f;
        if (bwd) t += static_cast<double>((static_cast<long long>(c) * items_b + n_sm - 1) / n_sm) * n_steps * step_b;
        return t;
    };
    int cmin = static_cast<int>(0.4e-3 / t_x + 0.999);
    if (cmin < 1) cmin = 1;
    if (cmin > count) cmin = count;
    while (cmin < count && kernels(cmin) > 0.8 * t_x * cmin) ++cmin;
    // large chunks: ~20 MiB per tensor copy, but never more than 1/8 of the job (the pipeline needs chunks)
    int cbig = static_cast<int>(20971520.0 / slab_bytes + 0.999);
    if (cbig > count / 8) cbig = count / 8;
    if (const char* e = getenv("FA2_CHUNK_BIG")) cbig = atoi(e);      // tuning knob (tools/e2e_trace.py)
    if (const char* e = getenv("FA2_CHUNK_MIN")) cmin = atoi(e);
    if (cbig < cmin) cbig = cmin;
    std::vector<int> ramp;
    int ramp_sum = 0;
    for (long long c = cmin; c < cbig; c *= 2) { ramp.push_back(static_cast<int>(c)); ramp_sum += static_cast<int>(c); }
    if (ramp.empty() || 2 * ramp_sum + cbig > count) {
        // short job: equal chunks of the smallest size
        int left = count;
        while (left > 0) { const int n = left < cmin ? left : cmin; sizes.push_back(n); left -= n; }
        return sizes;
    }
    const int mid = count - 2 * ramp_sum, n_mid = (mid + cbig - 1) / cbig;
    sizes = ramp;
    for (int i = 0; i < n_mid; ++i) sizes.push_back(mid / n_mid + (i < mid % n_mid ? 1 : 0));
    for (size_t i = ramp.size(); i-- > 0;) sizes.push_back(ramp[i]);
    return sizes;
}

int host_worker(const HostJob& job, int dev, int bh0, int count, float* ms_out, std::string* err_out) {
    if (dev < 0 || dev >= kMaxDevices) { *err_out = "device ordinal out of range"; return FA2_ERR_UNSUPPORTED; }
    // Serialise host calls per device: the arenas and the cached streams belong to one call at a time.
    std::lock_guard<std::mutex> device_lock(g_pipes[dev].mu);
    PipeSet& ps = g_pipes[dev];
    // Whatever way run() is left, nothing may still be in flight on the buffers the next call will reuse.
    struct Drain {
        PipeSet& s;
        ~Drain() {
            if (!s.ready) return;
            cudaStreamSynchronize(s.s_in); cudaStreamSynchronize(s.s_comp); cudaStreamSynchronize(s.s_out);
        }
    } drain{ps};
    auto run = [&]() -> int {
        FA2_CUDA(cudaSetDevice(dev));
        const size_t slab = static_cast<size_t>(job.S) * job.D;           // floats per (b,h)
        const bool fwd = job.mode != FA2_MODE_BACKWARD, bwd = job.mode != FA2_MODE_FORWARD;
        std::vector<int> sizes = plan_chunks(count, job.S, job.D, fwd, bwd);
        int chunk_bh = 0;
        for (int c : sizes) chunk_bh = c > chunk_bh ? c : chunk_bh;
        const int n_chunks = static_cast<int>(sizes.size());
        const size_t tb = align_up(slab * chunk_bh * 4, 1024), lb = align_up(static_cast<size_t>(job.S) * chunk_bh * 4, 1024);
        const size_t set_bytes = 4 * tb + lb + (bwd ? 4 * tb : 0);       // Q K V O LSE [dO dQ dK dV]
        void* base = nullptr;
        int rc = arena_reserve(g_io, dev, kSets * set_bytes, &base);
        if (rc) return rc;
        {   // keep one-time costs (module load, workspace growth) out of the timed region
            Prepared warm;
            if ((rc = prepare(&warm, 1, chunk_bh, job.S, job.D, job.precision, bwd))) return rc;
            FA2_CUDA(warm_fwd());
            FA2_CUDA(warm_bwd());
        }
        FA2_CUDA(ps.init());
        FA2_CUDA(ps.grow(static_cast<size_t>(n_chunks)));
        cudaStream_t s_in = ps.s_in, s_comp = ps.s_comp, s_out = ps.s_out;
        cudaEvent_t* ev_in = ps.ev_in;
        cudaEvent_t* ev_comp = ps.ev_comp;
        cudaEvent_t* ev_out = ps.ev_out;
        std::vector<cudaEvent_t>& k0 = ps.k0;
        std::vector<cudaEvent_t>& k1 = ps.k1;
        // FA2_HOST_TRACE=1: print when each chunk's copies and kernels ran (pipeline tuning aid)
        static const bool trace = getenv("FA2_HOST_TRACE") != nullptr;
        const auto wall0 = std::chrono::steady_clock::now();
        std::vector<cudaEvent_t> tr;
        if (trace) {
            tr.resize(3 * n_chunks + 1);
            for (auto& e : tr) FA2_CUDA(cudaEventCreate(&e));
            FA2_CUDA(cudaEventRecord(tr[3 * n_chunks], s_in));
        }
        int cb0 = bh0;
        for (int c = 0; c < n_chunks; cb0 += sizes[c], ++c) {
            const int set = c % kSets;
            const int cnt = sizes[c];
            const size_t n = slab * cnt, nl = static_cast<size_t>(job.S) * cnt;
            const size_t off = static_cast<size_t>(cb0) * slab, offl = static_cast<size_t>(cb0) * job.S;
            uint8_t* b8 = static_cast<uint8_t*>(base) + set * set_bytes;
            float* dQin = reinterpret_cast<float*>(b8);
            float* dKin = reinterpret_cast<float*>(b8 + tb);
            float* dVin = reinterpret_cast<float*>(b8 + 2 * tb);
            float* dO_ = reinterpret_cast<float*>(b8 + 3 * tb);
            float* dL = reinterpret_cast<float*>(b8 + 4 * tb);
            float* ddO = bwd ? reinterpret_cast<float*>(b8 + 4 * tb + lb) : nullptr;
            float* dQ_ = bwd ? reinterpret_cast<float*>(b8 + 5 * tb + lb) : nullptr;
            float* dK_ = bwd ? reinterpret_cast<float*>(b8 + 6 * tb + lb) : nullptr;
            float* dV_ = bwd ? reinterpret_cast<float*>(b8 + 7 * tb + lb) : nullptr;
            // inputs of this buffer set were consumed by the kernels of chunk c-kSets
            if (c >= kSets) FA2_CUDA(cudaStreamWaitEvent(s_in, ev_comp[set], 0));
            FA2_CUDA(cudaMemcpyAsync(dQin, job.Q + off, n * 4, cudaMemcpyHostToDevice, s_in));
            FA2_CUDA(cudaMemcpyAsync(dKin, job.K + off, n * 4, cudaMemcpyHostToDevice, s_in));
            FA2_CUDA(cudaMemcpyAsync(dVin, job.V + off, n * 4, cudaMemcpyHostToDevice, s_in));
            if (job.mode == FA2_MODE_BACKWARD) {
                FA2_CUDA(cudaMemcpyAsync(dO_, job.O_in + off, n * 4, cudaMemcpyHostToDevice, s_in));
                FA2_CUDA(cudaMemcpyAsync(dL, job.LSE_in + offl, nl * 4, cudaMemcpyHostToDevice, s_in));
            }
            if (bwd) FA2_CUDA(cudaMemcpyAsync(ddO, job.dO + off, n * 4, cudaMemcpyHostToDevice, s_in));
            FA2_CUDA(cudaEventRecord(ev_in[set], s_in));
            if (trace) FA2_CUDA(cudaEventRecord(tr[3 * c], s_in));
            // kernels: need the inputs, and the outputs of chunk c-kSets must have left this buffer set
            FA2_CUDA(cudaStreamWaitEvent(s_comp, ev_in[set], 0));
            if (c >= kSets) FA2_CUDA(cudaStreamWaitEvent(s_comp, ev_out[set], 0));
            FA2_CUDA(cudaEventRecord(k0[c], s_comp));
            if (job.mode == FA2_MODE_FORWARD)
                rc = fa2_forward(dQin, dKin, dVin, dO_, dL, 1, cnt, job.S, job.D, job.precision, s_comp);
            else if (job.mode == FA2_MODE_BACKWARD)
                rc = fa2_backward(dQin, dKin, dVin, dO_, ddO, dL, dQ_, dK_, dV_, 1, cnt, job.S, job.D, job.precision, s_comp);
            else
                rc = fa2_forward_backward(dQin, dKin, dVin, ddO, dO_, dL, dQ_, dK_, dV_, 1, cnt, job.S, job.D,
                                          job.precision, s_comp);
            if (rc) return rc;
            FA2_CUDA(cudaEventRecord(k1[c], s_comp));
            FA2_CUDA(cudaEventRecord(ev_comp[set], s_comp));
            FA2_CUDA(cudaStreamWaitEvent(s_out, ev_comp[set], 0));
            if (fwd) {
                FA2_CUDA(cudaMemcpyAsync(job.O + off, dO_, n * 4, cudaMemcpyDeviceToHost, s_out));
                FA2_CUDA(cudaMemcpyAsync(job.LSE + offl, dL, nl * 4, cudaMemcpyDeviceToHost, s_out));
            }
            if (bwd) {
                FA2_CUDA(cudaMemcpyAsync(job.dQ + off, dQ_, n * 4, cudaMemcpyDeviceToHost, s_out));
                FA2_CUDA(cudaMemcpyAsync(job.dK + off, dK_, n * 4, cudaMemcpyDeviceToHost, s_out));
                FA2_CUDA(cudaMemcpyAsync(job.dV + off, dV_, n * 4, cudaMemcpyDeviceToHost, s_out));
            }
            FA2_CUDA(cudaEventRecord(ev_out[set], s_out));
            if (trace) {
                FA2_CUDA(cudaEventRecord(tr[3 * c + 1], s_comp));
                FA2_CUDA(cudaEventRecord(tr[3 * c + 2], s_out));
            }
        }
        const auto wall1 = std::chrono::steady_clock::now();
        FA2_CUDA(cudaStreamSynchronize(s_in));
        FA2_CUDA(cudaStreamSynchronize(s_comp));
        FA2_CUDA(cudaStreamSynchronize(s_out));
        if (trace) {
            const auto wall2 = std::chrono::steady_clock::now();
            auto ms_of = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[fa2 host trace] dev %d: %d chunks, enqueue %.2f ms, drained at %.2f ms (wall, after setup)\n",
                    dev, n_chunks, ms_of(wall0, wall1), ms_of(wall0, wall2));
            for (int c = 0; c < n_chunks; ++c) {
                float a = 0, b = 0, d = 0, k = 0;
                cudaEventElapsedTime(&a, tr[3 * n_chunks], tr[3 * c]);
                cudaEventElapsedTime(&b, tr[3 * n_chunks], tr[3 * c + 1]);
                cudaEventElapsedTime(&d, tr[3 * n_chunks], tr[3 * c + 2]);
                cudaEventElapsedTime(&k, k0[c], k1[c]);
                fprintf(stderr, "  chunk %2d (%3d slabs): H2D done %7.2f  kernels done %7.2f (%.2f ms)  D2H done %7.2f\n", c,
                        sizes[c], a, b, k, d);
            }
            for (auto& e : tr) cudaEventDestroy(e);
        }
        float total_ms = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            float ms = 0.f;
            FA2_CUDA(cudaEventElapsedTime(&ms, k0[c], k1[c]));
            total_ms += ms;
        }
        *ms_out = total_ms;
        return FA2_OK;
    };
    int rc = run();
    if (rc) *err_out = g_last_error;
    return rc;
}

// ------------------------------------------------------------------------------------------
// sequence-split path: G = G_bh x G_s devices; a group of G_s devices shares a slab range and splits its rows
// ------------------------------------------------------------------------------------------
// Used when there are fewer (b,h) slabs than devices (or an uneven handful): within a group every device holds the
// group's slabs in full, runs the FORWARD on its own range of query rows (K / V replicated, no exchange) and the
// BACKWARD on the same range of key/value rows against all query rows.  That leaves one partial dQ per device:
// the only collective of the whole design, a reduce-scatter over the group done by P2P loads across NVLink
// (dq_peer_reduce_kernel), after an all-gather of the 2 x S floats per slab of D_i and LSE * log2(e) the forward
// ranges produced (peer copies).  Each device returns O / LSE / dQ / dK / dV rows of its own range only.
// Inputs are replicated over PCIe (each device of a group uploads the group's slabs), which costs end-to-end time;
// what the split buys is kernel time when B*H cannot occupy the devices.
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned gen = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

// Row range of part `part` of `parts` over S rows, boundaries on multiples of 256 rows (one forward work item).
void seq_range(int S, int parts, int part, int* r0, int* r1) {
    auto cut = [&](int i) {
        if (i <= 0) return 0;
        if (i >= parts) return S;
        long long x = (static_cast<long long>(S) * i / parts + 128) / 256 * 256;
        return static_cast<int>(x > S ? S : x);
    };
    *r0 = cut(part);
    *r1 = cut(part + 1);
}

// (G_bh, G_s) for BH slabs on G devices.  With at least one slab per device the plain slab split is used: the host
// path is PCIe-bound and a sequence split replicates a group's inputs on each of its devices.  With fewer slabs than
// devices the rows are split only if that shortens the kernels: the persistent kernels need ceil(items / SMs) rounds
// of one work item each (an item walks all S / 128 tiles of the other operand), so splitting rows pays once a
// device's share exceeds one round -- a single slab has S / 256 forward items, i.e. beyond S = 32k.
// FA2_SEQ_SPLIT=k forces G_s = k (k must divide G; used by the tests and tools/seq_split_bench.py).
void choose_split(int BH, int S, int G, int* g_bh, int* g_s) {
    int forced = 0;
    if (const char* e = getenv("FA2_SEQ_SPLIT")) forced = atoi(e);
    *g_bh = G < BH ? G : BH;
    *g_s = 1;
    if (forced <= 0 && BH >= G) return;
    const int n_sm = 148;
    const double n_steps = (S + 127) / 128;
    double best = 1e30;
    for (int gs = 1; gs <= 8 && gs <= G; gs *= 2) {
        if (G % gs) continue;
        const int gb = G / gs < BH ? G / gs : BH;
        if (gs > 1 && S / gs < 256) continue;
        if (forced > 0 && gs != forced) continue;
        const long long c = (BH + gb - 1) / gb, rows = (S + gs - 1) / gs;
        const long long items_f = c * ((rows + 255) / 256), items_b = c * ((rows + 127) / 128);
        const double t = ((items_f + n_sm - 1) / n_sm * 1.65 + (items_b + n_sm - 1) / n_sm * 2.15) * n_steps * (gs > 1 ? 1.05 : 1.0);
        if (t < best - 1e-9) { best = t; *g_bh = gb; *g_s = gs; }
    }
}

struct SeqShared {                  // what the devices of one call publish to each other
    std::vector<float*> delta, lse2, dq;       // per device: workspace D_i / LSE*log2e (full rows), partial dQ
    std::vector<cudaEvent_t> ev_side, ev_bwd;  // "my D_i / LSE rows are final", "my backward kernel is done"
    std::vector<int> rc;
    std::vector<std::string> err;
    std::vector<float> ms;
    std::atomic<bool> failed{false};
};

int host_dispatch_seqsplit(const HostJob& job, int G_bh, int G_s, float* kernel_ms) {
    const int G = G_bh * G_s, BH = job.B * job.H, S = job.S, D = job.D;
    const bool fwd = job.mode != FA2_MODE_BACKWARD, bwd = job.mode != FA2_MODE_FORWARD;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    // peer access inside every group (NVLink P2P loads of the dQ reduce and the side-input copies)
    for (int d = 0; d < G; ++d) {
        cudaSetDevice(d);
        for (int q = d / G_s * G_s; q < d / G_s * G_s + G_s; ++q) {
            if (q == d) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, d, q);
            if (!can) { cudaSetDevice(prev_dev); return fail(FA2_ERR_UNSUPPORTED, "sequence split needs peer access between devices %d and %d", d, q); }
            const cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(prev_dev); return fail(FA2_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", d, q, cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    SeqShared sh;
    sh.delta.assign(G, nullptr); sh.lse2.assign(G, nullptr); sh.dq.assign(G, nullptr);
    sh.ev_side.assign(G, nullptr); sh.ev_bwd.assign(G, nullptr);
    sh.rc.assign(G, 0); sh.err.assign(G, ""); sh.ms.assign(G, 0.f);
    HostBarrier bar(G);

    auto worker = [&](int d) {
        const int gb = d / G_s, gs = d % G_s;
        int bh0 = 0, cnt = 0, r0 = 0, r1 = 0;
        fa2_partition(BH, G_bh, gb, &bh0, &cnt);
        seq_range(S, G_s, gs, &r0, &r1);
        const RowRange rr{r0, r1};
        const size_t slab = static_cast<size_t>(S) * D, n = slab * cnt, nl = static_cast<size_t>(S) * cnt;
        const size_t off = static_cast<size_t>(bh0) * slab, offl = static_cast<size_t>(bh0) * S;
        std::unique_lock<std::mutex> device_lock(g_pipes[d].mu);        // one host call per device at a time
        cudaStream_t st = nullptr;
        cudaEvent_t k0 = nullptr, k1 = nullptr;
        float *dQ_ = nullptr, *dK_ = nullptr, *dV_ = nullptr, *dO_ = nullptr, *dQin = nullptr, *dKin = nullptr, *dVin = nullptr,
              *dOut = nullptr, *dL = nullptr;
        Prepared pr;
        // phase A: upload, cast, forward on the own query rows, side inputs of the own rows
        auto phase_a = [&]() -> int {
            FA2_CUDA(cudaSetDevice(d));
            FA2_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            FA2_CUDA(cudaEventCreate(&k0));
            FA2_CUDA(cudaEventCreate(&k1));
            FA2_CUDA(cudaEventCreateWithFlags(&sh.ev_side[d], cudaEventDisableTiming));
            FA2_CUDA(cudaEventCreateWithFlags(&sh.ev_bwd[d], cudaEventDisableTiming));
            const size_t tb = align_up(n * 4, 1024), lb = align_up(nl * 4, 1024);
            void* base = nullptr;
            int rc = arena_reserve(g_io, d, 4 * tb + lb + (bwd ? 4 * tb : 0), &base);
            if (rc) return rc;
            uint8_t* b8 = static_cast<uint8_t*>(base);
            dQin = reinterpret_cast<float*>(b8); dKin = reinterpret_cast<float*>(b8 + tb); dVin = reinterpret_cast<float*>(b8 + 2 * tb);
            dOut = reinterpret_cast<float*>(b8 + 3 * tb); dL = reinterpret_cast<float*>(b8 + 4 * tb);
            if (bwd) {
                dO_ = reinterpret_cast<float*>(b8 + 4 * tb + lb); dQ_ = reinterpret_cast<float*>(b8 + 5 * tb + lb);
                dK_ = reinterpret_cast<float*>(b8 + 6 * tb + lb); dV_ = reinterpret_cast<float*>(b8 + 7 * tb + lb);
            }
            if ((rc = prepare(&pr, 1, cnt, S, D, job.precision, bwd))) return rc;
            FA2_CUDA(warm_fwd());
            FA2_CUDA(warm_bwd());
            sh.delta[d] = reinterpret_cast<float*>(pr.work + pr.wl.off_delta);
            sh.lse2[d] = reinterpret_cast<float*>(pr.work + pr.wl.off_lse2);
            sh.dq[d] = dQ_;
            FA2_CUDA(cudaMemcpyAsync(dQin, job.Q + off, n * 4, cudaMemcpyHostToDevice, st));
            FA2_CUDA(cudaMemcpyAsync(dKin, job.K + off, n * 4, cudaMemcpyHostToDevice, st));
            FA2_CUDA(cudaMemcpyAsync(dVin, job.V + off, n * 4, cudaMemcpyHostToDevice, st));
            if (bwd) FA2_CUDA(cudaMemcpyAsync(dO_, job.dO + off, n * 4, cudaMemcpyHostToDevice, st));
            if (job.mode == FA2_MODE_BACKWARD) {
                FA2_CUDA(cudaMemcpyAsync(dOut, job.O_in + off, n * 4, cudaMemcpyHostToDevice, st));
                FA2_CUDA(cudaMemcpyAsync(dL, job.LSE_in + offl, nl * 4, cudaMemcpyHostToDevice, st));
            }
            FA2_CUDA(cudaEventRecord(k0, st));
            if ((rc = run_cast(pr, dQin, dKin, dVin, st))) return rc;
            if (fwd && (rc = run_fwd_main(pr, dOut, dL, st, nullptr, nullptr, 0, &rr))) return rc;
            if (bwd) {
                // dO cast + dQ zero-fill (and dO's scale decision) over ALL rows; D_i and LSE*log2e over the rows whose
                // O / LSE this device has: its own range after a forward here, everything in backward-only mode
                if ((rc = run_bwd_prepass(pr, dOut, dO_, dL, dQ_, 1, st))) return rc;
                if ((rc = run_bwd_prepass(pr, dOut, dO_, dL, dQ_, 2, st, fwd ? &rr : nullptr))) return rc;
                if ((rc = run_fix_do(pr, dO_, st))) return rc;
            }
            FA2_CUDA(cudaEventRecord(sh.ev_side[d], st));
            return FA2_OK;
        };
        // phase B: gather the peers' D_i / LSE rows, backward on the own key/value rows
        auto phase_b = [&]() -> int {
            if (!bwd) return FA2_OK;
            if (fwd) {
                for (int q = gb * G_s; q < gb * G_s + G_s; ++q) {
                    if (q == d) continue;
                    int q0, q1;
                    seq_range(S, G_s, q % G_s, &q0, &q1);
                    if (q1 <= q0) continue;
                    FA2_CUDA(cudaStreamWaitEvent(st, sh.ev_side[q], 0));
                    const size_t w = static_cast<size_t>(q1 - q0) * 4, pitch = static_cast<size_t>(S) * 4;
                    FA2_CUDA(cudaMemcpy2DAsync(sh.delta[d] + q0, pitch, sh.delta[q] + q0, pitch, w, cnt, cudaMemcpyDefault, st));
                    FA2_CUDA(cudaMemcpy2DAsync(sh.lse2[d] + q0, pitch, sh.lse2[q] + q0, pitch, w, cnt, cudaMemcpyDefault, st));
                }
            }
            int rc = FA2_OK;
            if (r1 > r0 && (rc = run_bwd_main(pr, dQ_, dK_, dV_, st, &rr))) return rc;
            FA2_CUDA(cudaEventRecord(sh.ev_bwd[d], st));
            return FA2_OK;
        };
        // phase C: reduce-scatter of dQ (own rows += the peers' partials, read over NVLink), results to the host
        auto phase_c = [&]() -> int {
            const size_t w = static_cast<size_t>(r1 - r0) * D * 4, pitch = slab * 4;
            if (bwd && r1 > r0) {
                const float* peers[8];
                int np = 0;
                for (int q = gb * G_s; q < gb * G_s + G_s; ++q) {
                    if (q == d) continue;
                    FA2_CUDA(cudaStreamWaitEvent(st, sh.ev_bwd[q], 0));
                    peers[np++] = sh.dq[q] + static_cast<size_t>(r0) * D;
                }
                FA2_CUDA(launch_dq_peer_reduce(dQ_ + static_cast<size_t>(r0) * D, peers, np, static_cast<size_t>(r1 - r0) * D, slab, cnt, st));
            }
            FA2_CUDA(cudaEventRecord(k1, st));
            if (r1 > r0) {
                const size_t e0 = static_cast<size_t>(r0) * D;
                if (fwd) {
                    FA2_CUDA(cudaMemcpy2DAsync(job.O + off + e0, pitch, dOut + e0, pitch, w, cnt, cudaMemcpyDeviceToHost, st));
                    FA2_CUDA(cudaMemcpy2DAsync(job.LSE + offl + r0, static_cast<size_t>(S) * 4, dL + r0, static_cast<size_t>(S) * 4,
                                               static_cast<size_t>(r1 - r0) * 4, cnt, cudaMemcpyDeviceToHost, st));
                }
                if (bwd) {
                    FA2_CUDA(cudaMemcpy2DAsync(job.dQ + off + e0, pitch, dQ_ + e0, pitch, w, cnt, cudaMemcpyDeviceToHost, st));
                    FA2_CUDA(cudaMemcpy2DAsync(job.dK + off + e0, pitch, dK_ + e0, pitch, w, cnt, cudaMemcpyDeviceToHost, st));
                    FA2_CUDA(cudaMemcpy2DAsync(job.dV + off + e0, pitch, dV_ + e0, pitch, w, cnt, cudaMemcpyDeviceToHost, st));
                }
            }
            FA2_CUDA(cudaStreamSynchronize(st));
            FA2_CUDA(cudaEventElapsedTime(&sh.ms[d], k0, k1));
            return FA2_OK;
        };
        // every device reaches every barrier, whatever failed (a peer waiting for an event would hang otherwise)
        auto step = [&](const std::function<int()>& f) {
            if (!sh.failed.load() && sh.rc[d] == 0) {
                const int rc = f();
                if (rc) { sh.rc[d] = rc; sh.err[d] = g_last_error; sh.failed.store(true); }
            }
            bar.arrive_and_wait();
        };
        step(phase_a);
        step(phase_b);
        step(phase_c);
        if (st) cudaStreamSynchronize(st);
        bar.arrive_and_wait();                   // nobody frees or reuses a buffer a peer may still be reading
        for (cudaEvent_t e : {k0, k1, sh.ev_side[d], sh.ev_bwd[d]}) if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
    };
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d) th.emplace_back(worker, d);
    for (auto& t : th) t.join();
    cudaSetDevice(prev_dev);
    float mx = 0.f;
    for (int d = 0; d < G; ++d) {
        if (sh.rc[d]) return fail(sh.rc[d], "device %d: %s", d, sh.err[d].c_str());
        if (sh.ms[d] > mx) mx = sh.ms[d];
    }
    if (kernel_ms) *kernel_ms = mx;
    return FA2_OK;
}

int host_dispatch(const HostJob& job, int n_gpus, float* kernel_ms) {
    int rc = check_shape(job.B, job.H, job.S, job.D);
    if (rc) return rc;
    if ((rc = check_precision(job.precision))) return rc;
    if (!job.Q || !job.K || !job.V) return fail(FA2_ERR_INVALID_ARGUMENT, "null input pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(FA2_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (n_gpus <= 0) n_gpus = 1;
    if (n_gpus > ndev) return fail(FA2_ERR_INVALID_ARGUMENT, "n_gpus=%d but only %d device(s) visible", n_gpus, ndev);
    const int BH = job.B * job.H;
    if (n_gpus > 1) {
        int g_bh = n_gpus, g_s = 1;
        choose_split(BH, job.S, n_gpus, &g_bh, &g_s);
        if (g_s > 1) return host_dispatch_seqsplit(job, g_bh, g_s, kernel_ms);
    }
    if (n_gpus > BH) n_gpus = BH;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    std::vector<std::thread> th;
    std::vector<int> rcs(n_gpus, 0);
    std::vector<float> ms(n_gpus, 0.f);
    std::vector<std::string> errs(n_gpus);
    for (int g = 0; g < n_gpus; ++g) {
        int bh0 = 0, cnt = 0;
        fa2_partition(BH, n_gpus, g, &bh0, &cnt);
        const int dev = (n_gpus == 1) ? prev_dev : g;      // one GPU: the caller's current device
        th.emplace_back([&, g, dev, bh0, cnt] { rcs[g] = host_worker(job, dev, bh0, cnt, &ms[g], &errs[g]); });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(prev_dev);
    float mx = 0.f;
    for (int g = 0; g < n_gpus; ++g) {
        if (rcs[g]) return fail(rcs[g], "device %d: %s", g, errs[g].c_str());
        if (ms[g] > mx) mx = ms[g];
    }
    if (kernel_ms) *kernel_ms = mx;
    return FA2_OK;
}

}  // namespace
}  // namespace fa2

using namespace fa2;

extern "C" {

int fa2_version(void) { return 100; }

const char* fa2_last_error(void) { return g_last_error.c_str(); }

size_t fa2_workspace_bytes(int B, int H, int S, int D, int mode) {
    if (B <= 0 || H <= 0 || S <= 0 || (D != 32 && D != 64 && D != 128)) return 0;
    return work_layout(static_cast<size_t>(B) * H * S, padded_head_dim(D), mode != FA2_MODE_FORWARD).total;   // incl. the 1 KB range block
}

int fa2_partition(int BH, int n_parts, int part, int* bh0, int* count) {
    if (BH < 0 || n_parts <= 0 || part < 0 || part >= n_parts || !bh0 || !count)
        return fail(FA2_ERR_INVALID_ARGUMENT, "bad partition request BH=%d n_parts=%d part=%d", BH, n_parts, part);
    const long long lo = static_cast<long long>(part) * BH / n_parts;
    const long long hi = static_cast<long long>(part + 1) * BH / n_parts;
    *bh0 = static_cast<int>(lo);
    *count = static_cast<int>(hi - lo);
    return FA2_OK;
}

int fa2_plan_chunks(int count, int S, int D, int mode, int* sizes, int max_chunks) {
    if (count < 0 || S <= 0 || D <= 0 || mode < FA2_MODE_FORWARD || mode > FA2_MODE_FORWARD_BACKWARD || (!sizes && max_chunks > 0))
        return -1;
    const std::vector<int> v = plan_chunks(count, S, D, mode != FA2_MODE_BACKWARD, mode != FA2_MODE_FORWARD);
    for (size_t i = 0; i < v.size() && static_cast<int>(i) < max_chunks; ++i) sizes[i] = v[i];
    return static_cast<int>(v.size());
}

int fa2_plan_split(int BH, int S, int n_gpus, int* g_bh, int* g_s) {
    if (BH <= 0 || S <= 0 || n_gpus <= 0 || !g_bh || !g_s) return fail(FA2_ERR_INVALID_ARGUMENT, "bad split request");
    if (n_gpus == 1) { *g_bh = 1; *g_s = 1; return FA2_OK; }
    choose_split(BH, S, n_gpus, g_bh, g_s);
    return FA2_OK;
}

int fa2_seq_range(int S, int parts, int part, int* r0, int* r1) {
    if (S <= 0 || parts <= 0 || part < 0 || part >= parts || !r0 || !r1) return fail(FA2_ERR_INVALID_ARGUMENT, "bad range request");
    seq_range(S, parts, part, r0, r1);
    return FA2_OK;
}

int fa2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

namespace {
std::mutex g_host_mu;
std::vector<void*> g_pinned;   // pointers that came from cudaMallocHost
}

void* fa2_host_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) == cudaSuccess && p) {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_pinned.push_back(p);
        return p;
    }
    cudaGetLastError();
    p = malloc(bytes);
    if (!p) fail(FA2_ERR_INVALID_ARGUMENT, "host allocation of %zu bytes failed", bytes);
    return p;
}

void fa2_host_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        for (size_t i = 0; i < g_pinned.size(); ++i)
            if (g_pinned[i] == p) {
                g_pinned.erase(g_pinned.begin() + i);
                cudaFreeHost(p);
                return;
            }
    }
    free(p);
}

#ifdef FA2_TIMELINE
int fa2_debug_set_timeline(void* dev_ptr) {
    g_timeline = static_cast<unsigned long long*>(dev_ptr);
    return FA2_OK;
}
#endif

int fa2_profile_enable(int on) {
    g_profile.store(on != 0);
    return FA2_OK;
}

int fa2_profile_read(float* ms, int* launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (const ProfSpan& sp : g_spans) {
        float t = 0.f;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) {
            if (ms) ms[sp.kind] += t;
            if (launches) launches[sp.kind] += 1;
        }
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    g_spans.clear();
    return FA2_OK;
}

long long fa2_profile_kernel_launches(void) { return g_kernel_launches.exchange(0); }

int fa2_release_workspaces(void) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return FA2_OK; }
    for (int d = 0; d < kMaxDevices; ++d) {
        std::lock_guard<std::mutex> pl(g_pipes[d].mu);      // waits for a host call in flight on this device
        std::lock_guard<std::mutex> lk(g_mu);                // same order as host_worker -> arena_reserve
        if (g_pipes[d].ready) {
            cudaSetDevice(d);
            cudaDeviceSynchronize();
            g_pipes[d].destroy();
        }
        for (Arena* a : {&g_work[d], &g_io[d]}) {
            if (a->ptr) {
                cudaSetDevice(d);
                cudaDeviceSynchronize();
                cudaFree(a->ptr);
                a->ptr = nullptr;
                a->bytes = 0;
            }
        }
    }
    cudaSetDevice(prev);
    return FA2_OK;
}

int fa2_forward(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int H, int S, int D,
                int precision, void* cuda_stream) {
    if (!Q || !K || !V || !O || !LSE) return fail(FA2_ERR_INVALID_ARGUMENT, "null pointer argument");
    if (!aligned16({Q, K, V, O})) return fail(FA2_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
    Prepared pr;
    int rc = prepare(&pr, B, H, S, D, precision, false);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if ((rc = is_small(pr) ? run_cast_small(pr, Q, K, V, nullptr, nullptr, st) : run_cast(pr, Q, K, V, st))) return rc;
    return run_fwd_main(pr, O, LSE, st);
}

int fa2_backward(const float* Q, const float* K, const float* V, const float* O, const float* dO, const float* LSE,
                 float* dQ, float* dK, float* dV, int B, int H, int S, int D, int precision, void* cuda_stream) {
    if (!Q || !K || !V || !O || !dO || !LSE || !dQ || !dK || !dV)
        return fail(FA2_ERR_INVALID_ARGUMENT, "null pointer argument");
    if (!aligned16({Q, K, V, O, dO, dQ, dK, dV})) return fail(FA2_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
    Prepared pr;
    int rc = prepare(&pr, B, H, S, D, precision, true);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (is_small(pr)) {
        if ((rc = run_cast_small(pr, Q, K, V, dO, dQ, st))) return rc;          // incl. dO cast, dQ = 0, all scales
        if ((rc = run_bwd_prepass(pr, O, dO, LSE, dQ, 2, st))) return rc;       // D_i, LSE * log2(e)
        return run_bwd_main(pr, dQ, dK, dV, st);
    }
    if ((rc = run_cast(pr, Q, K, V, st))) return rc;
    if ((rc = run_bwd_prepass(pr, O, dO, LSE, dQ, 3, st))) return rc;
    if ((rc = run_fix_do(pr, dO, st))) return rc;
    return run_bwd_main(pr, dQ, dK, dV, st);
}

int fa2_forward_backward(const float* Q, const float* K, const float* V, const float* dO, float* O, float* LSE,
                         float* dQ, float* dK, float* dV, int B, int H, int S, int D, int precision,
                         void* cuda_stream) {
    if (!Q || !K || !V || !O || !dO || !LSE || !dQ || !dK || !dV)
        return fail(FA2_ERR_INVALID_ARGUMENT, "null pointer argument");
    if (!aligned16({Q, K, V, O, dO, dQ, dK, dV})) return fail(FA2_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
    Prepared pr;
    int rc = prepare(&pr, B, H, S, D, precision, true);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // The backward's pre-pass is folded into the forward kernel here: its register-donor warps cast dO and
    // zero-fill dQ in the shadow of the tensor-core loop, its epilogue forms D_i and LSE*log2(e).
    if (is_small(pr)) {
        // one cooperative launch prepares everything that does not depend on the forward; the forward's epilogue
        // still forms D_i and LSE * log2(e) (unless dO is only 16-byte aligned, see below)
        int mask = fuse_mask() & 2;
        if (reinterpret_cast<uintptr_t>(dO) & 31u) mask = 0;
        if ((rc = run_cast_small(pr, Q, K, V, dO, dQ, st))) return rc;
        if ((rc = run_fwd_main(pr, O, LSE, st, dO, dQ, mask))) return rc;
        if (!mask && (rc = run_bwd_prepass(pr, O, dO, LSE, dQ, 2, st))) return rc;
        return run_bwd_main(pr, dQ, dK, dV, st);
    }
    if ((rc = run_cast(pr, Q, K, V, st))) return rc;           // one 16-bit copy serves both passes
    {
        int mask = fuse_mask();
        // the fused forward reads dO rows with 256-bit loads (ldg256_stream): a dO view that is only 16-byte aligned
        // goes through the stand-alone pre-pass instead
        if (reinterpret_cast<uintptr_t>(dO) & 31u) mask = 0;
        if ((rc = run_fwd_main(pr, O, LSE, st, dO, dQ, mask))) return rc;  // mask != 0: also prepares dO(16 bit) / D_i, LSE*log2e / dQ = 0
        if (mask != 3 && (rc = run_bwd_prepass(pr, O, dO, LSE, dQ, 3 & ~mask, st))) return rc;
    }
    if ((rc = run_fix_do(pr, dO, st))) return rc;
    return run_bwd_main(pr, dQ, dK, dV, st);
}

int fa2_host_forward(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int H, int S,
                     int D, int precision, int n_gpus, float* kernel_ms) {
    if (!O || !LSE) return fail(FA2_ERR_INVALID_ARGUMENT, "null output pointer");
    HostJob j{Q, K, V, nullptr, nullptr, nullptr, O, LSE, nullptr, nullptr, nullptr, B, H, S, D, precision,
              FA2_MODE_FORWARD};
    return host_dispatch(j, n_gpus, kernel_ms);
}

int fa2_host_backward(const float* Q, const float* K, const float* V, const float* O, const float* dO,
                      const float* LSE, float* dQ, float* dK, float* dV, int B, int H, int S, int D, int precision,
                      int n_gpus, float* kernel_ms) {
    if (!O || !dO || !LSE || !dQ || !dK || !dV) return fail(FA2_ERR_INVALID_ARGUMENT, "null pointer argument");
    HostJob j{Q, K, V, O, dO, LSE, nullptr, nullptr, dQ, dK, dV, B, H, S, D, precision, FA2_MODE_BACKWARD};
    return host_dispatch(j, n_gpus, kernel_ms);
}

int fa2_host_forward_backward(const float* Q, const float* K, const float* V, const float* dO, float* O, float* LSE,
                              float* dQ, float* dK, float* dV, int B, int H, int S, int D, int precision,
                              int n_gpus, float* kernel_ms) {
    if (!O || !dO || !LSE || !dQ || !dK || !dV) return fail(FA2_ERR_INVALID_ARGUMENT, "null pointer argument");
    HostJob j{Q, K, V, nullptr, dO, nullptr, O, LSE, dQ, dK, dV, B, H, S, D, precision, FA2_MODE_FORWARD_BACKWARD};
    return host_dispatch(j, n_gpus, kernel_ms);
}

}  // extern "C"
