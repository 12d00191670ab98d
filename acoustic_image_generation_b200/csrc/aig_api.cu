// libaig.so - C ABI (include/aig.h) over the sm_100a kernels of this directory.
//
// Host-side responsibilities: argument validation, pointer classification (device / pinned /
// pageable host), staging of host buffers through handle-owned device scratch (double-buffered
// on a copy stream for the large spectrum input), TMA descriptor encoding, launch geometry
// (persistent grids sized from the SM count), error reporting.  No compute happens on the host
// except aig_auc's 10-term trapezoid.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aig.h"
#include "host_staging.h"
#include "aig_common.cuh"
#include "energy_kernel.cuh"
#include "frontend_kernel.cuh"
#include "fused_kernel.cuh"
#include "heatmap_kernel.cuh"
#include "mfcc_kernel.cuh"
#include "score_kernel.cuh"
#include "mask_packed_kernel.cuh"

#ifndef AIG_BUILD_ID_STRING
#define AIG_BUILD_ID_STRING "unstamped"
#endif

namespace {

using namespace aig;

// ---- reference tables as host data (generated) --------------------------------------------------
const unsigned short kRefNzBin[AIG_REF_NNZ] = AIG_REF_NZ_BIN;
const unsigned char kRefNzCol[AIG_REF_NNZ] = AIG_REF_NZ_COL;
const double kRefNzVal[AIG_REF_NNZ] = AIG_REF_NZ_VAL;
const double kRefDct[AIG_REF_FILTER_NUM * AIG_REF_MFCC_NUM] = AIG_REF_DCT;
const double kRefLifter[AIG_REF_MFCC_NUM] = AIG_REF_LIFTER;
const double kRefMfnorm = AIG_REF_MFNORM;

thread_local std::string g_create_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum MemKind { kDevice, kHostPinned, kHostPageable };

MemKind classify(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return kHostPageable;
    }
    switch (attr.type) {
        case cudaMemoryTypeDevice:
        case cudaMemoryTypeManaged:
            return kDevice;
        case cudaMemoryTypeHost:
            return kHostPinned;
        default:
            return kHostPageable;
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Entry points that touch several devices, or a device other than the caller's, put the calling thread's current device
// back (torch and other CUDA users of the thread rely on it).  Calls on a handle leave that handle's device current.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) { cudaSetDevice(prev); cudaGetLastError(); } }
};

// ---- NCCL through dlopen --------------------------------------------------------------------------
struct NcclId { char internal[AIG_COMM_ID_BYTES]; };
struct NcclApi {
    bool tried = false, ok = false;
    int (*get_unique_id)(NcclId*) = nullptr;
    int (*comm_init_rank)(void**, int, NcclId, int) = nullptr;
    int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*comm_destroy)(void*) = nullptr;
    int (*comm_init_all)(void**, int, const int*) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    const char* (*get_error_string)(int) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.tried) return api;
    api.tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return api;
    api.get_unique_id = reinterpret_cast<decltype(api.get_unique_id)>(dlsym(lib, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<decltype(api.comm_init_rank)>(dlsym(lib, "ncclCommInitRank"));
    api.all_reduce = reinterpret_cast<decltype(api.all_reduce)>(dlsym(lib, "ncclAllReduce"));
    api.comm_destroy = reinterpret_cast<decltype(api.comm_destroy)>(dlsym(lib, "ncclCommDestroy"));
    api.get_error_string = reinterpret_cast<decltype(api.get_error_string)>(dlsym(lib, "ncclGetErrorString"));
    api.comm_init_all = reinterpret_cast<decltype(api.comm_init_all)>(dlsym(lib, "ncclCommInitAll"));
    api.group_start = reinterpret_cast<decltype(api.group_start)>(dlsym(lib, "ncclGroupStart"));
    api.group_end = reinterpret_cast<decltype(api.group_end)>(dlsym(lib, "ncclGroupEnd"));
    api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
    return api;
}
constexpr int kNcclInt64 = 4;   // ncclInt64
constexpr int kNcclSum = 0;     // ncclSum

}  // namespace

constexpr size_t kStagedUploadMinBytes = size_t(8) << 20;      // default of option staged_min_bytes
constexpr size_t kStagedDownloadMinBytes = size_t(8) << 20;
struct aig_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};
    int sm_count = 0;
    std::string err;
    // tables
    bool tables_set = false, tables_ref = false;
    int fft_len = 0, filter_num = 0, mfcc_num = 0;
    double mfnorm = 0.0;
    double* d_tables = nullptr;   // bank | dct | lifter (generic kernel)
    // scratch: one grow-only arena plus per-call overflow blocks folded in at the next reset
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    std::vector<std::pair<void*, size_t>> overflow;
    // small host buffers (a frame, a few vectors) skip cudaMemcpy altogether: they are copied into this pinned, device-
    // mapped arena and the kernels read / write it directly over PCIe (zero-copy); see Io
    char* pinned = nullptr;
    size_t pinned_cap = 0, pinned_off = 0;
    int small_host_bytes = 128 * 1024;  // per-buffer limit of the zero-copy path; 0 switches it off
    // chained MFCC -> energy pipeline
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_chain[4] = {nullptr, nullptr, nullptr, nullptr};   // start, mfcc[2], energy-done
    int chain_chunk_frames = 512;
    bool chain_overlap = true;
    int chain_mode = 2;                 // 2: one fused persistent kernel; 1/0: two kernels (see chain_overlap)
    int fused_variant = 2;              // 96 KiB stages x 2 measured best (profiles/r01_tune_chain.txt)
    int64_t launch_row_limit = (int64_t(1) << 31) - 1024;   // TMA coordinates are int32: rows per launch (testable via option)
    int heatmap_exact = 0;              // 1: float64 replica of the oracle's bilinear; 0: float32 fast path
    bool heat_attr_set = false;
    bool heat_exact_attr_set = false;
    bool heat_stream_attr_set[8] = {false, false, false, false, false, false, false, false};
    bool stage2_attr_set[2] = {false, false};
    int heat_bulk_store = 1;            // 0: round-1 per-thread-store kernel for every shape (comparison runs)
    bool norm_bulk_attr_set = false;
    bool packed_attr_set[4] = {};
    int packed_ctas[4] = {};
    int mask_packed = 1;                // aig_resize_mask / aig_ciou_sweep at 224 x 298 and 224 x 224 as the packed kernels (0: the generic kernels)
    bool energy_heat_ws_attr_set[5] = {};
    int energy_wide = 1;                // aig_energy / aig_acivw_batch on large batches: eight frames (four pairs) per 512-thread CTA around one conflict-free exp table (0: stage2_kernel, eight CTAs of 64 (four of 128) threads per SM)
    bool stage2_wide_attr_set[2] = {};
    int64_t acivw_wide_pairs = 8192;    // aig_acivw_batch takes the wide form from this many pairs up
    int overlay_luma = 1;               // aig_overlay keeps the luma plane of a frame in shared memory between its passes (0: BGR read twice)
    bool overlay_luma_attr_set = false;
    int energy_heat_ws = 1;             // aig_energy_heatmap as the warp-specialised kernel (0: heat_stream_kernel<true>, phases in sequence)
    int norm_bulk_copy = 1;             // aig_normalize_images with the frame resident in shared memory (0: two-pass per-thread kernel)
    int small_batch_frames = 0;         // below this many frames a frame is split over a cluster of 8 CTAs; 0: SM count
    unsigned int debug_jitter = 0;      // non-zero: seed of the jittered build of the fused kernel (race stress tests)
    bool mask_attr_set = false;
    int l2_evict_first = 0;             // L2 evict-first hint on the spectrum loads (measured slower: off)
    int keep_mfcc_in_l2 = 1;            // fused kernel: evict-last hint on the MFCC stores the energy warps re-read
    bool fused_attr_set[8] = {false, false, false, false, false, false, false, false};
    bool fused_heat_attr_set[3] = {false, false, false};
    int chain_energy_ctas_per_sm = 3;   // footprint of the overlapped energy kernel (measured: profiles/r01_tune_chain.txt)
    // profiling: (start, stop) event pairs per launch, by kernel kind
    bool profile = false;
    struct Span { cudaEvent_t start, stop; int kind; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double2* d_twiddle = nullptr;       // exp(-2*pi*i*k/1024), k < 512 (aig_power_spectrum)
    // pageable host inputs are staged through a pinned ring by a few copy threads (host_staging.h)
    aig::StagedUploader uploader;
    size_t staged_min_bytes = kStagedUploadMinBytes;   // pageable copies from this size on go through the staging ring
    int host_copy_threads = -1;         // -1: min(6, hardware threads / 2); 0: leave pageable copies to the driver
    // NCCL communicator (resolved with dlopen; see aig_comm_init)
    void* comm = nullptr;
    int comm_world = 1;
    // misc
    int variant = -1;
    int64_t launches = 0;
    EncodeTiledFn encode = nullptr;
    bool smem_attr_set[16] = {false};

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int fail_cuda(cudaError_t e, const char* what) {
        cudaGetLastError();
        return fail(AIG_ERR_CUDA_BASE - static_cast<int>(e), "%s: %s (%s)", what, cudaGetErrorName(e),
                    cudaGetErrorString(e));
    }
};

#define AIG_CK(call)                                                   \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return h->fail_cuda(e_, #call);         \
    } while (0)

namespace {

// ---- scratch ------------------------------------------------------------------------------------
int scratch_reset(aig_handle* h) {
    if (!h->overflow.empty()) {
        size_t total = h->arena_cap;
        for (auto& b : h->overflow) total += align_up(b.second, 256);
        AIG_CK(cudaStreamSynchronize(h->stream));
        AIG_CK(cudaStreamSynchronize(h->copy_stream));
        for (auto& b : h->overflow) cudaFree(b.first);
        h->overflow.clear();
        if (h->arena) cudaFree(h->arena);
        h->arena = nullptr;
        h->arena_cap = 0;
        total = align_up(total + total / 4, 1 << 20);
        if (cudaMalloc(&h->arena, total) != cudaSuccess) {
            cudaGetLastError();
            return h->fail(AIG_ERR_ALLOC, "cudaMalloc of %zu scratch bytes failed", total);
        }
        h->arena_cap = total;
    }
    h->arena_off = 0;
    h->pinned_off = 0;
    return AIG_OK;
}

// A slot of the pinned zero-copy arena, or null when the buffer is too large / the arena is full or unavailable.
constexpr size_t kPinnedArenaBytes = size_t(1) << 20;
void* pinned_slot(aig_handle* h, size_t bytes) {
    if (h->small_host_bytes <= 0 || bytes > static_cast<size_t>(h->small_host_bytes)) return nullptr;
    if (h->pinned == nullptr) {
        if (h->pinned_cap == size_t(-1)) return nullptr;                   // allocation failed before: do not retry
        if (cudaHostAlloc(reinterpret_cast<void**>(&h->pinned), kPinnedArenaBytes, cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            h->pinned = nullptr;
            h->pinned_cap = size_t(-1);
            return nullptr;
        }
        h->pinned_cap = kPinnedArenaBytes;
    }
    const size_t need = align_up(std::max<size_t>(bytes, 1), 256);
    if (h->pinned_off + need > h->pinned_cap) return nullptr;
    void* p = h->pinned + h->pinned_off;
    h->pinned_off += need;
    return p;
}

void* scratch(aig_handle* h, size_t bytes) {
    bytes = align_up(std::max<size_t>(bytes, 1), 256);
    if (h->arena_off + bytes <= h->arena_cap) {
        void* p = h->arena + h->arena_off;
        h->arena_off += bytes;
        return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        h->fail(AIG_ERR_ALLOC, "cudaMalloc of %zu scratch bytes failed", bytes);
        return nullptr;
    }
    h->overflow.emplace_back(p, bytes);
    return p;
}

// Host -> device copy of `bytes` on `stream`.  Large pageable sources go through the handle's StagedUploader (4-5x the
// driver's own pageable path); pinned sources and small copies are plain cudaMemcpyAsync.
bool use_host_staging(aig_handle* h, size_t bytes, size_t min_bytes, MemKind kind);
cudaError_t upload_async(aig_handle* h, void* dst, const void* src, size_t bytes, MemKind kind, cudaStream_t stream) {
    if (use_host_staging(h, bytes, h->staged_min_bytes, kind)) return h->uploader.upload(dst, src, bytes, stream);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
}

// Device -> host copy after the work enqueued on `stream`; large pageable destinations are drained through the same
// pinned ring (returns with the data in place), everything else is a plain asynchronous copy.
bool use_host_staging(aig_handle* h, size_t bytes, size_t min_bytes, MemKind kind) {
    if (kind != kHostPageable || bytes < min_bytes || h->host_copy_threads == 0) return false;
    int threads = h->host_copy_threads;
    if (threads < 0) threads = static_cast<int>(std::min(6u, std::max(1u, std::thread::hardware_concurrency() / 2)));   // 4-6 fill the link
    return h->uploader.start(threads);
}
cudaError_t download(aig_handle* h, void* dst, const void* src, size_t bytes, MemKind kind, cudaStream_t stream) {
    // (results under 8 MiB stay with the driver: draining needs the worker threads, whose wake-up a small result cannot repay)
    if (use_host_staging(h, bytes, kStagedDownloadMinBytes, kind)) return h->uploader.download(dst, src, bytes, stream);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream);
}

// ---- host/device staging of small and medium buffers ----------------------------------------------
struct Io {
    aig_handle* h;
    bool any_host = false;
    bool failed = false;
    bool zero_copy_ok = false;       // set by latency-bound entry points (one-frame find_logen): small host buffers are then
                                     // read / written by the kernel in pinned mapped memory instead of being copied
    int code = AIG_OK;               // first failure's status
    struct Pending { void* host; void* dev; size_t bytes; bool zero_copy; };
    std::vector<Pending> outs;
    explicit Io(aig_handle* handle) : h(handle) {}
    // Error exit of an entry point: copies out of caller memory may already be in flight - they must have finished
    // before the caller gets its buffers back.
    int abort(int rc) {
        if (any_host) { cudaStreamSynchronize(h->stream); cudaGetLastError(); }
        return rc;
    }

    template <typename T>
    const T* in(const T* p, size_t count) {
        if (p == nullptr) return nullptr;
        const MemKind kind = classify(p);
        if (kind == kDevice) return p;
        any_host = true;
        if (void* z = zero_copy_ok ? pinned_slot(h, count * sizeof(T)) : nullptr) {   // the kernel reads the pinned copy itself
            std::memcpy(z, p, count * sizeof(T));
            return static_cast<const T*>(z);
        }
        void* d = scratch(h, count * sizeof(T));
        if (!d) { failed = true; code = AIG_ERR_ALLOC; return nullptr; }
        const cudaError_t status = upload_async(h, d, p, count * sizeof(T), kind, h->stream);
        if (status != cudaSuccess) {
            code = h->fail_cuda(status, "host-to-device copy");
            failed = true;
            return nullptr;
        }
        return static_cast<const T*>(d);
    }
    // in/out buffer (accumulators): copied in now, copied back at finish()
    template <typename T>
    T* inout(T* p, size_t count) {
        if (p == nullptr) return nullptr;
        if (classify(p) == kDevice) return p;
        const bool keep = zero_copy_ok;
        zero_copy_ok = false;                    // accumulators take device atomics: always a device copy
        T* d = const_cast<T*>(in(const_cast<const T*>(p), count));
        zero_copy_ok = keep;
        if (d) outs.push_back({p, d, count * sizeof(T), false});
        return d;
    }
    template <typename T>
    T* out(T* p, size_t count) {
        if (p == nullptr) return nullptr;
        if (classify(p) == kDevice) return p;
        any_host = true;
        if (void* z = zero_copy_ok ? pinned_slot(h, count * sizeof(T)) : nullptr) {   // the kernel writes the pinned slot itself
            outs.push_back({p, z, count * sizeof(T), true});
            return static_cast<T*>(z);
        }
        void* d = scratch(h, count * sizeof(T));
        if (!d) { failed = true; code = AIG_ERR_ALLOC; return nullptr; }
        outs.push_back({p, d, count * sizeof(T), false});
        return static_cast<T*>(d);
    }
    int finish() {
        if (failed) return abort(code != AIG_OK ? code : h->fail(AIG_ERR_ALLOC, "staging failed"));
        for (auto& o : outs) {
            if (o.zero_copy) continue;
            const cudaError_t status = download(h, o.host, o.dev, o.bytes, classify(o.host), h->stream);
            if (status != cudaSuccess) return abort(h->fail_cuda(status, "device-to-host copy"));
        }
        if (any_host) AIG_CK(cudaStreamSynchronize(h->stream));
        for (auto& o : outs)
            if (o.zero_copy) std::memcpy(o.host, o.dev, o.bytes);           // the kernel wrote the pinned slot; hand it over
        return AIG_OK;
    }
};

enum KernelKind { kKindMfcc = 0, kKindEnergy = 1, kKindOther = 2 };

cudaEvent_t pooled_event(aig_handle* h) {
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Brackets one launch with events on the launching stream when profiling is on.
struct LaunchScope {
    aig_handle* h;
    cudaStream_t stream;
    cudaEvent_t start = nullptr;
    int kind;
    LaunchScope(aig_handle* handle, cudaStream_t s, int k) : h(handle), stream(s), kind(k) {
        if (h->profile) {
            start = pooled_event(h);
            cudaEventRecord(start, stream);
        }
    }
    int done(const char* name) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return h->fail_cuda(e, name);
        h->launches += 1;
        if (start) {
            cudaEvent_t stop = pooled_event(h);
            cudaEventRecord(stop, stream);
            h->spans.push_back({start, stop, kind});
        }
        return AIG_OK;
    }
};

int frames_grid(const aig_handle* h, int64_t n_frames, int per_sm) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_frames, static_cast<int64_t>(h->sm_count) * per_sm)));
}

// ---- fused MFCC kernel variants -------------------------------------------------------------------
struct Variant { int rows, slabs, stages, ctas; };
constexpr int kNumVariants = 10;
// Ring geometry {spectra per tile, 32-bin slabs per stage, stages, CTAs per SM}.  A stage count that
// divides the 16/slabs stages of a tile lets the compiler resolve every ring address statically.
constexpr Variant kVariants[kNumVariants] = {
    {128, 1, 4, 3},    // 0: 16 KiB stages, 64 KiB ring, 3 CTAs/SM (192 KiB of loads in flight per SM)
    {128, 2, 2, 3},    // 1: 32 KiB stages (256 B contiguous per spectrum)
    {128, 1, 8, 1},    // 2: one CTA/SM, 128 KiB ring
    {64, 2, 4, 3},     // 3
    {64, 4, 2, 3},     // 4: 512 B contiguous per spectrum
    {128, 4, 2, 1},    // 5: 64 KiB stages
    {64, 1, 4, 6},     // 6: many small CTAs
    {32, 16, 1, 3},    // 7: whole spectra per stage, pipelining across CTAs only
    {256, 1, 4, 1},    // 8: 256-spectrum tiles
    {128, 1, 6, 2},    // 9: run-time ring index (6 does not divide 16)
};
constexpr int kDefaultVariant = 8;   // 256-spectrum tiles, one CTA/SM: best for aig_mfcc alone (profiles/r01_tune_fused_and_hint.txt)

template <int V>
int launch_banded_variant(aig_handle* h, const CUtensorMap& map, float* out, unsigned n_rows, int flip180,
                          unsigned frame_pixels) {
    constexpr Variant v = kVariants[V];
    using P = MfccPipe<v.rows, v.slabs, v.stages>;
    auto kernel = mfcc_banded_kernel<v.rows, v.slabs, v.stages, v.ctas>;
    if (!h->smem_attr_set[V]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kSmemBytes));
        h->smem_attr_set[V] = true;
    }
    const unsigned n_tiles = (n_rows + v.rows - 1) / v.rows;
    const unsigned grid = std::min<unsigned>(n_tiles, static_cast<unsigned>(h->sm_count * v.ctas));
    LaunchScope scope(h, h->stream, kKindMfcc);
    kernel<<<grid, P::kThreads, P::kSmemBytes, h->stream>>>(map, out, n_rows, n_tiles, flip180, frame_pixels,
                                                            h->l2_evict_first);
    return scope.done("mfcc_banded_kernel");
}

int encode_spectrum_map(aig_handle* h, const float* d_power, uint64_t n_rows, int box_rows, CUtensorMap* map) {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kFftLen), static_cast<cuuint64_t>(n_rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kFftLen) * sizeof(float)};
    const cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t elem[2] = {1u, 1u};
    CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_power), dims, strides, box,
                           elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return h->fail(AIG_ERR_CUDA_BASE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return AIG_OK;
}

// d_power / d_out are device pointers; rows are processed in launches of < 2^31 rows.
int launch_mfcc(aig_handle* h, const float* d_power, int64_t n_rows, float* d_out, int flip180, int frame_pixels) {
    if (n_rows == 0) return AIG_OK;
    if (!h->tables_ref) {
        const double* bank = h->d_tables;
        const double* dct = bank + static_cast<size_t>(h->fft_len) * h->filter_num;
        const double* lifter = dct + static_cast<size_t>(h->filter_num) * h->mfcc_num;
        const int64_t blocks = std::min<int64_t>((n_rows + kGenericWarps - 1) / kGenericWarps,
                                                 static_cast<int64_t>(h->sm_count) * 8);
        LaunchScope scope(h, h->stream, kKindMfcc);
        mfcc_generic_kernel<<<static_cast<unsigned>(blocks), kGenericWarps * 32, 0, h->stream>>>(
            d_power, n_rows, h->fft_len, h->filter_num, h->mfcc_num, bank, dct, lifter, h->mfnorm, d_out, flip180,
            frame_pixels);
        return scope.done("mfcc_generic_kernel");
    }
    if ((reinterpret_cast<uintptr_t>(d_power) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc: device buffers must be 16-byte aligned");
    const int variant = h->variant < 0 ? kDefaultVariant : h->variant;
    const int64_t unit = flip180 ? frame_pixels : 256;
    const int64_t max_rows = std::max<int64_t>(unit, h->launch_row_limit / unit * unit);
    for (int64_t done = 0; done < n_rows; done += max_rows) {
        const int64_t rows = std::min(max_rows, n_rows - done);
        CUtensorMap map;
        int rc = encode_spectrum_map(h, d_power + done * kFftLen, static_cast<uint64_t>(rows),
                                     kVariants[variant].rows, &map);
        if (rc != AIG_OK) return rc;
        float* out = d_out + done * kMfccNum;
        const unsigned r = static_cast<unsigned>(rows), fp = static_cast<unsigned>(frame_pixels);
        switch (variant) {
            case 0: rc = launch_banded_variant<0>(h, map, out, r, flip180, fp); break;
            case 1: rc = launch_banded_variant<1>(h, map, out, r, flip180, fp); break;
            case 2: rc = launch_banded_variant<2>(h, map, out, r, flip180, fp); break;
            case 3: rc = launch_banded_variant<3>(h, map, out, r, flip180, fp); break;
            case 4: rc = launch_banded_variant<4>(h, map, out, r, flip180, fp); break;
            case 5: rc = launch_banded_variant<5>(h, map, out, r, flip180, fp); break;
            case 6: rc = launch_banded_variant<6>(h, map, out, r, flip180, fp); break;
            case 7: rc = launch_banded_variant<7>(h, map, out, r, flip180, fp); break;
            case 8: rc = launch_banded_variant<8>(h, map, out, r, flip180, fp); break;
            case 9: rc = launch_banded_variant<9>(h, map, out, r, flip180, fp); break;
            default: rc = h->fail(AIG_ERR_ARGUMENT, "unknown MFCC kernel variant %d", variant);
        }
        if (rc != AIG_OK) return rc;
    }
    return AIG_OK;
}

// Frames below which one frame is spread over a thread-block cluster of 8 CTAs (stage2_cluster_kernel): with fewer frames
// than SMs a CTA per frame leaves SMs idle and a call costs a whole frame's 27 rounds.
int64_t small_batch_limit(const aig_handle* h) { return h->small_batch_frames > 0 ? h->small_batch_frames : h->sm_count; }

// aig_energy (GROUPS = 1) and aig_acivw_batch (GROUPS = 2) on device buffers.
template <int GROUPS>
int launch_stage2(aig_handle* h, cudaStream_t stream, const Stage2Args& args, int ctas_per_sm = 8) {
    if (args.n_frames == 0) return AIG_OK;
    if (!h->stage2_attr_set[GROUPS - 1]) {
        // 8 CTAs of 20 KiB (4 of 45 KiB) per SM only fit with the shared-memory carveout at its maximum
        cudaFuncSetAttribute(stage2_kernel<GROUPS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaGetLastError();
        h->stage2_attr_set[GROUPS - 1] = true;
    }
    LaunchScope scope(h, stream, kKindEnergy);
    if (args.n_frames < small_batch_limit(h)) {
        const int64_t clusters = std::min<int64_t>(args.n_frames, 4 * h->sm_count);
        stage2_cluster_kernel<GROUPS><<<static_cast<unsigned>(clusters * kClusterSize), GROUPS * kClusterGroupThreads, 0, stream>>>(args);
        return scope.done("stage2_cluster_kernel");
    }
    // (the pair form only pays from 8192 pairs up: 6.65 against 6.53 M pairs/s there, 5.74 against 6.06 M at 2048 - tools/energy_sizes_probe.py)
    if (ctas_per_sm >= 8 && h->energy_wide && (GROUPS == 1 || args.n_frames >= h->acivw_wide_pairs)) {
        // eight frames (four pairs) per CTA of 512 threads, one CTA per SM, around one conflict-free exponential table
        if (!h->stage2_wide_attr_set[GROUPS - 1]) {
            AIG_CK(cudaFuncSetAttribute(stage2_wide_kernel<GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(WideShared<GROUPS>))));
            h->stage2_wide_attr_set[GROUPS - 1] = true;
        }
        const int64_t ctas = std::min<int64_t>(args.n_frames, h->sm_count);
        stage2_wide_kernel<GROUPS><<<static_cast<unsigned>(ctas), kWideGroups * kEnergyThreads, sizeof(WideShared<GROUPS>), stream>>>(args);
        return scope.done("stage2_wide_kernel");
    }
    stage2_kernel<GROUPS><<<frames_grid(h, args.n_frames, std::min(ctas_per_sm, 8 / GROUPS)), GROUPS * kEnergyThreads, 0, stream>>>(args);
    return scope.done("stage2_kernel");
}

int launch_energy(aig_handle* h, cudaStream_t stream, const float* d_images, int64_t n_frames, int normalize_first,
                  float* d_scaled, double* d_energy, uint8_t* d_mask, double* d_mean, int ctas_per_sm = 8) {
    Stage2Args args = {};
    args.img[0] = d_images; args.scaled[0] = d_scaled; args.energy[0] = d_energy; args.mask[0] = d_mask; args.mean[0] = d_mean;
    args.n_frames = n_frames; args.normalize_first = normalize_first;
    return launch_stage2<1>(h, stream, args, ctas_per_sm);
}

// resize_mask_kernel / ciou_sweep_kernel keep 36 blended rows of out_w uint16 in shared memory: up to 184 KiB at 2048 x 2048
constexpr size_t kMaskSmemLimit = 200 * 1024;
int allow_mask_smem(aig_handle* h) {
    if (!h->mask_attr_set) {
        AIG_CK(cudaFuncSetAttribute(resize_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaskSmemLimit)));
        AIG_CK(cudaFuncSetAttribute(ciou_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaskSmemLimit)));
        h->mask_attr_set = true;
    }
    return AIG_OK;
}
int mask_ctas_per_sm(size_t smem) {
    return static_cast<int>(std::max<size_t>(1, std::min<size_t>(8, kMaskSmemLimit / (smem + 4 * 1024))));
}

// The reference's two output sizes run the packed mask kernels (mask_packed_kernel.cuh).
bool packed_size(const aig_handle* h, int out_h, int out_w) { return h->mask_packed && out_h == 224 && (out_w == 298 || out_w == 224); }
// CTAs of a packed kernel that are resident per SM (asked of the runtime once): the grid is exactly one wave, a partial
// second wave cost 25 % (ncu: 1.5 waves with a grid sized by shared memory alone).
template <typename Kernel>
int packed_ctas_per_sm(aig_handle* h, Kernel kernel, size_t smem, int slot) {
    if (h->packed_ctas[slot] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kPackedThreads, smem) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
        h->packed_ctas[slot] = n;
    }
    return h->packed_ctas[slot];
}
template <int W, int H>
int launch_resize_packed(aig_handle* h, const uint8_t* d_mask, int64_t n_frames, uint8_t* d_up, int slot) {
    auto kernel = resize_mask_packed_kernel<W, H>;
    const size_t smem = sizeof(ResizePackedSmem<W, H>);
    if (!h->packed_attr_set[slot]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        h->packed_attr_set[slot] = true;
    }
    LaunchScope scope(h, h->stream, kKindOther);
    kernel<<<frames_grid(h, n_frames, packed_ctas_per_sm(h, kernel, smem, slot)), kPackedThreads, smem, h->stream>>>(d_mask, n_frames, d_up);
    return scope.done("resize_mask_packed_kernel");
}

// heat_stream_kernel needs every output row pair to start on a 16-byte boundary and be a 16-byte multiple long.
bool heat_stream_ok(const aig_handle* h, int out_h, int out_w, const float* d_heat, bool fused) {
    return h->heat_bulk_store && !h->heatmap_exact && out_w % 2 == 0 && (static_cast<long long>(out_h) * out_w) % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(d_heat) & 15u) == 0 && heat_stream_layout(out_h, out_w, fused).total <= 120 * 1024;
}

template <bool FUSED, int VEC, int W, int H>
int launch_heat_stream_variant(aig_handle* h, const HeatStreamArgs& args, int slot) {
    auto kernel = heat_stream_kernel<FUSED, VEC, W, H>;
    if (!h->heat_stream_attr_set[slot]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        h->heat_stream_attr_set[slot] = true;
    }
    const size_t smem = heat_stream_layout(args.out_h, args.out_w, FUSED).total;
    const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (smem + 1024 + 512))));
    const int grid = frames_grid(h, args.n_frames, per_sm);
    LaunchScope scope(h, h->stream, FUSED ? kKindEnergy : kKindOther);
    kernel<<<grid, kStreamThreads, smem, h->stream>>>(args);
    return scope.done("heat_stream_kernel");
}

// The reference's two output sizes run builds with the size as a compile-time constant; anything else the generic one.
template <bool FUSED>
int launch_heat_stream(aig_handle* h, const HeatStreamArgs& args) {
    const int base = FUSED ? 4 : 0;
    if (args.out_h == 224 && args.out_w == 298) return launch_heat_stream_variant<FUSED, 2, 298, 224>(h, args, base + 0);
    if (args.out_h == 224 && args.out_w == 224) return launch_heat_stream_variant<FUSED, 4, 224, 224>(h, args, base + 1);
    if (args.out_w % 4 == 0) return launch_heat_stream_variant<FUSED, 4, 0, 0>(h, args, base + 2);
    return launch_heat_stream_variant<FUSED, 2, 0, 0>(h, args, base + 3);
}

// aig_energy_heatmap as one warp-specialised launch (energy_heat_ws_kernel): float64 warps and heat-map warps of a CTA
// overlap, two CTAs of 352 + 160 threads per SM (WsTwin; the other splits measured are listed in heatmap_kernel.cuh).
// Taken when its rows and staging slots fit in half an SM's shared memory (both of the reference's sizes do); otherwise,
// or with option energy_heat_ws = 0, heat_stream_kernel<true>.
template <typename C>
size_t ws_smem_limit() { return std::min<size_t>((228 * 1024 - C::CTAS * 1024) / C::CTAS / 16 * 16, 227 * 1024); }
template <typename C>
bool energy_heat_ws_fits(int out_h, int out_w) { return energy_heat_ws_smem<C>(out_h, out_w) <= ws_smem_limit<C>(); }
bool energy_heat_ws_ok(const aig_handle* h, int out_h, int out_w, const float* d_heat) {
    return h->energy_heat_ws && heat_stream_ok(h, out_h, out_w, d_heat, false) && energy_heat_ws_fits<WsTwin>(out_h, out_w);
}
template <typename C, int VEC, int W, int H>
int launch_energy_heat_ws_variant(aig_handle* h, const HeatStreamArgs& args, int slot) {
    auto kernel = energy_heat_ws_kernel<C, VEC, W, H>;
    if (!h->energy_heat_ws_attr_set[slot]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ws_smem_limit<C>())));
        h->energy_heat_ws_attr_set[slot] = true;
    }
    LaunchScope scope(h, h->stream, kKindEnergy);
    kernel<<<frames_grid(h, args.n_frames, C::CTAS), C::THREADS, energy_heat_ws_smem<C>(args.out_h, args.out_w), h->stream>>>(args);
    return scope.done("energy_heat_ws_kernel");
}
template <typename C>
int launch_energy_heat_ws(aig_handle* h, const HeatStreamArgs& args, int base) {
    if (h->debug_jitter != 0) {                  // the jittered build exists for the generic-size kernel only (race stress tests)
        HeatStreamArgs jittered = args;
        jittered.jitter_seed = h->debug_jitter;
        auto kernel = energy_heat_ws_kernel<C, 2, 0, 0, true>;
        if (!h->energy_heat_ws_attr_set[base + 4]) {
            AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ws_smem_limit<C>())));
            h->energy_heat_ws_attr_set[base + 4] = true;
        }
        LaunchScope scope(h, h->stream, kKindEnergy);
        kernel<<<frames_grid(h, args.n_frames, C::CTAS), C::THREADS, energy_heat_ws_smem<C>(args.out_h, args.out_w), h->stream>>>(jittered);
        return scope.done("energy_heat_ws_kernel<jitter>");
    }
    if (args.out_h == 224 && args.out_w == 298) return launch_energy_heat_ws_variant<C, 2, 298, 224>(h, args, base + 0);
    if (args.out_h == 224 && args.out_w == 224) return launch_energy_heat_ws_variant<C, 4, 224, 224>(h, args, base + 1);
    if (args.out_w % 4 == 0) return launch_energy_heat_ws_variant<C, 4, 0, 0>(h, args, base + 2);
    return launch_energy_heat_ws_variant<C, 2, 0, 0>(h, args, base + 3);
}

int launch_heatmap(aig_handle* h, const double* d_energy, int64_t n_frames, int out_h, int out_w, float* d_heat) {
    if (heat_stream_ok(h, out_h, out_w, d_heat, false)) {
        HeatStreamArgs args = {};
        args.energy_in = d_energy; args.n_frames = n_frames; args.out_h = out_h; args.out_w = out_w; args.heat = d_heat;
        return launch_heat_stream<false>(h, args);
    }
    const size_t fast_smem = (static_cast<size_t>(kFrameH) * out_w + 2 * static_cast<size_t>(out_w + out_h)) * sizeof(float);
    LaunchScope scope(h, h->stream, kKindOther);
    if (!h->heatmap_exact && fast_smem <= 200 * 1024) {
        if (!h->heat_attr_set) {
            AIG_CK(cudaFuncSetAttribute(heatmap_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            AIG_CK(cudaFuncSetAttribute(heatmap_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            AIG_CK(cudaFuncSetAttribute(heatmap_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            // the kernels use no L1-cached global loads worth the space: give all of it to shared memory so 4 CTAs fit
            cudaFuncSetAttribute(heatmap_fast_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            cudaFuncSetAttribute(heatmap_fast_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            cudaFuncSetAttribute(heatmap_fast_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            h->heat_attr_set = true;
        }
        // 228 KiB per SM; each CTA also holds ~7.3 KiB of static shared memory and 1 KiB of system reserve
        const int per_sm = static_cast<int>(std::max<size_t>(1, std::min<size_t>(6, (228 * 1024) / (fast_smem + 9 * 1024))));
        const int grid = frames_grid(h, n_frames, per_sm);
        const bool aligned = (reinterpret_cast<uintptr_t>(d_heat) & 15u) == 0;
        if (aligned && out_w % 4 == 0)
            heatmap_fast_kernel<4><<<grid, kHeatThreads, fast_smem, h->stream>>>(d_energy, n_frames, out_h, out_w, d_heat);
        else if (aligned && out_w % 2 == 0)
            heatmap_fast_kernel<2><<<grid, kHeatThreads, fast_smem, h->stream>>>(d_energy, n_frames, out_h, out_w, d_heat);
        else
            heatmap_fast_kernel<1><<<grid, kHeatThreads, fast_smem, h->stream>>>(d_energy, n_frames, out_h, out_w, d_heat);
    } else {
        const size_t smem = static_cast<size_t>(out_w + out_h) * (sizeof(double) + sizeof(int));
        if (!h->heat_exact_attr_set) {      // up to 48 KiB of taps at 2048 x 2048, on top of 14 KiB of static shared memory
            AIG_CK(cudaFuncSetAttribute(heatmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            h->heat_exact_attr_set = true;
        }
        heatmap_kernel<<<frames_grid(h, n_frames, 4), kHeatThreads, smem, h->stream>>>(d_energy, n_frames, out_h, out_w, d_heat);
    }
    return scope.done("heatmap_kernel");
}

// ---- fused MFCC + energy kernel ---------------------------------------------------------------------
struct FusedVariant { int slabs, stages; };
constexpr int kNumFusedVariants = 3;
constexpr FusedVariant kFusedVariants[kNumFusedVariants] = {
    {1, 8},    // 0: 24 KiB stages x 8  = 192 KiB ring
    {2, 4},    // 1: 48 KiB stages x 4
    {4, 2},    // 2: 96 KiB stages x 2
};

template <int V, bool JITTER>
int launch_fused_variant(aig_handle* h, const CUtensorMap& map, float* d_mfcc, unsigned n_frames, int flip180,
                         int normalize_first, double* d_energy, uint8_t* d_mask, double* d_mean) {
    constexpr FusedVariant v = kFusedVariants[V];
    using F = FusedPipe<v.slabs, v.stages>;
    auto kernel = mfcc_energy_fused_kernel<v.slabs, v.stages, JITTER>;
    if (!h->fused_attr_set[V + (JITTER ? 4 : 0)]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F::kSmemBytes));
        h->fused_attr_set[V + (JITTER ? 4 : 0)] = true;
    }
    const unsigned grid = std::min<unsigned>(n_frames, static_cast<unsigned>(h->sm_count));
    LaunchScope scope(h, h->stream, kKindMfcc);
    kernel<<<grid, kFusedThreads, F::kSmemBytes, h->stream>>>(map, d_mfcc, n_frames, flip180, normalize_first, d_energy,
                                                             d_mask, d_mean, h->l2_evict_first, h->keep_mfcc_in_l2,
                                                             h->debug_jitter, FusedHeatOut{nullptr, 0, 0});
    return scope.done("mfcc_energy_fused_kernel");
}

// The opt-in persistent kernel with the heat-map phase: ring of 4 x 24 KiB stages (the heat phase needs 92 KiB at 224 x 298).
constexpr int kHeatFusedSlabs = 1, kHeatFusedStages = 4;
bool fused_heat_ok(const aig_handle* h, int out_h, int out_w, const float* d_heat) {
    using F = FusedPipe<kHeatFusedSlabs, kHeatFusedStages>;
    return !h->heatmap_exact && out_w % 2 == 0 && (static_cast<long long>(out_h) * out_w) % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(d_heat) & 15u) == 0 && F::smem_with_heat(out_h, out_w) <= 227 * 1024;
}
template <int VEC, int W, int H>
int launch_fused_heat_variant(aig_handle* h, const CUtensorMap& map, float* d_mfcc, unsigned n_frames, int flip180,
                              int normalize_first, double* d_energy, uint8_t* d_mask, double* d_mean, const FusedHeatOut& heat,
                              int slot) {
    using F = FusedPipe<kHeatFusedSlabs, kHeatFusedStages>;
    auto kernel = mfcc_energy_fused_kernel<kHeatFusedSlabs, kHeatFusedStages, false, VEC, W, H>;
    if (!h->fused_heat_attr_set[slot]) {
        AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        h->fused_heat_attr_set[slot] = true;
    }
    const size_t smem = F::smem_with_heat(heat.out_h, heat.out_w);
    const unsigned grid = std::min<unsigned>(n_frames, static_cast<unsigned>(h->sm_count));
    LaunchScope scope(h, h->stream, kKindMfcc);
    kernel<<<grid, kFusedThreads, smem, h->stream>>>(map, d_mfcc, n_frames, flip180, normalize_first, d_energy, d_mask, d_mean,
                                                    h->l2_evict_first, h->keep_mfcc_in_l2, 0u, heat);
    return scope.done("mfcc_energy_fused_kernel<heat>");
}
int launch_fused_heat(aig_handle* h, const float* d_power, int64_t n_frames, float* d_mfcc, int flip180, int normalize_first,
                      double* d_energy, uint8_t* d_mask, double* d_mean, float* d_heat, int out_h, int out_w) {
    if ((reinterpret_cast<uintptr_t>(d_power) & 15u) || (reinterpret_cast<uintptr_t>(d_mfcc) & 15u))
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc_energy_heatmap: device buffers must be 16-byte aligned");
    const int64_t max_frames = std::max<int64_t>(1, h->launch_row_limit / kFramePixels);
    for (int64_t done = 0; done < n_frames; done += max_frames) {
        const int64_t frames = std::min(max_frames, n_frames - done);
        CUtensorMap map;
        int rc = encode_spectrum_map(h, d_power + done * kFramePixels * kFftLen, static_cast<uint64_t>(frames) * kFramePixels,
                                     kFusedRows, &map);
        if (rc != AIG_OK) return rc;
        float* mf = d_mfcc + done * kFrameValues;
        double* en = d_energy ? d_energy + done * kFramePixels : nullptr;
        uint8_t* mk = d_mask ? d_mask + done * kFramePixels : nullptr;
        double* mn = d_mean ? d_mean + done : nullptr;
        const FusedHeatOut heat = {d_heat + done * out_h * out_w, out_h, out_w};
        const unsigned f = static_cast<unsigned>(frames);
        if (out_h == 224 && out_w == 298) rc = launch_fused_heat_variant<2, 298, 224>(h, map, mf, f, flip180, normalize_first, en, mk, mn, heat, 0);
        else if (out_h == 224 && out_w == 224) rc = launch_fused_heat_variant<4, 224, 224>(h, map, mf, f, flip180, normalize_first, en, mk, mn, heat, 1);
        else rc = launch_fused_heat_variant<2, 0, 0>(h, map, mf, f, flip180, normalize_first, en, mk, mn, heat, 2);
        if (rc != AIG_OK) return rc;
    }
    return AIG_OK;
}

int launch_fused(aig_handle* h, const float* d_power, int64_t n_frames, float* d_mfcc, int flip180, int normalize_first,
                 double* d_energy, uint8_t* d_mask, double* d_mean) {
    if ((reinterpret_cast<uintptr_t>(d_power) & 15u) || (reinterpret_cast<uintptr_t>(d_mfcc) & 15u))
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc_energy: device buffers must be 16-byte aligned");
    const int64_t max_frames = std::max<int64_t>(1, h->launch_row_limit / kFramePixels);
    for (int64_t done = 0; done < n_frames; done += max_frames) {
        const int64_t frames = std::min(max_frames, n_frames - done);
        CUtensorMap map;
        int rc = encode_spectrum_map(h, d_power + done * kFramePixels * kFftLen, static_cast<uint64_t>(frames) * kFramePixels,
                                     kFusedRows, &map);
        if (rc != AIG_OK) return rc;
        float* mf = d_mfcc + done * kFrameValues;
        double* en = d_energy ? d_energy + done * kFramePixels : nullptr;
        uint8_t* mk = d_mask ? d_mask + done * kFramePixels : nullptr;
        double* mn = d_mean ? d_mean + done : nullptr;
        const unsigned f = static_cast<unsigned>(frames);
        if (h->debug_jitter != 0) {              // the jittered build exists for the default ring geometry only
            rc = launch_fused_variant<2, true>(h, map, mf, f, flip180, normalize_first, en, mk, mn);
        } else {
            switch (h->fused_variant) {
                case 0: rc = launch_fused_variant<0, false>(h, map, mf, f, flip180, normalize_first, en, mk, mn); break;
                case 1: rc = launch_fused_variant<1, false>(h, map, mf, f, flip180, normalize_first, en, mk, mn); break;
                case 2: rc = launch_fused_variant<2, false>(h, map, mf, f, flip180, normalize_first, en, mk, mn); break;
                default: rc = h->fail(AIG_ERR_ARGUMENT, "unknown fused kernel variant %d", h->fused_variant);
            }
        }
        if (rc != AIG_OK) return rc;
    }
    return AIG_OK;
}

int require(aig_handle* h) {
    if (h == nullptr) return AIG_ERR_ARGUMENT;
    h->err.clear();
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return h->fail_cuda(e, "cudaSetDevice");
    return scratch_reset(h);
}

// Host spectra -> device in frame-aligned chunks, double-buffered on the copy stream so the H2D
// copy of chunk i+1 overlaps the kernels of chunk i.  `consume(chunk device ptr, first row, rows)`
// enqueues the work of one chunk on h->stream.
template <typename Consume>
int stream_host_rows(aig_handle* h, const float* host, int64_t n_rows, int row_floats, int64_t rows_per_chunk,
                     Consume consume) {
    const size_t chunk_bytes = static_cast<size_t>(rows_per_chunk) * row_floats * sizeof(float);
    float* buf[2] = {static_cast<float*>(scratch(h, chunk_bytes)), static_cast<float*>(scratch(h, chunk_bytes))};
    if (!buf[0] || !buf[1]) return AIG_ERR_ALLOC;
    const MemKind kind = classify(host);
    // the scratch may still be in use by earlier work on h->stream
    AIG_CK(cudaEventRecord(h->ev_consumed[0], h->stream));
    AIG_CK(cudaEventRecord(h->ev_consumed[1], h->stream));
    int slot = 0;
    for (int64_t row = 0; row < n_rows; row += rows_per_chunk, slot ^= 1) {
        const int64_t rows = std::min(rows_per_chunk, n_rows - row);
        AIG_CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0));
        AIG_CK(upload_async(h, buf[slot], host + row * row_floats, static_cast<size_t>(rows) * row_floats * sizeof(float),
                            kind, h->copy_stream));
        AIG_CK(cudaEventRecord(h->ev_copied[slot], h->copy_stream));
        AIG_CK(cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
        int rc = consume(buf[slot], row, rows);
        if (rc != AIG_OK) return rc;
        AIG_CK(cudaEventRecord(h->ev_consumed[slot], h->stream));
    }
    return AIG_OK;
}

// Rows per H2D chunk of a host-input call: ~64 MiB for long calls (a copy engine is at the link rate from 4 MiB up and
// fewer chunks mean fewer launches); a pinned call is cut into at least six chunks down to 8 MiB each, so that a
// 16-frame call (BASELINE configs[0], 56.6 MB) still overlaps its copies with its kernels and result copies.
int64_t host_chunk_rows(int64_t unit_rows, int row_floats, int64_t total_rows, MemKind kind) {
    const int64_t unit_bytes = unit_rows * row_floats * 4;
    const int64_t total_bytes = total_rows * row_floats * 4;
    int64_t target = int64_t(64) << 20;
    // pageable sources already stream through the staging ring in 1-4 MiB pieces: cutting the call further only adds
    // start-up cost per chunk (measured: 16 pageable frames 4.0 ms in six chunks, 1.8 ms in one)
    if (kind == kHostPinned) target = std::min(target, std::max<int64_t>(int64_t(8) << 20, total_bytes / 6));
    return std::max<int64_t>(1, target / unit_bytes) * unit_rows;
}

}  // namespace

// =====================================================================================================
extern "C" {

int aig_abi_version(void) { return AIG_ABI_VERSION; }

// The marker prefix lets the Python loader read the stamp from the file without dlopen (see _lib.binary_build_id).
const char* aig_build_id(void) {
    static const char stamp[] = "AIG_BUILD_ID=" AIG_BUILD_ID_STRING;
    return stamp + 13;
}

const char* aig_last_error(const aig_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int aig_create(int device, uint64_t stream, aig_handle** out) {
    if (out == nullptr) { g_create_error = "aig_create: out is null"; return AIG_ERR_ARGUMENT; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_create_error = "aig_create: no CUDA device available (this library has no CPU fallback)";
        return AIG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        g_create_error = "aig_create: device index out of range";
        return AIG_ERR_ARGUMENT;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        g_create_error = "aig_create: cannot select the device";
        return AIG_ERR_NO_DEVICE;
    }
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "aig_create: device %d is sm_%d%d; libaig is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
        g_create_error = buf;
        return AIG_ERR_NO_DEVICE;
    }
    aig_handle* h = new aig_handle();
    h->device = device;
    h->stream = reinterpret_cast<cudaStream_t>(stream);
    h->sm_count = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr) {
        cudaGetLastError();
        g_create_error = "aig_create: the driver does not export cuTensorMapEncodeTiled";
        delete h;
        return AIG_ERR_NO_DEVICE;
    }
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
    bool ok = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&h->ev_chain[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        g_create_error = "aig_create: stream / event creation failed";
        aig_destroy(h);
        return AIG_ERR_ALLOC;
    }
    *out = h;
    return AIG_OK;
}

int aig_destroy(aig_handle* h) {
    if (h == nullptr) return AIG_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->uploader.shutdown();
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    for (int i = 0; i < 4; ++i) if (h->ev_chain[i]) cudaEventDestroy(h->ev_chain[i]);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.start); cudaEventDestroy(sp.stop); }
    for (auto& e : h->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
        if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
    }
    if (h->comm != nullptr) { nccl().comm_destroy(h->comm); h->comm = nullptr; }
    for (auto& b : h->overflow) cudaFree(b.first);
    if (h->arena) cudaFree(h->arena);
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_twiddle) cudaFree(h->d_twiddle);
    if (h->pinned) cudaFreeHost(h->pinned);
    cudaGetLastError();
    delete h;
    return AIG_OK;
}

int aig_synchronize(aig_handle* h) {
    if (h == nullptr) return AIG_ERR_ARGUMENT;
    AIG_CK(cudaSetDevice(h->device));
    AIG_CK(cudaStreamSynchronize(h->copy_stream));
    AIG_CK(cudaStreamSynchronize(h->stream));
    AIG_CK(cudaStreamSynchronize(h->aux_stream));
    return AIG_OK;
}

int64_t aig_launch_count(const aig_handle* h) { return h ? h->launches : -1; }

int aig_set_option(aig_handle* h, const char* name, int64_t value) {
    if (h == nullptr || name == nullptr) return AIG_ERR_ARGUMENT;
    const std::string key(name);
    if (key == "mfcc_variant") {
        if (value < -1 || value >= kNumVariants) return h->fail(AIG_ERR_ARGUMENT, "unknown MFCC kernel variant %lld", (long long)value);
        h->variant = static_cast<int>(value);
    } else if (key == "chain_chunk_frames") {
        if (value < 1 || value > (1 << 20)) return h->fail(AIG_ERR_ARGUMENT, "chain_chunk_frames out of range");
        h->chain_chunk_frames = static_cast<int>(value);
    } else if (key == "chain_overlap") {
        h->chain_overlap = value != 0;
    } else if (key == "chain_mode") {
        if (value < 0 || value > 2) return h->fail(AIG_ERR_ARGUMENT, "chain_mode must be 0, 1 or 2");
        h->chain_mode = static_cast<int>(value);
        if (value < 2) h->chain_overlap = value == 1;
    } else if (key == "fused_variant") {
        if (value < 0 || value >= kNumFusedVariants) return h->fail(AIG_ERR_ARGUMENT, "unknown fused kernel variant %lld", (long long)value);
        h->fused_variant = static_cast<int>(value);
    } else if (key == "host_copy_threads") {
        if (value < -1 || value > 64) return h->fail(AIG_ERR_ARGUMENT, "host_copy_threads out of range (-1 auto, 0 off, 1..64)");
        if (static_cast<int>(value) != h->host_copy_threads) h->uploader.shutdown();
        h->host_copy_threads = static_cast<int>(value);
    } else if (key == "staged_min_bytes") {
        if (value < (64 << 10) || value > (int64_t(1) << 40)) return h->fail(AIG_ERR_ARGUMENT, "staged_min_bytes out of range (from 65536)");
        h->staged_min_bytes = static_cast<size_t>(value);
    } else if (key == "staged_solo_bytes") {
        if (value < 0 || value > (int64_t(1) << 40)) return h->fail(AIG_ERR_ARGUMENT, "staged_solo_bytes out of range");
        h->uploader.set_solo_bytes(static_cast<size_t>(value));
    } else if (key == "staged_small_piece_bytes") {
        if (value < (64 << 10) || value > (4 << 20)) return h->fail(AIG_ERR_ARGUMENT, "staged_small_piece_bytes out of range (64 KiB .. 4 MiB)");
        h->uploader.set_small_piece(static_cast<size_t>(value));
    } else if (key == "host_copy_streaming") {
        if (value < -1 || value > 1) return h->fail(AIG_ERR_ARGUMENT, "host_copy_streaming must be -1 (by job size), 0 or 1");
        h->uploader.set_streaming_fill(static_cast<int>(value));   // staging-slot fills with non-temporal stores (host_copy.cpp)
    } else if (key == "chain_energy_ctas_per_sm") {
        if (value < 1 || value > 16) return h->fail(AIG_ERR_ARGUMENT, "chain_energy_ctas_per_sm out of range");
        h->chain_energy_ctas_per_sm = static_cast<int>(value);
    } else if (key == "launch_row_limit") {
        if (value < 1 || value > (int64_t(1) << 31) - 1024) return h->fail(AIG_ERR_ARGUMENT, "launch_row_limit out of range");
        h->launch_row_limit = value;
    } else if (key == "heatmap_exact") {
        h->heatmap_exact = value != 0;
    } else if (key == "keep_mfcc_in_l2") {
        h->keep_mfcc_in_l2 = value != 0;
    } else if (key == "l2_evict_first") {
        h->l2_evict_first = value != 0;
    } else if (key == "small_host_bytes") {
        if (value < 0 || value > static_cast<int64_t>(kPinnedArenaBytes)) return h->fail(AIG_ERR_ARGUMENT, "small_host_bytes out of range");
        h->small_host_bytes = static_cast<int>(value);
    } else if (key == "heat_bulk_store") {
        h->heat_bulk_store = value != 0;
    } else if (key == "energy_wide") {
        h->energy_wide = value != 0;
    } else if (key == "acivw_wide_pairs") {
        if (value < 0) return h->fail(AIG_ERR_ARGUMENT, "acivw_wide_pairs out of range");
        h->acivw_wide_pairs = value;
    } else if (key == "overlay_luma") {
        h->overlay_luma = value != 0;
    } else if (key == "energy_heat_ws") {
        h->energy_heat_ws = value != 0;
    } else if (key == "mask_packed") {
        h->mask_packed = value != 0;
    } else if (key == "norm_bulk_copy") {
        h->norm_bulk_copy = value != 0;

    } else if (key == "small_batch_frames") {
        if (value < 0 || value > (1 << 20)) return h->fail(AIG_ERR_ARGUMENT, "small_batch_frames out of range");
        h->small_batch_frames = static_cast<int>(value);
    } else if (key == "debug_jitter") {
        if (value < 0 || value > 0x7fffffff) return h->fail(AIG_ERR_ARGUMENT, "debug_jitter: seed out of range");
        h->debug_jitter = static_cast<unsigned int>(value);
    } else if (key == "profile") {
        h->profile = value != 0;
    } else {
        return h->fail(AIG_ERR_ARGUMENT, "unknown option '%s'", name);
    }
    return AIG_OK;
}

int aig_selftest(aig_handle* h, int which, uint64_t* out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (out == nullptr || which < 0 || which > 2) return h->fail(AIG_ERR_ARGUMENT, "aig_selftest: bad arguments");
    unsigned long long* d_out = static_cast<unsigned long long*>(scratch(h, 4 * sizeof(unsigned long long)));
    if (!d_out) return AIG_ERR_ALLOC;
    AIG_CK(cudaMemsetAsync(d_out, 0, 4 * sizeof(unsigned long long), h->stream));
    LaunchScope scope(h, h->stream, kKindOther);
    if (which == 0)
        selftest_division_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(d_out);
    else if (which == 1)
        selftest_exp_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(1ull << 26, d_out);
    else
        selftest_norm_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(d_out);
    rc = scope.done("selftest kernel");
    if (rc != AIG_OK) return rc;
    AIG_CK(cudaMemcpyAsync(out, d_out, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    AIG_CK(cudaStreamSynchronize(h->stream));
    return AIG_OK;
}

int aig_profile_read(aig_handle* h, double* ms_out, int64_t* launches_out) {
    if (h == nullptr || ms_out == nullptr || launches_out == nullptr) return AIG_ERR_ARGUMENT;
    int rc = aig_synchronize(h);
    if (rc != AIG_OK) return rc;
    for (int k = 0; k < 3; ++k) { ms_out[k] = 0.0; launches_out[k] = 0; }
    for (auto& sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess) {
            ms_out[sp.kind] += ms;
            launches_out[sp.kind] += 1;
        } else {
            cudaGetLastError();
        }
        h->event_pool.push_back(sp.start);
        h->event_pool.push_back(sp.stop);
    }
    h->spans.clear();
    return AIG_OK;
}

int aig_set_tables(aig_handle* h, const double* filter_mat, int fft_len, int filter_num, const double* dct_base,
                   int mfcc_num, const double* lifter, double mfnorm) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (!filter_mat || !dct_base || !lifter) return h->fail(AIG_ERR_ARGUMENT, "aig_set_tables: null table");
    if (fft_len < 1 || fft_len > 4096 || filter_num < 1 || filter_num > kGenericMaxFilters || mfcc_num < 1 ||
        mfcc_num > 32)
        return h->fail(AIG_ERR_TABLES, "aig_set_tables: unsupported geometry fft_len=%d filter_num=%d mfcc_num=%d",
                       fft_len, filter_num, mfcc_num);
    // Is this exactly the reference configuration the banded mel program was generated from?
    bool ref = fft_len == AIG_REF_FFT_LEN && filter_num == AIG_REF_FILTER_NUM && mfcc_num == AIG_REF_MFCC_NUM &&
               mfnorm == kRefMfnorm;
    if (ref) {
        ref = std::memcmp(dct_base, kRefDct, sizeof kRefDct) == 0 && std::memcmp(lifter, kRefLifter, sizeof kRefLifter) == 0;
    }
    if (ref) {
        std::vector<double> dense(static_cast<size_t>(AIG_REF_FFT_LEN) * AIG_REF_FILTER_NUM, 0.0);
        for (int i = 0; i < AIG_REF_NNZ; ++i) dense[kRefNzBin[i] * AIG_REF_FILTER_NUM + kRefNzCol[i]] = kRefNzVal[i];
        for (size_t i = 0; i < dense.size() && ref; ++i) ref = dense[i] == filter_mat[i];
    }
    const size_t n_bank = static_cast<size_t>(fft_len) * filter_num, n_dct = static_cast<size_t>(filter_num) * mfcc_num;
    AIG_CK(cudaStreamSynchronize(h->stream));
    if (h->d_tables) { cudaFree(h->d_tables); h->d_tables = nullptr; }
    if (cudaMalloc(&h->d_tables, (n_bank + n_dct + mfcc_num) * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return h->fail(AIG_ERR_ALLOC, "aig_set_tables: cudaMalloc failed");
    }
    AIG_CK(cudaMemcpy(h->d_tables, filter_mat, n_bank * sizeof(double), cudaMemcpyHostToDevice));
    AIG_CK(cudaMemcpy(h->d_tables + n_bank, dct_base, n_dct * sizeof(double), cudaMemcpyHostToDevice));
    AIG_CK(cudaMemcpy(h->d_tables + n_bank + n_dct, lifter, mfcc_num * sizeof(double), cudaMemcpyHostToDevice));
    h->fft_len = fft_len; h->filter_num = filter_num; h->mfcc_num = mfcc_num; h->mfnorm = mfnorm;
    h->tables_set = true;
    h->tables_ref = ref;
    return AIG_OK;
}

int aig_tables_are_reference(const aig_handle* h) {
    if (h == nullptr) return AIG_ERR_ARGUMENT;
    if (!h->tables_set) return AIG_ERR_TABLES;
    return h->tables_ref ? 1 : 0;
}

int aig_mfcc(aig_handle* h, const float* power, int64_t n_rows, float* mfcc_out, int flip180, int frame_pixels) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (!h->tables_set) return h->fail(AIG_ERR_TABLES, "aig_mfcc: call aig_set_tables first");
    if (n_rows < 0 || (n_rows > 0 && (!power || !mfcc_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc: bad buffers");
    if (flip180 && (frame_pixels < 1 || n_rows % frame_pixels != 0))
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc: flip180 needs n_rows (%lld) to be a multiple of frame_pixels (%d)",
                       (long long)n_rows, frame_pixels);
    if (!flip180 && frame_pixels < 1) frame_pixels = 1;
    if (n_rows == 0) return AIG_OK;
    const int in_floats = h->fft_len, out_floats = h->mfcc_num;
    const bool in_dev = classify(power) == kDevice, out_dev = classify(mfcc_out) == kDevice;
    float* d_out = out_dev ? mfcc_out : static_cast<float*>(scratch(h, static_cast<size_t>(n_rows) * out_floats * 4));
    if (!d_out) return AIG_ERR_ALLOC;
    if (in_dev) {
        rc = launch_mfcc(h, power, n_rows, d_out, flip180, frame_pixels);
        if (rc != AIG_OK) return rc;
    } else {
        const int64_t chunk = host_chunk_rows(flip180 ? frame_pixels : 128, in_floats, n_rows, classify(power));
        rc = stream_host_rows(h, power, n_rows, in_floats, chunk, [&](const float* d_chunk, int64_t row, int64_t rows) {
            return launch_mfcc(h, d_chunk, rows, d_out + row * out_floats, flip180, frame_pixels);
        });
        if (rc != AIG_OK) return rc;
    }
    if (!out_dev)
        AIG_CK(download(h, mfcc_out, d_out, static_cast<size_t>(n_rows) * out_floats * 4, classify(mfcc_out), h->stream));
    if (!out_dev || !in_dev) AIG_CK(cudaStreamSynchronize(h->stream));
    return AIG_OK;
}

int aig_normalize_images(aig_handle* h, const float* images, int64_t n_frames, float* out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!images || !out))) return h->fail(AIG_ERR_ARGUMENT, "aig_normalize_images: bad buffers");
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t count = static_cast<size_t>(n_frames) * kFrameValues;
    const float* d_in = io.in(images, count);
    float* d_out = io.out(out, count);
    if (io.failed) return io.finish();
    if (((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) != 0)
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_normalize_images: device buffers must be 16-byte aligned"));
    const bool bulk = h->norm_bulk_copy != 0;
    if (bulk && !h->norm_bulk_attr_set) {
        AIG_CK(cudaFuncSetAttribute(normalize_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kNormBulkSmem)));
        h->norm_bulk_attr_set = true;
    }
    LaunchScope scope(h, h->stream, kKindOther);
    if (bulk)
        normalize_bulk_kernel<<<frames_grid(h, n_frames, 1), kNormBulkThreads, kNormBulkSmem, h->stream>>>(d_in, n_frames, d_out);
    else
        normalize_kernel<<<frames_grid(h, n_frames, 8), 256, 0, h->stream>>>(d_in, n_frames, d_out);
    rc = scope.done("normalize_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_energy(aig_handle* h, const float* images, int64_t n_frames, int normalize_first, float* scaled_out,
               double* energy_out, uint8_t* mask_out, double* mean_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && !images)) return h->fail(AIG_ERR_ARGUMENT, "aig_energy: bad buffers");
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    io.zero_copy_ok = n_frames <= 1;             // the find_logen drop-in: a frame in, a frame and a map out, latency-bound
    const size_t n = static_cast<size_t>(n_frames);
    const float* d_in = io.in(images, n * kFrameValues);
    float* d_scaled = io.out(scaled_out, n * kFrameValues);
    double* d_energy = io.out(energy_out, n * kFramePixels);
    uint8_t* d_mask = io.out(mask_out, n * kFramePixels);
    double* d_mean = io.out(mean_out, n);
    if (io.failed) return io.finish();
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (reinterpret_cast<uintptr_t>(d_scaled) & 7u))
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_energy: device image buffers must be 16-byte aligned"));
    rc = launch_energy(h, h->stream, d_in, n_frames, normalize_first, d_scaled, d_energy, d_mask, d_mean);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_heatmap(aig_handle* h, const double* energy, int64_t n_frames, int out_h, int out_w, float* heat_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!energy || !heat_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_heatmap: bad buffers");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut)
        return h->fail(AIG_ERR_ARGUMENT, "aig_heatmap: output size %dx%d outside 1..%d", out_h, out_w, kMaxOut);
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_frames);
    const double* d_energy = io.in(energy, n * kFramePixels);
    float* d_heat = io.out(heat_out, n * out_h * out_w);
    if (io.failed) return io.finish();
    rc = launch_heatmap(h, d_energy, n_frames, out_h, out_w, d_heat);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_energy_heatmap(aig_handle* h, const float* images, int64_t n_frames, int normalize_first, double* energy_out,
                       uint8_t* mask_out, float* heat_out, int out_h, int out_w) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!images || !heat_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_energy_heatmap: bad buffers");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut)
        return h->fail(AIG_ERR_ARGUMENT, "aig_energy_heatmap: output size %dx%d outside 1..%d", out_h, out_w, kMaxOut);
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_frames);
    const float* d_in = io.in(images, n * kFrameValues);
    double* d_energy = io.out(energy_out, n * kFramePixels);
    uint8_t* d_mask = io.out(mask_out, n * kFramePixels);
    float* d_heat = io.out(heat_out, n * out_h * out_w);
    if (io.failed) return io.finish();
    // One launch: the energy map goes from the float64 warps to the up-sampling passes through shared memory.  Small
    // batches (fewer frames than SMs) and shapes the streaming kernel cannot take run the two kernels back to back.
    if (heat_stream_ok(h, out_h, out_w, d_heat, true) && n_frames >= small_batch_limit(h)) {
        HeatStreamArgs args = {};
        args.s2.img[0] = d_in; args.s2.energy[0] = d_energy; args.s2.mask[0] = d_mask;
        args.s2.n_frames = n_frames; args.s2.normalize_first = normalize_first;
        args.n_frames = n_frames; args.out_h = out_h; args.out_w = out_w; args.heat = d_heat;
        rc = energy_heat_ws_ok(h, out_h, out_w, d_heat) ? launch_energy_heat_ws<WsTwin>(h, args, 0) : launch_heat_stream<true>(h, args);
        if (rc != AIG_OK) return io.abort(rc);
        return io.finish();
    }
    if (d_energy == nullptr) d_energy = static_cast<double*>(scratch(h, n * kFramePixels * sizeof(double)));
    if (d_energy == nullptr) return io.abort(AIG_ERR_ALLOC);
    rc = launch_energy(h, h->stream, d_in, n_frames, normalize_first, nullptr, d_energy, d_mask, nullptr);
    if (rc != AIG_OK) return io.abort(rc);
    rc = launch_heatmap(h, d_energy, n_frames, out_h, out_w, d_heat);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_acivw_batch(aig_handle* h, const float* real, const float* reconstructed, int64_t n_frames, int normalize_first,
                    const double* thr, int k, int64_t* inter_out, int64_t* union_out, int64_t* pos_inout,
                    int64_t* num_inout, double* energy_real_out, double* energy_recon_out, uint8_t* mask_real_out,
                    uint8_t* mask_recon_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || k < 0 || k > kMaxThresholds)
        return h->fail(AIG_ERR_ARGUMENT, "aig_acivw_batch: n=%lld k=%d out of range (k <= %d)", (long long)n_frames, k, kMaxThresholds);
    if ((k > 0 && (!thr || !pos_inout)) || !num_inout) return h->fail(AIG_ERR_ARGUMENT, "aig_acivw_batch: null threshold / count buffers");
    if (n_frames > 0 && (!real || !reconstructed)) return h->fail(AIG_ERR_ARGUMENT, "aig_acivw_batch: null images");
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_frames);
    Stage2Args args = {};
    args.img[0] = io.in(real, n * kFrameValues);
    args.img[1] = io.in(reconstructed, n * kFrameValues);
    args.energy[0] = io.out(energy_real_out, n * kFramePixels);
    args.energy[1] = io.out(energy_recon_out, n * kFramePixels);
    args.mask[0] = io.out(mask_real_out, n * kFramePixels);
    args.mask[1] = io.out(mask_recon_out, n * kFramePixels);
    args.thr = io.in(thr, static_cast<size_t>(k));
    args.k_thr = k;
    args.inter = reinterpret_cast<long long*>(io.out(inter_out, n));
    args.uni = reinterpret_cast<long long*>(io.out(union_out, n));
    args.pos = reinterpret_cast<unsigned long long*>(io.inout(pos_inout, static_cast<size_t>(k)));
    args.num = reinterpret_cast<unsigned long long*>(io.inout(num_inout, 1));
    args.n_frames = n_frames;
    args.normalize_first = normalize_first;
    if (io.failed) return io.finish();
    if ((reinterpret_cast<uintptr_t>(args.img[0]) & 15u) || (reinterpret_cast<uintptr_t>(args.img[1]) & 15u))
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_acivw_batch: image buffers must be 16-byte aligned"));
    rc = launch_stage2<2>(h, h->stream, args);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_resize_mask(aig_handle* h, const uint8_t* mask, int64_t n_frames, int out_h, int out_w, uint8_t* mask_up) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!mask || !mask_up))) return h->fail(AIG_ERR_ARGUMENT, "aig_resize_mask: bad buffers");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut)
        return h->fail(AIG_ERR_ARGUMENT, "aig_resize_mask: output size %dx%d outside 1..%d", out_h, out_w, kMaxOut);
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_frames);
    const uint8_t* d_mask = io.in(mask, n * kFramePixels);
    uint8_t* d_up = io.out(mask_up, n * out_h * out_w);
    if (io.failed) return io.finish();
    if (packed_size(h, out_h, out_w) && (reinterpret_cast<uintptr_t>(d_up) & 15u) == 0) {
        rc = out_w == 298 ? launch_resize_packed<298, 224>(h, d_mask, n_frames, d_up, 0) : launch_resize_packed<224, 224>(h, d_mask, n_frames, d_up, 1);
        if (rc != AIG_OK) return io.abort(rc);
        return io.finish();
    }
    const size_t smem = MaskTaps::bytes(out_h, out_w);
    rc = allow_mask_smem(h);
    if (rc != AIG_OK) return io.abort(rc);
    LaunchScope scope(h, h->stream, kKindOther);
    resize_mask_kernel<<<frames_grid(h, n_frames, mask_ctas_per_sm(smem)), kHeatThreads, smem, h->stream>>>(d_mask, n_frames, out_h, out_w, d_up);
    rc = scope.done("resize_mask_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_mfcc_energy(aig_handle* h, const float* power, int64_t n_frames, int flip180, int normalize_first,
                    float* mfcc_out, double* energy_out, uint8_t* mask_out, double* mean_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (!h->tables_set || !h->tables_ref)
        return h->fail(AIG_ERR_TABLES, "aig_mfcc_energy: needs the reference tables (aig_set_tables)");
    if (n_frames < 0 || (n_frames > 0 && (!power || !mfcc_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc_energy: bad buffers");
    if (n_frames == 0) return AIG_OK;
    const size_t n = static_cast<size_t>(n_frames);
    const bool in_dev = classify(power) == kDevice;
    Io io(h);
    float* d_mfcc = io.out(mfcc_out, n * kFrameValues);
    double* d_energy = io.out(energy_out, n * kFramePixels);
    uint8_t* d_mask = io.out(mask_out, n * kFramePixels);
    double* d_mean = io.out(mean_out, n);
    if (io.failed) return io.finish();
    // The energy kernel is FP64-compute-bound and touches 2 % of the bytes; the MFCC kernel is HBM-bound
    // and leaves the FP64 pipe idle.  Run them chunk-wise on two streams so that the energy kernel of
    // chunk i executes underneath the MFCC kernel of chunk i+1 (its input is still in L2).
    const bool fused = h->chain_mode == 2;
    const bool overlap = h->chain_overlap && !fused;
    cudaStream_t energy_stream = overlap ? h->aux_stream : h->stream;
    if (overlap) {
        AIG_CK(cudaEventRecord(h->ev_chain[0], h->stream));              // order after earlier work
        AIG_CK(cudaStreamWaitEvent(h->aux_stream, h->ev_chain[0], 0));
    }
    int64_t chunk_index = 0;
    // The persistent kernel gives one CTA a whole frame; with fewer frames than SMs the tiled MFCC kernel followed by the
    // cluster-per-frame energy kernel uses the whole GPU instead (same results bit for bit).
    auto run = [&](const float* d_power, int64_t frame0, int64_t frames) -> int {
        if (fused && (frames >= small_batch_limit(h) || h->debug_jitter != 0))
            return launch_fused(h, d_power, frames, d_mfcc + frame0 * kFrameValues, flip180, normalize_first,
                                d_energy ? d_energy + frame0 * kFramePixels : nullptr,
                                d_mask ? d_mask + frame0 * kFramePixels : nullptr, d_mean ? d_mean + frame0 : nullptr);
        for (int64_t f = 0; f < frames; f += h->chain_chunk_frames, ++chunk_index) {
            const int64_t cf = std::min<int64_t>(h->chain_chunk_frames, frames - f);
            const int64_t g = frame0 + f;
            float* mf = d_mfcc + g * kFrameValues;
            int r = launch_mfcc(h, d_power + f * kFramePixels * kFftLen, cf * kFramePixels, mf, flip180, kFramePixels);
            if (r != AIG_OK) return r;
            if (overlap) {
                cudaEvent_t ev = h->ev_chain[1 + (chunk_index & 1)];
                AIG_CK(cudaEventRecord(ev, h->stream));
                AIG_CK(cudaStreamWaitEvent(energy_stream, ev, 0));
            }
            r = launch_energy(h, energy_stream, mf, cf, normalize_first, nullptr,
                              d_energy ? d_energy + g * kFramePixels : nullptr,
                              d_mask ? d_mask + g * kFramePixels : nullptr, d_mean ? d_mean + g : nullptr,
                              overlap ? h->chain_energy_ctas_per_sm : 8);
            if (r != AIG_OK) return r;
        }
        return AIG_OK;
    };
    bool d2h_overlapped = false;
    if (in_dev) {
        rc = run(power, 0, n_frames);
    } else {
        io.any_host = true;
        // Host outputs of a host-input call: instead of one device-to-host copy at the very end, ship each chunk's
        // results on the (otherwise idle, in fused mode) second stream while the next chunk's spectra are still coming in -
        // PCIe is full duplex, so the result copies disappear behind the 40x larger input copies.
        struct HostOut { char* host; char* dev; size_t bytes_per_frame; };
        std::vector<HostOut> host_outs;
        if (fused) {
            const size_t per_frame[4] = {kFrameValues * sizeof(float), kFramePixels * sizeof(double), kFramePixels, sizeof(double)};
            void* const user[4] = {mfcc_out, energy_out, mask_out, mean_out};
            void* const devp[4] = {d_mfcc, d_energy, d_mask, d_mean};
            // Pinned destinations only: a copy into pageable memory would block this thread on the chunk's kernel and
            // stall the uploads behind it; those results go out through the staging ring at the end (io.finish()).
            for (int i = 0; i < 4; ++i)
                if (user[i] != nullptr && devp[i] != user[i] && classify(user[i]) == kHostPinned) {
                    host_outs.push_back({static_cast<char*>(user[i]), static_cast<char*>(devp[i]), per_frame[i]});
                    io.outs.erase(std::remove_if(io.outs.begin(), io.outs.end(),
                                                 [&](const Io::Pending& pd) { return pd.host == user[i]; }), io.outs.end());
                }
            d2h_overlapped = !host_outs.empty();
        }
        const int64_t chunk_rows = host_chunk_rows(kFramePixels, kFftLen, n_frames * kFramePixels, classify(power));
        rc = stream_host_rows(h, power, n_frames * kFramePixels, kFftLen, chunk_rows,
                              [&](const float* d_chunk, int64_t row, int64_t rows) {
                                  const int64_t f0 = row / kFramePixels, nf = rows / kFramePixels;
                                  int r = run(d_chunk, f0, nf);
                                  if (r != AIG_OK || host_outs.empty()) return r;
                                  AIG_CK(cudaEventRecord(h->ev_chain[1], h->stream));
                                  AIG_CK(cudaStreamWaitEvent(h->aux_stream, h->ev_chain[1], 0));
                                  for (auto& o : host_outs)
                                      AIG_CK(cudaMemcpyAsync(o.host + f0 * o.bytes_per_frame, o.dev + f0 * o.bytes_per_frame,
                                                             static_cast<size_t>(nf) * o.bytes_per_frame, cudaMemcpyDeviceToHost,
                                                             h->aux_stream));
                                  return AIG_OK;
                              });
    }
    if (d2h_overlapped) AIG_CK(cudaStreamSynchronize(h->aux_stream));
    if (rc != AIG_OK) return io.abort(rc);
    if (overlap) {
        AIG_CK(cudaEventRecord(h->ev_chain[3], h->aux_stream));          // results visible to the handle's stream
        AIG_CK(cudaStreamWaitEvent(h->stream, h->ev_chain[3], 0));
    }
    return io.finish();
}

int aig_mfcc_energy_heatmap(aig_handle* h, const float* power, int64_t n_frames, int flip180, int normalize_first,
                            float* mfcc_out, double* energy_out, uint8_t* mask_out, double* mean_out, float* heat_out,
                            int out_h, int out_w) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (!h->tables_set || !h->tables_ref)
        return h->fail(AIG_ERR_TABLES, "aig_mfcc_energy_heatmap: needs the reference tables (aig_set_tables)");
    if (n_frames < 0 || (n_frames > 0 && (!power || !mfcc_out || !heat_out)))
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc_energy_heatmap: bad buffers");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut)
        return h->fail(AIG_ERR_ARGUMENT, "aig_mfcc_energy_heatmap: output size %dx%d outside 1..%d", out_h, out_w, kMaxOut);
    if (n_frames == 0) return AIG_OK;
    const size_t n = static_cast<size_t>(n_frames);
    Io io(h);
    const float* d_power = io.in(power, n * kFramePixels * kFftLen);
    float* d_mfcc = io.out(mfcc_out, n * kFrameValues);
    double* d_energy = io.out(energy_out, n * kFramePixels);
    uint8_t* d_mask = io.out(mask_out, n * kFramePixels);
    double* d_mean = io.out(mean_out, n);
    float* d_heat = io.out(heat_out, n * out_h * out_w);
    if (io.failed) return io.finish();
    if (fused_heat_ok(h, out_h, out_w, d_heat) && n_frames >= small_batch_limit(h) && h->chain_mode == 2) {
        rc = launch_fused_heat(h, d_power, n_frames, d_mfcc, flip180, normalize_first, d_energy, d_mask, d_mean, d_heat, out_h, out_w);
        if (rc != AIG_OK) return io.abort(rc);
        return io.finish();
    }
    // small batches, odd shapes, exact mode: the same arithmetic as separate launches
    if (d_energy == nullptr) d_energy = static_cast<double*>(scratch(h, n * kFramePixels * sizeof(double)));
    if (d_energy == nullptr) return io.abort(AIG_ERR_ALLOC);
    if (n_frames >= small_batch_limit(h) && h->chain_mode == 2) {
        rc = launch_fused(h, d_power, n_frames, d_mfcc, flip180, normalize_first, d_energy, d_mask, d_mean);
    } else {
        rc = launch_mfcc(h, d_power, n_frames * kFramePixels, d_mfcc, flip180, kFramePixels);
        if (rc == AIG_OK) rc = launch_energy(h, h->stream, d_mfcc, n_frames, normalize_first, nullptr, d_energy, d_mask, d_mean);
    }
    if (rc != AIG_OK) return io.abort(rc);
    rc = launch_heatmap(h, d_energy, n_frames, out_h, out_w, d_heat);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

static int check_sweep_args(aig_handle* h, const char* who, int64_t n, const double* thr, int k, int64_t* pos, int64_t* num) {
    if (n < 0 || k < 0 || k > kMaxThresholds) return h->fail(AIG_ERR_ARGUMENT, "%s: n=%lld k=%d out of range (k <= %d)", who, (long long)n, k, kMaxThresholds);
    if ((k > 0 && (!thr || !pos)) || !num) return h->fail(AIG_ERR_ARGUMENT, "%s: null threshold / count buffers", who);
    return AIG_OK;
}

int aig_iou_sweep(aig_handle* h, const uint8_t* mask_a, const uint8_t* mask_b, int64_t n, const double* thr, int k,
                  int64_t* inter_out, int64_t* union_out, int64_t* pos_inout, int64_t* num_inout) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    rc = check_sweep_args(h, "aig_iou_sweep", n, thr, k, pos_inout, num_inout);
    if (rc != AIG_OK) return rc;
    if (n > 0 && (!mask_a || !mask_b)) return h->fail(AIG_ERR_ARGUMENT, "aig_iou_sweep: null masks");
    if (n == 0) return AIG_OK;
    Io io(h);
    const size_t cnt = static_cast<size_t>(n);
    const uint8_t* d_a = io.in(mask_a, cnt * kFramePixels);
    const uint8_t* d_b = io.in(mask_b, cnt * kFramePixels);
    const double* d_thr = io.in(thr, static_cast<size_t>(k));
    int64_t* d_inter = io.out(inter_out, cnt);
    int64_t* d_union = io.out(union_out, cnt);
    int64_t* d_pos = io.inout(pos_inout, static_cast<size_t>(k));
    int64_t* d_num = io.inout(num_inout, 1);
    if (io.failed) return io.finish();
    if ((reinterpret_cast<uintptr_t>(d_a) & 3u) || (reinterpret_cast<uintptr_t>(d_b) & 3u))
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_iou_sweep: mask buffers must be 4-byte aligned"));
    const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(h->sm_count) * 4));
    LaunchScope scope(h, h->stream, kKindOther);
    iou_sweep_kernel<<<blocks, kIouThreads, 0, h->stream>>>(d_a, d_b, n, d_thr, k, reinterpret_cast<long long*>(d_inter),
                                                            reinterpret_cast<long long*>(d_union),
                                                            reinterpret_cast<unsigned long long*>(d_pos),
                                                            reinterpret_cast<unsigned long long*>(d_num));
    rc = scope.done("iou_sweep_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_iou_sweep_clips(aig_handle* h, const uint8_t* mask_a, const uint8_t* mask_b, int64_t n, int64_t frames_per_clip,
                        const double* thr, int k, int64_t* inter_out, int64_t* union_out, int64_t* pos_inout) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n < 0 || k < 1 || k > kMaxThresholds || frames_per_clip < 1 || !thr || !pos_inout)
        return h->fail(AIG_ERR_ARGUMENT, "aig_iou_sweep_clips: bad arguments (n=%lld k=%d frames_per_clip=%lld)", (long long)n, k,
                       (long long)frames_per_clip);
    if (n > 0 && (!mask_a || !mask_b)) return h->fail(AIG_ERR_ARGUMENT, "aig_iou_sweep_clips: null masks");
    if (n == 0) return AIG_OK;
    Io io(h);
    const size_t cnt = static_cast<size_t>(n);
    const size_t n_clips = static_cast<size_t>((n + frames_per_clip - 1) / frames_per_clip);
    const uint8_t* d_a = io.in(mask_a, cnt * kFramePixels);
    const uint8_t* d_b = io.in(mask_b, cnt * kFramePixels);
    const double* d_thr = io.in(thr, static_cast<size_t>(k));
    int64_t* d_inter = io.out(inter_out, cnt);
    int64_t* d_union = io.out(union_out, cnt);
    int64_t* d_pos = io.inout(pos_inout, n_clips * k);
    if (io.failed) return io.finish();
    if ((reinterpret_cast<uintptr_t>(d_a) & 3u) || (reinterpret_cast<uintptr_t>(d_b) & 3u))
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_iou_sweep_clips: mask buffers must be 4-byte aligned"));
    const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(h->sm_count) * 8));
    LaunchScope scope(h, h->stream, kKindOther);
    iou_sweep_clips_kernel<<<blocks, kIouThreads, 0, h->stream>>>(d_a, d_b, n, frames_per_clip, d_thr, k,
                                                                  reinterpret_cast<long long*>(d_inter),
                                                                  reinterpret_cast<long long*>(d_union),
                                                                  reinterpret_cast<unsigned long long*>(d_pos));
    rc = scope.done("iou_sweep_clips_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_ciou_sweep(aig_handle* h, const uint8_t* mask, const int32_t* xmin, const int32_t* xmax, const int32_t* ymin,
                   const int32_t* ymax, int64_t n, int out_h, int out_w, const double* thr, int k, int64_t* inter2_out,
                   int64_t* union2_out, int64_t* pos_inout, int64_t* num_inout) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    rc = check_sweep_args(h, "aig_ciou_sweep", n, thr, k, pos_inout, num_inout);
    if (rc != AIG_OK) return rc;
    if (n > 0 && (!mask || !xmin || !xmax || !ymin || !ymax)) return h->fail(AIG_ERR_ARGUMENT, "aig_ciou_sweep: null inputs");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut)
        return h->fail(AIG_ERR_ARGUMENT, "aig_ciou_sweep: output size %dx%d outside 1..%d", out_h, out_w, kMaxOut);
    if (n == 0) return AIG_OK;
    Io io(h);
    const size_t cnt = static_cast<size_t>(n);
    const uint8_t* d_mask = io.in(mask, cnt * kFramePixels);
    const int32_t* d_xmin = io.in(xmin, cnt * 3);
    const int32_t* d_xmax = io.in(xmax, cnt * 3);
    const int32_t* d_ymin = io.in(ymin, cnt * 3);
    const int32_t* d_ymax = io.in(ymax, cnt * 3);
    const double* d_thr = io.in(thr, static_cast<size_t>(k));
    int64_t* d_inter = io.out(inter2_out, cnt);
    int64_t* d_union = io.out(union2_out, cnt);
    int64_t* d_pos = io.inout(pos_inout, static_cast<size_t>(k));
    int64_t* d_num = io.inout(num_inout, 1);
    if (io.failed) return io.finish();
    if (packed_size(h, out_h, out_w)) {
        auto go = [&](auto kernel, size_t smem, int slot) {
            if (!h->packed_attr_set[slot]) {
                AIG_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                h->packed_attr_set[slot] = true;
            }
            LaunchScope scope(h, h->stream, kKindOther);
            kernel<<<frames_grid(h, n, packed_ctas_per_sm(h, kernel, smem, slot)), kPackedThreads, smem, h->stream>>>(
                d_mask, d_xmin, d_xmax, d_ymin, d_ymax, n, d_thr, k, reinterpret_cast<long long*>(d_inter),
                reinterpret_cast<long long*>(d_union), reinterpret_cast<unsigned long long*>(d_pos),
                reinterpret_cast<unsigned long long*>(d_num));
            return scope.done("ciou_packed_kernel");
        };
        rc = out_w == 298 ? go(ciou_packed_kernel<298, 224>, sizeof(CiouPackedSmem<298, 224>), 2)
                          : go(ciou_packed_kernel<224, 224>, sizeof(CiouPackedSmem<224, 224>), 3);
        if (rc != AIG_OK) return io.abort(rc);
        return io.finish();
    }
    const size_t smem = MaskTaps::bytes(out_h, out_w) + static_cast<size_t>(out_w + out_h);
    rc = allow_mask_smem(h);
    if (rc != AIG_OK) return io.abort(rc);
    LaunchScope scope(h, h->stream, kKindOther);
    ciou_sweep_kernel<<<frames_grid(h, n, mask_ctas_per_sm(smem)), kIouThreads, smem, h->stream>>>(
        d_mask, d_xmin, d_xmax, d_ymin, d_ymax, n, out_h, out_w, d_thr, k, reinterpret_cast<long long*>(d_inter),
        reinterpret_cast<long long*>(d_union), reinterpret_cast<unsigned long long*>(d_pos),
        reinterpret_cast<unsigned long long*>(d_num));
    rc = scope.done("ciou_sweep_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

static int ensure_twiddles(aig_handle* h) {
    if (h->d_twiddle != nullptr) return AIG_OK;
    std::vector<double2> tw(kAudioSamples / 2);
    for (int k = 0; k < kAudioSamples / 2; ++k) {
        const double ang = -2.0 * 3.14159265358979323846 * k / kAudioSamples;
        tw[k] = make_double2(std::cos(ang), std::sin(ang));
    }
    if (cudaMalloc(&h->d_twiddle, tw.size() * sizeof(double2)) != cudaSuccess) {
        cudaGetLastError();
        return h->fail(AIG_ERR_ALLOC, "aig_power_spectrum: cudaMalloc failed");
    }
    AIG_CK(cudaMemcpy(h->d_twiddle, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    return AIG_OK;
}

static int launch_spectrum(aig_handle* h, const float* d_in, int audio_is_int32, int64_t n_rows, const double* d_win, float* d_out) {
    const int grid = frames_grid(h, n_rows, 8);
    LaunchScope scope(h, h->stream, kKindOther);
    if (audio_is_int32)
        spectrum_kernel<int><<<grid, kSpectrumThreads, 0, h->stream>>>(reinterpret_cast<const int*>(d_in), n_rows, d_win, h->d_twiddle, d_out);
    else
        spectrum_kernel<float><<<grid, kSpectrumThreads, 0, h->stream>>>(d_in, n_rows, d_win, h->d_twiddle, d_out);
    return scope.done("spectrum_kernel");
}

int aig_power_spectrum(aig_handle* h, const void* audio, int audio_is_int32, int64_t n_rows, const double* window,
                       float* power_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_rows < 0 || (n_rows > 0 && (!audio || !power_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_power_spectrum: bad buffers");
    if (n_rows == 0) return AIG_OK;
    rc = ensure_twiddles(h);
    if (rc != AIG_OK) return rc;
    Io io(h);
    const size_t n = static_cast<size_t>(n_rows);
    const float* d_in = io.in(static_cast<const float*>(audio), n * kAudioSamples);      // 4-byte samples either way
    const double* d_win = io.in(window, kAudioSamples);
    float* d_out = io.out(power_out, n * (kAudioSamples / 2));
    if (io.failed) return io.finish();
    rc = launch_spectrum(h, d_in, audio_is_int32, n_rows, d_win, d_out);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_audio_mfcc(aig_handle* h, const void* audio, int audio_is_int32, int64_t n_rows, const double* window,
                   float* mfcc_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (!h->tables_set) return h->fail(AIG_ERR_TABLES, "aig_audio_mfcc: aig_set_tables has not been called");
    if (h->fft_len != kAudioSamples / 2)
        return h->fail(AIG_ERR_TABLES, "aig_audio_mfcc: the tables are for %d bins, the 1024-sample spectrum has %d", h->fft_len, kAudioSamples / 2);
    if (n_rows < 0 || (n_rows > 0 && (!audio || !mfcc_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_audio_mfcc: bad buffers");
    if (n_rows == 0) return AIG_OK;
    rc = ensure_twiddles(h);
    if (rc != AIG_OK) return rc;
    Io io(h);
    const size_t n = static_cast<size_t>(n_rows);
    const float* d_in = io.in(static_cast<const float*>(audio), n * kAudioSamples);
    const double* d_win = io.in(window, kAudioSamples);
    float* d_out = io.out(mfcc_out, n * h->mfcc_num);
    float* d_power = static_cast<float*>(scratch(h, n * (kAudioSamples / 2) * sizeof(float)));   // never leaves the device
    if (io.failed || !d_power) { io.failed = true; return io.finish(); }
    rc = launch_spectrum(h, d_in, audio_is_int32, n_rows, d_win, d_power);
    if (rc != AIG_OK) return io.abort(rc);
    rc = launch_mfcc(h, d_power, n_rows, d_out, 0, 1);
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_filtfilt(aig_handle* h, const void* x, int x_is_int32, int64_t n_rows, int length, const double* b, const double* a,
                 const double* zi, int ntaps, float* y_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_rows < 0 || !b || !a || !zi || (n_rows > 0 && (!x || !y_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_filtfilt: bad buffers");
    if (ntaps < 2 || ntaps > kMaxTaps) return h->fail(AIG_ERR_ARGUMENT, "aig_filtfilt: ntaps %d outside 2..%d", ntaps, kMaxTaps);
    const int pad = 3 * ntaps;
    if (length <= pad) return h->fail(AIG_ERR_ARGUMENT, "aig_filtfilt: rows of %d samples are not longer than padlen %d", length, pad);
    if (a[0] == 0.0) return h->fail(AIG_ERR_ARGUMENT, "aig_filtfilt: a[0] is zero");
    if (n_rows == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_rows);
    std::vector<double> coef(3 * ntaps - 1);
    for (int i = 0; i < ntaps; ++i) { coef[i] = b[i]; coef[ntaps + i] = a[i]; }
    for (int i = 0; i < ntaps - 1; ++i) coef[2 * ntaps + i] = zi[i];
    const float* d_x = io.in(static_cast<const float*>(x), n * length);
    const double* d_coef = io.in(coef.data(), coef.size());
    float* d_y = io.out(y_out, n * length);
    double* d_scratch = static_cast<double*>(scratch(h, n * (length + 2 * pad) * sizeof(double)));
    if (io.failed || !d_scratch) return io.finish();
    io.any_host = true;                          // `coef` is a stack-lifetime host buffer: complete before returning
    const int threads = 32, blocks = static_cast<int>((n_rows + threads - 1) / threads);
    LaunchScope scope(h, h->stream, kKindOther);
    if (x_is_int32)
        filtfilt_kernel<int><<<blocks, threads, 0, h->stream>>>(reinterpret_cast<const int*>(d_x), n_rows, length, ntaps, pad, d_coef, d_scratch, d_y);
    else
        filtfilt_kernel<float><<<blocks, threads, 0, h->stream>>>(d_x, n_rows, length, ntaps, pad, d_coef, d_scratch, d_y);
    rc = scope.done("filtfilt_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_normalize_mfcc(aig_handle* h, const float* mfcc, int64_t n, float* out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n < 0 || (n > 0 && (!mfcc || !out))) return h->fail(AIG_ERR_ARGUMENT, "aig_normalize_mfcc: bad buffers");
    if (n == 0) return AIG_OK;
    Io io(h);
    const float* d_in = io.in(mfcc, static_cast<size_t>(n) * kMfccNum);
    float* d_out = io.out(out, static_cast<size_t>(n) * kMfccNum);
    if (io.failed) return io.finish();
    LaunchScope scope(h, h->stream, kKindOther);
    normalize_mfcc_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, h->stream>>>(d_in, n, d_out);
    rc = scope.done("normalize_mfcc_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_tile_mfcc(aig_handle* h, const float* mfcc, int64_t n, int normalize, float* map_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n < 0 || (n > 0 && (!mfcc || !map_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_tile_mfcc: bad buffers");
    if (n == 0) return AIG_OK;
    Io io(h);
    const float* d_in = io.in(mfcc, static_cast<size_t>(n) * kMfccNum);
    float* d_out = io.out(map_out, static_cast<size_t>(n) * kFrameValues);
    if (io.failed) return io.finish();
    if (((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15u) != 0)
        return io.abort(h->fail(AIG_ERR_ARGUMENT, "aig_tile_mfcc: device buffers must be 16-byte aligned"));
    LaunchScope scope(h, h->stream, kKindOther);
    tile_mfcc_kernel<<<frames_grid(h, n, 8), kTileThreads, 0, h->stream>>>(d_in, n, normalize, d_out);
    rc = scope.done("tile_mfcc_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_split_triplets(aig_handle* h, const float* images, int64_t n_frames, float* triplets_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!images || !triplets_out)))
        return h->fail(AIG_ERR_ARGUMENT, "aig_split_triplets: bad buffers");
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t values = static_cast<size_t>(n_frames) * kFrameValues;
    const float* d_in = io.in(images, values);
    float* d_out = io.out(triplets_out, values);
    if (io.failed) return io.finish();
    const long long n_pixels = static_cast<long long>(n_frames) * kFramePixels;
    const int grid = static_cast<int>(std::min<long long>((n_pixels + kTripletTile - 1) / kTripletTile,
                                                          static_cast<long long>(h->sm_count) * 8));
    LaunchScope scope(h, h->stream, kKindOther);
    split_triplets_kernel<<<grid, kTripletThreads, 0, h->stream>>>(d_in, n_pixels, d_out);
    rc = scope.done("split_triplets_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_triplet_mse(aig_handle* h, const float* a, const float* b, int64_t n_frames, double* mse_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames <= 0 || !a || !b || !mse_out) return h->fail(AIG_ERR_ARGUMENT, "aig_triplet_mse: bad buffers");
    Io io(h);
    const size_t values = static_cast<size_t>(n_frames) * kFrameValues;
    const float* d_a = io.in(a, values);
    const float* d_b = io.in(b, values);
    double* d_out = io.out(mse_out, 5);
    const long long n_pixels = static_cast<long long>(n_frames) * kFramePixels;
    const int grid = static_cast<int>(std::min<long long>((n_pixels + kTripletThreads - 1) / kTripletThreads,
                                                          static_cast<long long>(h->sm_count) * 8));
    double* d_partial = static_cast<double*>(scratch(h, static_cast<size_t>(grid) * 4 * sizeof(double)));
    if (io.failed || !d_partial) { io.failed = true; return io.finish(); }
    {
        LaunchScope scope(h, h->stream, kKindOther);
        triplet_mse_kernel<<<grid, kTripletThreads, 0, h->stream>>>(d_a, d_b, n_pixels, d_partial);
        rc = scope.done("triplet_mse_kernel");
        if (rc != AIG_OK) return io.abort(rc);
    }
    {
        LaunchScope scope(h, h->stream, kKindOther);
        triplet_mse_finish_kernel<<<1, 32, 0, h->stream>>>(d_partial, grid, n_pixels, d_out);
        rc = scope.done("triplet_mse_finish_kernel");
        if (rc != AIG_OK) return io.abort(rc);
    }
    return io.finish();
}

int aig_overlay(aig_handle* h, const float* heat, const uint8_t* bgr, int64_t n_frames, int out_h, int out_w, float alpha,
                const uint8_t* jet_lut, uint8_t* rgb_out) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (n_frames < 0 || !jet_lut || (n_frames > 0 && (!heat || !rgb_out))) return h->fail(AIG_ERR_ARGUMENT, "aig_overlay: bad buffers");
    if (out_h < 1 || out_w < 1 || out_h > kMaxOut || out_w > kMaxOut || !(alpha >= 0.f && alpha <= 1.f))
        return h->fail(AIG_ERR_ARGUMENT, "aig_overlay: bad geometry %dx%d or alpha %g", out_h, out_w, alpha);
    if (n_frames == 0) return AIG_OK;
    Io io(h);
    const size_t n = static_cast<size_t>(n_frames), px = static_cast<size_t>(out_h) * out_w;
    const float* d_heat = io.in(heat, n * px);
    const uint8_t* d_bgr = io.in(bgr, n * px * 3);
    const uint8_t* d_lut = io.in(jet_lut, 768);
    uint8_t* d_out = io.out(rgb_out, n * px * 3);
    if (io.failed) return io.finish();
    LaunchScope scope(h, h->stream, kKindOther);
    const bool vec = px % 4 == 0 && (reinterpret_cast<uintptr_t>(d_heat) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(d_bgr) & 3u) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3u) == 0;
    // one CTA per frame, handed out by the hardware as CTAs retire: a persistent grid of SM-count multiples ends on a
    // partial wave worth up to a whole frame time (2048 frames on 1184 CTAs: 2 rounds for 1.73 frames per CTA)
    const unsigned overlay_grid = static_cast<unsigned>(std::min<int64_t>(n_frames, 1 << 20));
    // with a frame of up to 80 k pixels the luma plane stays in shared memory between the two passes (two CTAs per SM)
    const size_t luma_bytes = px;
    if (vec && d_bgr != nullptr && h->overlay_luma && luma_bytes <= kOverlayLumaMaxBytes) {
        auto kernel = overlay_kernel<4, true>;
        if (!h->overlay_luma_attr_set) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kOverlayLumaMaxBytes));
            h->overlay_luma_attr_set = true;
        }
        kernel<<<overlay_grid, kOverlayThreads, luma_bytes, h->stream>>>(d_heat, d_bgr, n_frames, static_cast<int>(px), alpha, d_lut, d_out);
    } else if (vec)
        overlay_kernel<4><<<overlay_grid, kOverlayThreads, 0, h->stream>>>(d_heat, d_bgr, n_frames, static_cast<int>(px), alpha, d_lut, d_out);
    else
        overlay_kernel<1><<<overlay_grid, kOverlayThreads, 0, h->stream>>>(d_heat, d_bgr, n_frames, static_cast<int>(px), alpha, d_lut, d_out);
    rc = scope.done("overlay_kernel");
    if (rc != AIG_OK) return io.abort(rc);
    return io.finish();
}

int aig_comm_unique_id(uint8_t* id_out) {
    if (id_out == nullptr) return AIG_ERR_ARGUMENT;
    NcclApi& api = nccl();
    if (!api.ok) { g_create_error = "aig_comm_unique_id: libnccl.so.2 could not be loaded"; return AIG_ERR_NO_DEVICE; }
    NcclId id;
    const int r = api.get_unique_id(&id);
    if (r != 0) { g_create_error = "aig_comm_unique_id: ncclGetUniqueId failed"; return AIG_ERR_CUDA_BASE; }
    std::memcpy(id_out, id.internal, AIG_COMM_ID_BYTES);
    return AIG_OK;
}

int aig_comm_destroy(aig_handle* h) {
    if (h == nullptr) return AIG_ERR_ARGUMENT;
    if (h->comm != nullptr) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        nccl().comm_destroy(h->comm);
        h->comm = nullptr;
        h->comm_world = 1;
    }
    return AIG_OK;
}

int aig_comm_init(aig_handle* h, const uint8_t* id, int rank, int world) {
    if (h == nullptr) return AIG_ERR_ARGUMENT;
    if (id == nullptr || world < 1 || rank < 0 || rank >= world) return h->fail(AIG_ERR_ARGUMENT, "aig_comm_init: bad rank/world %d/%d", rank, world);
    NcclApi& api = nccl();
    if (!api.ok) return h->fail(AIG_ERR_NO_DEVICE, "aig_comm_init: libnccl.so.2 could not be loaded");
    aig_comm_destroy(h);
    AIG_CK(cudaSetDevice(h->device));
    NcclId nid;
    std::memcpy(nid.internal, id, AIG_COMM_ID_BYTES);
    const int r = api.comm_init_rank(&h->comm, world, nid, rank);
    if (r != 0) {
        h->comm = nullptr;
        return h->fail(AIG_ERR_CUDA_BASE, "ncclCommInitRank failed: %s", api.get_error_string ? api.get_error_string(r) : "?");
    }
    h->comm_world = world;
    return AIG_OK;
}

int aig_allreduce_counts(aig_handle* h, int64_t* counts, int n) {
    int rc = require(h);
    if (rc != AIG_OK) return rc;
    if (counts == nullptr || n < 0) return h->fail(AIG_ERR_ARGUMENT, "aig_allreduce_counts: bad buffer");
    if (h->comm == nullptr || n == 0) return AIG_OK;    // no communicator: identity (a one-rank communicator still runs NCCL)
    Io io(h);
    int64_t* d = io.inout(counts, static_cast<size_t>(n));
    if (io.failed) return io.finish();
    const int r = nccl().all_reduce(d, d, static_cast<size_t>(n), kNcclInt64, kNcclSum, h->comm, h->stream);
    if (r != 0) return h->fail(AIG_ERR_CUDA_BASE, "ncclAllReduce failed: %s", nccl().get_error_string ? nccl().get_error_string(r) : "?");
    return io.finish();
}

int aig_comm_init_all(aig_handle** handles, int n) {
    if (handles == nullptr || n < 1) return AIG_ERR_ARGUMENT;
    for (int i = 0; i < n; ++i) if (handles[i] == nullptr) return AIG_ERR_ARGUMENT;
    aig_handle* h0 = handles[0];
    DeviceGuard guard;
    NcclApi& api = nccl();
    if (!api.ok || !api.comm_init_all) return h0->fail(AIG_ERR_NO_DEVICE, "aig_comm_init_all: libnccl.so.2 could not be loaded");
    std::vector<int> devices(n);
    std::vector<void*> comms(n, nullptr);
    for (int i = 0; i < n; ++i) {
        aig_comm_destroy(handles[i]);
        devices[i] = handles[i]->device;
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return h0->fail(AIG_ERR_ARGUMENT, "aig_comm_init_all: device %d listed twice", devices[i]);
    }
    const int r = api.comm_init_all(comms.data(), n, devices.data());
    if (r != 0) return h0->fail(AIG_ERR_CUDA_BASE, "ncclCommInitAll failed: %s", api.get_error_string ? api.get_error_string(r) : "?");
    for (int i = 0; i < n; ++i) { handles[i]->comm = comms[i]; handles[i]->comm_world = n; }
    return AIG_OK;
}

int aig_group_allreduce_counts(aig_handle** handles, int64_t* const* counts, int n_handles, int n) {
    if (handles == nullptr || counts == nullptr || n_handles < 1 || n < 0) return AIG_ERR_ARGUMENT;
    aig_handle* h0 = handles[0];
    if (h0 == nullptr) return AIG_ERR_ARGUMENT;
    NcclApi& api = nccl();
    if (!api.ok || !api.group_start || !api.group_end) return h0->fail(AIG_ERR_NO_DEVICE, "aig_group_allreduce_counts: NCCL is not available");
    for (int i = 0; i < n_handles; ++i) {
        if (handles[i] == nullptr || counts[i] == nullptr || handles[i]->comm == nullptr || handles[i]->comm_world != n_handles)
            return h0->fail(AIG_ERR_ARGUMENT, "aig_group_allreduce_counts: handle %d is not part of a %d-rank aig_comm_init_all communicator", i, n_handles);
        if (classify(counts[i]) != kDevice) return h0->fail(AIG_ERR_ARGUMENT, "aig_group_allreduce_counts: counts[%d] must be device memory", i);
    }
    if (n == 0) return AIG_OK;
    DeviceGuard guard;
    int r = api.group_start();
    for (int i = 0; i < n_handles && r == 0; ++i) {
        cudaSetDevice(handles[i]->device);
        r = api.all_reduce(counts[i], counts[i], static_cast<size_t>(n), kNcclInt64, kNcclSum, handles[i]->comm, handles[i]->stream);
    }
    const int r_end = api.group_end();
    if (r == 0) r = r_end;
    if (r != 0) return h0->fail(AIG_ERR_CUDA_BASE, "grouped ncclAllReduce failed: %s", api.get_error_string ? api.get_error_string(r) : "?");
    return AIG_OK;
}

int aig_auc(const double* thr, const double* value, int k, double* auc_out) {
    if (!thr || !value || !auc_out || k < 2) return AIG_ERR_ARGUMENT;
    // areaundercurve.py:32-37: both arrays reversed, then sklearn.metrics.auc: direction * sum(diff(x) * (y[1:] + y[:-1]) / 2)
    bool inc = true, dec = true;
    for (int i = 1; i < k; ++i) {
        const double d = thr[k - 1 - i] - thr[k - i];   // diff of the reversed x
        if (d > 0) dec = false;
        if (d < 0) inc = false;
    }
    if (!inc && !dec) return AIG_ERR_ARGUMENT;
    // NumPy's sum over k-1 terms: plain left-to-right for < 8 terms, else 8 strided partial sums
    std::vector<double> term(k - 1);
    for (int i = 0; i < k - 1; ++i) {
        const double dx = thr[k - 2 - i] - thr[k - 1 - i];
        term[i] = dx * (value[k - 2 - i] + value[k - 1 - i]) / 2.0;
    }
    const int n = k - 1;
    double sum;
    if (n < 8) {
        sum = 0.0;
        for (int i = 0; i < n; ++i) sum += term[i];
    } else {
        // pairwise_sum of numpy for n <= 128 (the sweep has at most kMaxThresholds terms: recurse above 128)
        struct Pw {
            static double run(const double* a, int n) {
                if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
                if (n <= 128) {
                    double r[8];
                    for (int j = 0; j < 8; ++j) r[j] = a[j];
                    int i = 8;
                    for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
                    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
                    for (; i < n; ++i) res += a[i];
                    return res;
                }
                int n2 = n / 2;
                n2 -= n2 % 8;
                return run(a, n2) + run(a + n2, n - n2);
            }
        };
        sum = Pw::run(term.data(), n);
    }
    *auc_out = ((dec && !inc) ? -1.0 : 1.0) * sum;
    return AIG_OK;
}

}  // extern "C"
