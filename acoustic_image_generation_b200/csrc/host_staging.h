// Host -> device upload of ordinary (pageable) memory at link speed.
//
// The reference hands this path NumPy arrays that TensorFlow's py_func allocated (dataloader/outdoor_data_mfcc.py:788):
// pageable memory.  cudaMemcpyAsync from pageable memory is staged by the driver on one thread and reaches 8-12 GB/s on
// the B200 boxes (tools/pageable_probe.py) against 55 GB/s for pinned memory - the whole end-to-end call is then 4-5x
// slower than the link allows.  StagedUploader does the staging itself: a few host threads copy 4 MiB pieces of the
// caller's array into a ring of pinned slots while the calling thread issues one asynchronous H2D copy per filled slot,
// in order, on the given stream.  download() is the mirror image for results returned into pageable arrays: D2H copies
// land in the pinned slots and the threads move them on into the caller's array.  CUDA is only ever called from the
// calling thread.
//
// Host-side runtime code, no device code: the kernels never see the difference.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// host_copy.cpp: slot fills with non-temporal stores (no read-for-ownership of the pinned slot)
extern "C" void aig_host_copy_streaming(void* dst, const void* src, size_t bytes);
extern "C" int aig_host_copy_streaming_supported();

namespace aig {

class StagedUploader {   // both directions; named after its first job
   public:
    static constexpr size_t kSlotBytes = size_t(4) << 20;
    static constexpr int kSlots = 8;
    // A copy engine needs 4 MiB transfers to reach the link rate (1 MiB: 45 GB/s, 4 MiB: 52.5, 16 MiB: 54.7 on this box,
    // tools/h2d_streams_probe.py); small jobs use 1 MiB pieces so that the first transfer starts sooner.
    // Jobs under 4 MiB (reached only when option staged_min_bytes is lowered: a batch of 16 acoustic images is 1.3 MB)
    // are cut finer still (small_piece_, default 512 KiB; tools/small_upload_probe.py: 128 KiB pieces cost 20 %).
    size_t piece_bytes(size_t bytes) const {
        return bytes >= (size_t(32) << 20) ? kSlotBytes : bytes >= (size_t(4) << 20) ? (size_t(1) << 20) : small_piece_;
    }
    void set_solo_bytes(size_t bytes) { solo_bytes_ = bytes; }
    void set_small_piece(size_t bytes) { small_piece_ = std::min(kSlotBytes, std::max<size_t>(bytes, 64 << 10)); }
    static constexpr int kLingerMicros = 400;
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    StagedUploader() = default;
    StagedUploader(const StagedUploader&) = delete;
    StagedUploader& operator=(const StagedUploader&) = delete;
    ~StagedUploader() { shutdown(); }

    // Lazily allocates the pinned ring and starts `threads` workers.  Returns false if pinned memory is unavailable.
    bool start(int threads) {
        if (ready_) return true;
        if (threads < 1) return false;
        if (ring_ == nullptr && cudaHostAlloc(reinterpret_cast<void**>(&ring_), kSlotBytes * kSlots, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            ring_ = nullptr;
            return false;
        }
        for (int s = 0; s < kSlots; ++s) {
            if (slot_done_[s] == nullptr && cudaEventCreateWithFlags(&slot_done_[s], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            slot_busy_[s] = false;
        }
        stop_ = false;
        try {
            for (int t = 0; t < threads; ++t) workers_.emplace_back([this] { worker(); });
        } catch (...) {                       // thread creation refused: run with the workers that did start, if any
            if (workers_.empty()) return false;
        }
        ready_ = true;
        return true;
    }

    void shutdown() {
        if (!workers_.empty()) {
            {
                std::lock_guard<std::mutex> lock(m_);
                stop_ = true;
                generation_hint_.store(~uint64_t(0), std::memory_order_release);
            }
            cv_.notify_all();
            for (auto& w : workers_) w.join();
            workers_.clear();
        }
        if (ring_ != nullptr) {
            for (int s = 0; s < kSlots; ++s)
                if (slot_done_[s]) { cudaEventSynchronize(slot_done_[s]); cudaEventDestroy(slot_done_[s]); slot_done_[s] = nullptr; }
            cudaFreeHost(ring_);
            ring_ = nullptr;
        }
        ready_ = false;
    }

    bool ready() const { return ready_; }
    // Upload fills: streaming stores or std::memcpy.  Set between jobs only.  mode 0: memcpy; 1: streaming stores
    // (where the CPU has AVX2); -1 (default): streaming stores for jobs up to kStreamingAutoBytes.  Measured on a B200
    // box (tools/c1_breakdown_probe.py, profiles/r02_c1_breakdown_streaming.txt, six threads): a 57 MB job whose source
    // is still in the host's caches 1.56 ms against 1.94 ms with memcpy (the slots stop evicting the source), the same
    // job from a cold source 1.95 against 2.17 ms, but a 906 MB job 22.6 against 21.1 ms - there the 32 MB ring stays
    // cache-resident under ordinary stores and the copy engine reads it from the cache, which streaming stores undo.
    static constexpr size_t kStreamingAutoBytes = size_t(128) << 20;
    void set_streaming_fill(int mode) { streaming_mode_ = aig_host_copy_streaming_supported() != 0 ? mode : 0; }
    int streaming_fill() const { return streaming_mode_; }
    int threads() const { return static_cast<int>(workers_.size()); }

    // Copies bytes from pageable `src` to device `dst`, ordered on `stream`.  Returns once every piece has been read
    // from `src` and its H2D copy enqueued (the caller may then reuse `src`; the device side is stream-ordered).
    cudaError_t upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
        if (bytes == 0) return cudaSuccess;
        const size_t piece = piece_bytes(bytes);
        const int64_t pieces = static_cast<int64_t>((bytes + piece - 1) / piece);
        bool solo = false;
        {
            std::unique_lock<std::mutex> lock(m_);
            cv_idle_.wait(lock, [this] { return active_ == 0; });     // a late waker of the previous job has left it
            src_ = static_cast<const char*>(src);
            user_dst_ = nullptr;
            bytes_ = bytes;
            piece_ = piece;
            streaming_job_ = streaming_mode_ > 0 || (streaming_mode_ < 0 && bytes <= kStreamingAutoBytes);
            worker_seen_.store(false, std::memory_order_relaxed);
            pieces_ = pieces;
            next_.store(0, std::memory_order_relaxed);
            freed_.store(0, std::memory_order_relaxed);
            for (int s = 0; s < kSlots; ++s) filled_[s].store(-1, std::memory_order_relaxed);
            // Jobs under solo_bytes_ (option staged_solo_bytes, default 0 = none) are copied by the calling thread alone, piece
            // by piece, each piece's H2D copy overlapping the next fill, without waking the workers.
            solo = bytes < solo_bytes_;
            if (!solo) {
                ++generation_;
                generation_hint_.store(generation_, std::memory_order_release);
            }
        }
        if (!solo) cv_.notify_all();
        // slots of an earlier upload may still be in flight on the DMA engine: their events gate the first reuse
        cudaError_t status = cudaSuccess;
        int64_t issued = 0, freed = 0;
        // piece i may be written into its slot once piece i - kSlots has left it; for the first kSlots pieces that is
        // the previous upload's copy out of the same slot
        int64_t primed = 0;
        const int64_t first_lap = std::min<int64_t>(pieces, kSlots);
        int64_t mine = -1;                       // piece claimed by the calling thread itself, waiting for its slot
        unsigned spins = 0;
        while (issued < pieces) {
            bool progressed = false;
            // release slots whose copies have completed, in order
            while (primed < first_lap) {
                const int s = static_cast<int>(primed % kSlots);
                if (slot_busy_[s] && cudaEventQuery(slot_done_[s]) == cudaErrorNotReady) break;
                slot_busy_[s] = false;
                ++primed;
                progressed = true;
            }
            while (freed < issued && cudaEventQuery(slot_done_[freed % kSlots]) != cudaErrorNotReady) {
                slot_busy_[freed % kSlots] = false;
                ++freed;
                progressed = true;
            }
            // a worker may fill piece i when i < primed (first lap) and i - kSlots < freed (later laps)
            freed_.store(primed < first_lap ? primed : freed + kSlots, std::memory_order_release);
            const int s = static_cast<int>(issued % kSlots);
            if (filled_[s].load(std::memory_order_acquire) == issued) {
                const size_t off = static_cast<size_t>(issued) * piece;
                const size_t len = std::min(piece, bytes - off);
                if (status == cudaSuccess) {
                    status = cudaMemcpyAsync(static_cast<char*>(dst) + off, ring_ + s * kSlotBytes, len, cudaMemcpyHostToDevice, stream);
                    if (status == cudaSuccess) status = cudaEventRecord(slot_done_[s], stream);
                    slot_busy_[s] = status == cudaSuccess;
                }
                ++issued;
                progressed = true;
            }
            // The calling thread copies too whenever it has nothing to issue: a sleeping worker takes 100-200 us to wake
            // on these (virtualised) hosts, and a 16-frame call is only a millisecond of copying.
            if (mine < 0 && !progressed && !worker_seen_.load(std::memory_order_relaxed) &&
                next_.load(std::memory_order_relaxed) < pieces) {
                const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
                if (i < pieces) mine = i;
            }
            if (mine >= 0 && mine < freed_.load(std::memory_order_relaxed)) {
                const size_t off = static_cast<size_t>(mine) * piece;
                const int ms = static_cast<int>(mine % kSlots);
                fill_slot(ring_ + ms * kSlotBytes, static_cast<const char*>(src) + off, std::min(piece, bytes - off));
                filled_[ms].store(mine, std::memory_order_release);
                mine = -1;
                progressed = true;
            }
            if (!progressed) {
                if (++spins > 64) { std::this_thread::yield(); spins = 0; }
            } else {
                spins = 0;
            }
        }
        // park the workers (they leave the job when next_ runs past pieces_) and wait until all have left it
        std::unique_lock<std::mutex> lock(m_);
        cv_idle_.wait(lock, [this] { return active_ == 0; });
        pieces_ = 0;
        if (status == cudaSuccess) cudaGetLastError();     // cudaEventQuery's cudaErrorNotReady is not an error
        return status;
    }

    // Copies bytes from device `src` to pageable `dst` after the work already enqueued on `stream`.  Returns when the data
    // is in `dst`.
    cudaError_t download(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
        if (bytes == 0) return cudaSuccess;
        const size_t piece = piece_bytes(bytes);
        const int64_t pieces = static_cast<int64_t>((bytes + piece - 1) / piece);
        {
            std::unique_lock<std::mutex> lock(m_);
            cv_idle_.wait(lock, [this] { return active_ == 0; });
            src_ = nullptr;
            user_dst_ = static_cast<char*>(dst);
            bytes_ = bytes;
            piece_ = piece;
            pieces_ = pieces;
            next_.store(0, std::memory_order_relaxed);
            freed_.store(0, std::memory_order_relaxed);              // download: pieces [0, freed_) have arrived in their slots
            for (int s = 0; s < kSlots; ++s) filled_[s].store(-1, std::memory_order_relaxed);   // download: piece drained from slot
            ++generation_;
            generation_hint_.store(generation_, std::memory_order_release);
        }
        cv_.notify_all();
        cudaError_t status = cudaSuccess;
        int64_t issued = 0, arrived = 0, drained = 0;
        unsigned spins = 0;
        while (drained < pieces) {
            bool progressed = false;
            while (arrived < issued && cudaEventQuery(slot_done_[arrived % kSlots]) != cudaErrorNotReady) {
                slot_busy_[arrived % kSlots] = false;
                ++arrived;
                progressed = true;
            }
            freed_.store(arrived, std::memory_order_release);
            while (drained < arrived && filled_[drained % kSlots].load(std::memory_order_acquire) == drained) {
                ++drained;
                progressed = true;
            }
            if (issued < pieces && issued < drained + kSlots) {
                const int s = static_cast<int>(issued % kSlots);
                // first lap: an earlier upload may still be reading this slot
                if (!(slot_busy_[s] && cudaEventQuery(slot_done_[s]) == cudaErrorNotReady)) {
                    const size_t off = static_cast<size_t>(issued) * piece;
                    const size_t len = std::min(piece, bytes - off);
                    if (status == cudaSuccess) {
                        status = cudaMemcpyAsync(ring_ + s * kSlotBytes, static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, stream);
                        if (status == cudaSuccess) status = cudaEventRecord(slot_done_[s], stream);
                    }
                    slot_busy_[s] = status == cudaSuccess;
                    if (status != cudaSuccess) {                        // nothing will arrive: let the workers run through
                        failed_.store(true, std::memory_order_release);
                    }
                    ++issued;
                    progressed = true;
                }
            }
            if (failed_.load(std::memory_order_acquire)) break;
            if (!progressed) {
                if (++spins > 64) { std::this_thread::yield(); spins = 0; }
            } else {
                spins = 0;
            }
        }
        if (failed_.load(std::memory_order_acquire)) {
            next_.store(pieces, std::memory_order_relaxed);             // no more claims
            freed_.store(pieces, std::memory_order_release);            // release any waiter (it copies stale bytes; the call fails anyway)
        }
        std::unique_lock<std::mutex> lock(m_);
        cv_idle_.wait(lock, [this] { return active_ == 0; });
        pieces_ = 0;
        failed_.store(false, std::memory_order_relaxed);
        if (status == cudaSuccess) cudaGetLastError();
        return status;
    }

   private:
    void fill_slot(char* slot, const char* src, size_t len) const {
        if (streaming_job_) aig_host_copy_streaming(slot, src, len);
        else std::memcpy(slot, src, len);
    }

    void worker() {
        uint64_t seen = 0;
        for (;;) {
            {
                // linger for a moment before sleeping: calls usually come back to back, and waking costs more than this
                const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(kLingerMicros);
                while (generation_hint_.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < until) {
                    for (int k = 0; k < 32; ++k) cpu_relax();
                }
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                ++active_;
                worker_seen_.store(true, std::memory_order_relaxed);
            }
            for (;;) {
                const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
                if (i >= pieces_) break;
                unsigned spins = 0;
                while (freed_.load(std::memory_order_acquire) <= i) {       // upload: slot not free yet; download: piece not here yet
                    if (++spins > 64) { std::this_thread::yield(); spins = 0; }
                }
                const size_t off = static_cast<size_t>(i) * piece_;
                const size_t len = std::min(piece_, bytes_ - off);
                const int s = static_cast<int>(i % kSlots);
                if (user_dst_ == nullptr) fill_slot(ring_ + s * kSlotBytes, src_ + off, len);       // upload: fill the slot
                else std::memcpy(user_dst_ + off, ring_ + s * kSlotBytes, len);                    // download: drain it
                filled_[s].store(i, std::memory_order_release);
            }
            {
                std::lock_guard<std::mutex> lock(m_);
                --active_;
            }
            cv_idle_.notify_all();
        }
    }

    bool ready_ = false;
    size_t small_piece_ = size_t(512) << 10;
    size_t solo_bytes_ = 0;                  // uploads under this size do not wake the workers
    int streaming_mode_ = aig_host_copy_streaming_supported() != 0 ? -1 : 0;
    bool streaming_job_ = false;             // the current upload job's choice (written before the workers are woken)
    char* ring_ = nullptr;
    cudaEvent_t slot_done_[kSlots] = {};
    bool slot_busy_[kSlots] = {};
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, cv_idle_;
    bool stop_ = false;
    uint64_t generation_ = 0;
    std::atomic<uint64_t> generation_hint_{0};   // copy of generation_ the lingering workers poll without the lock
    int active_ = 0;
    // current job
    const char* src_ = nullptr;              // upload: the caller's array
    char* user_dst_ = nullptr;               // download: the caller's array (nullptr selects upload in the workers)
    std::atomic<bool> failed_{false};
    size_t bytes_ = 0;
    size_t piece_ = kSlotBytes;              // bytes per piece of the current job (<= kSlotBytes)
    std::atomic<bool> worker_seen_{false};   // a worker has joined the current job: the calling thread stops copying
    int64_t pieces_ = 0;
    std::atomic<int64_t> next_{0};
    std::atomic<int64_t> freed_{0};          // pieces [0, freed_) may be written into their slots
    std::atomic<int64_t> filled_[kSlots];    // index of the piece a slot currently holds, -1 when stale
};

}  // namespace aig
