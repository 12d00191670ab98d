// Shared device helpers for libaig.so (sm_100a): mbarrier / TMA PTX wrappers and warp reductions.
#pragma once

#include <cuda.h>           // CUtensorMap (type only; the driver entry point is fetched at run time)
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace aig {

constexpr int kFrameH = 36;
constexpr int kFrameW = 48;
constexpr int kFramePixels = kFrameH * kFrameW;   // 1728
constexpr int kMfccNum = 12;
constexpr int kFilterNum = 24;
constexpr int kFftLen = 512;
constexpr int kFrameValues = kFramePixels * kMfccNum;   // 20736

// ---- shared-memory addressing ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Make barrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// For waits that usually last microseconds: the polling loop of mbar_wait costs two issue slots every ~30 cycles per
// waiting warp (ncu, energy_heat_ws_kernel: a fifth of all instructions issued were polls), which the warps being
// waited for could use.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned int ns) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---- TMA --------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
// 16-byte global store with an L2 cache-hint policy.
__device__ __forceinline__ void stg128_hint(float4* p, const float4& v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled bulk tensor load global -> shared, completion signalled on an mbarrier.
// policy == 0: default L2 policy (measured faster for this read-mostly stream, profiles/r01_read_peak_probe.txt);
// otherwise an L2 cache-hint policy from createpolicy.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int32_t x, int32_t y, uint64_t policy) {
    if (policy == 0) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
            : "memory");
        return;
    }
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "l"(policy)
        : "memory");
}

// ---- bulk asynchronous loads global -> shared (TMA unit, no tensor map) -------------------------------
// src (global) and dst (shared) 16-byte aligned, bytes a multiple of 16; completion arrives on the mbarrier as
// transaction bytes (arm it with mbar_arrive_expect_tx first).
__device__ __forceinline__ void bulk_load_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- bulk asynchronous stores shared -> global (TMA unit, SASS UBLKCP) ---------------------------
// Generic-proxy writes to shared memory (st.shared) must be fenced before the async proxy reads them.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// dst (global) and src (shared) 16-byte aligned, bytes a multiple of 16.  Completion is tracked by the issuing
// thread's bulk-copy groups.
__device__ __forceinline__ void bulk_store_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until at most N of this thread's groups are still reading their shared-memory source ...
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... or are incomplete altogether (writes performed).
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- L2 prefetch -------------------------------------------------------------------------------------------------
// Asks a line into L2 without a register or shared-memory slot waiting for it: the loads that follow find it there at a
// third of the DRAM latency, which is what a kernel short of bytes in flight needs (overlay_kernel: 7.3 -> 9.3 M frames/s).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- debug jitter ("debug_jitter" option) -------------------------------------------------------
// compute-sanitizer's racecheck is not available on the pool this library is tested on.  The hand-rolled mbarrier
// protocols are instead exercised by perturbing the relative timing of their roles: every call spins for a
// pseudo-random 0 .. 2^14 cycles derived from (seed, salt, CTA, warp, call count).  A protocol that is only correct by
// timing luck then produces results that differ from the sequential two-kernel path (tests/test_gpu_stress.py).
__device__ __forceinline__ void jitter_spin(unsigned int seed, unsigned int salt, unsigned int& counter) {
    unsigned int x = seed * 2654435761u ^ (salt + 0x9e3779b9u) * 40503u ^ (blockIdx.x * 977u + (threadIdx.x >> 5)) * 2246822519u ^
                     (++counter) * 3266489917u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    if ((x & 3u) != 0u) return;                            // three calls out of four pass straight through
    const long long until = clock64() + ((x >> 8) & 0x3fffu);
    while (clock64() < until) {
    }
}

// ---- shared-memory vector load ----------------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

// ---- warp reductions --------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace aig
