// Stages 1 + 2 in ONE persistent kernel: spectra -> MFCC image -> energy map + mean mask, a frame at a time.
//
// Why: the MFCC stage is HBM-bound and leaves the FP64 pipe idle; the energy stage (find_logen) is
// FP64-latency-bound and moves 2 % of the bytes.  Run as separate kernels they either serialise (+30 % time) or
// fight for SM residency when overlapped on two streams.  Here every CTA (one per SM) owns whole frames and is
// warp-specialised three ways:
//
//   warps 0-5   MFCC consumers   192 threads = one spectrum each per 192-row tile, 9 tiles per frame (mel_tile)
//   warp  6     TMA producer     one lane streams the frame's [192 x 32] slabs through the shared-memory ring
//   warps 7-14  energy warps     256 threads: find_logen + mean + mask of the frame the consumers finished last
//
// The consumers store each MFCC row to global memory (it is an output anyway), track the frame's min/max on the fly
// and publish them through a two-slot (full/empty mbarrier) hand-off; the energy warps read the rows back with
// ld.global.cg (L2 hits: the lines were written microseconds earlier by the same SM) so no second shared-memory
// copy of the frame is needed and the whole 192 KiB stays available for loads in flight.  While the energy warps work
// on frame i the consumers are already streaming frame i+1 (and may run up to two frames ahead).
#pragma once

#include "aig_common.cuh"
#include "energy_kernel.cuh"
#include "heatmap_kernel.cuh"
#include "mfcc_kernel.cuh"

namespace aig {

constexpr int kFusedRows = 192;                       // 9 tiles of 192 spectra = one 1728-pixel frame
constexpr int kFusedTilesPerFrame = kFramePixels / kFusedRows;
constexpr int kFusedConsumerWarps = kFusedRows / 32;  // 6
constexpr int kFusedEnergyWarps = 8;
constexpr int kFusedEnergyThreads = kFusedEnergyWarps * 32;
constexpr int kFusedThreads = kFusedRows + 32 + kFusedEnergyThreads;   // 480
static_assert(kFramePixels % kFusedRows == 0, "tiles must not straddle frames");

struct FusedShared {           // lives after the ring in dynamic shared memory
    double map[kFramePixels];
    double part[16][8];
    double leaf[16];
    double mean;
    EnergyTables tab;
    float minmax[2][kFusedConsumerWarps][2];
    unsigned long long bar[4];         // frame_full[2], frame_empty[2]
    double red64[2][kHeatMaxWarps];    // heat-map phase (HEAT builds): min / max reductions of the energy warps
    float red32[2][kHeatMaxWarps];
};

// Output of the opt-in heat-map phase (aig_mfcc_energy_heatmap).
struct FusedHeatOut {
    float* heat;               // [n_frames, out_h, out_w]
    int out_h, out_w;
};

template <int SLABS_PER_STAGE, int STAGES>
struct FusedPipe {
    using P = MfccPipe<kFusedRows, SLABS_PER_STAGE, STAGES>;
    static constexpr int kRingBytes = STAGES * P::kStageBytes;
    static constexpr int kSmemBytes = kRingBytes + 2 * STAGES * 8 + static_cast<int>(sizeof(FusedShared)) + 1024 + 64;
    // HEAT builds append the heat-map phase's rows, staging slots and taps (heat_stream_layout) behind FusedShared
    static __host__ __device__ size_t smem_with_heat(int out_h, int out_w) {
        return static_cast<size_t>(kSmemBytes) + 16 + heat_stream_layout(out_h, out_w, false, kFusedEnergyWarps).total;
    }
};

__device__ __forceinline__ void energy_group_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kFusedEnergyThreads) : "memory");
}

// JITTER = true is the "debug_jitter" build of the same kernel: every role spins for pseudo-random times before its
// barrier waits and arrives (aig_common.cuh: jitter_spin), which moves the three roles through every relative order the
// protocol allows; the production instantiation (JITTER = false) contains none of it.
// mfcc_out is written by the consumer warps and read back by the energy warps of the same CTA: no __restrict__, no
// read-only loads on it.
// HEAT_VEC != 0 (opt-in, aig_mfcc_energy_heatmap): the energy warps go on from the energy map to the normalised,
// up-sampled heat map of showvideo.py:226-228 (heat_phase() of heatmap_kernel.cuh: 2 or 4 pixels per lane, output size
// HW x HH as template constants or 0 = run-time), staging its rows in shared memory behind FusedShared and shipping them
// with bulk copies - 267 KB more per frame (+7 % of the bytes) for which the ring gives up half of its 192 KiB.
template <int SLABS_PER_STAGE, int STAGES, bool JITTER, int HEAT_VEC = 0, int HW = 0, int HH = 0>
__global__ void __launch_bounds__(kFusedThreads, 1)
mfcc_energy_fused_kernel(const __grid_constant__ CUtensorMap tmap, float* mfcc_out,
                         unsigned int n_frames, int flip180, int normalize_first,
                         double* __restrict__ energy_out, uint8_t* __restrict__ mask_out,
                         double* __restrict__ mean_out, int l2_evict_first, int keep_mfcc_in_l2,
                         unsigned int jitter_seed, const FusedHeatOut heat_out) {
    using P = MfccPipe<kFusedRows, SLABS_PER_STAGE, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = ring + STAGES * P::kStageBytes;
    const uint32_t bar_empty = bar_full + STAGES * 8;
    uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
    FusedShared& sh = *reinterpret_cast<FusedShared*>(ring_ptr + STAGES * P::kStageBytes + 2 * STAGES * 8 +
                                                      ((16 - (2 * STAGES * 8) % 16) % 16));
    const uint32_t frame_full = smem_u32(&sh.bar[0]);      // + 8 * slot
    const uint32_t frame_empty = smem_u32(&sh.bar[2]);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kFusedConsumerWarps);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(frame_full + 8 * s, kFusedRows);            // every consumer thread arrives (release of its stores)
            mbar_init(frame_empty + 8 * s, kFusedEnergyThreads);  // every energy thread arrives after reading min/max
        }
        mbar_fence_init();
    }
    load_energy_tables(sh.tab, threadIdx.x, kFusedThreads);
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    unsigned int jitter_counter = 0;

    if (warp == kFusedConsumerWarps) {
        // ------------------------------------ TMA producer -----------------------------------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = l2_evict_first ? l2_policy_evict_first() : 0ull;
            for (unsigned int frame = blockIdx.x; frame < n_frames; frame += gridDim.x)
#pragma unroll 1
                for (int t = 0; t < kFusedTilesPerFrame; ++t)
                    load_tile<kFusedRows, SLABS_PER_STAGE, STAGES>(
                        &tmap, ring, bar_full, bar_empty,
                        static_cast<int32_t>(frame * kFramePixels + t * kFusedRows), policy, stage, phase,
                        JITTER ? jitter_seed : 0u, &jitter_counter);
        }
        return;
    }

    if (warp < kFusedConsumerWarps) {
        // ------------------------------------ MFCC consumers ---------------------------------------
        const uint32_t row_off = threadIdx.x * 128u;
        const uint32_t sw = (threadIdx.x & 7u) << 4;
        // The energy warps read these rows back one frame-time (~80 us) later, by which time 500 MB of spectra have
        // streamed through the 126 MB L2: an evict-last policy keeps the 25 MB of in-flight MFCC frames resident.
        const uint64_t keep = keep_mfcc_in_l2 ? l2_policy_evict_last() : 0ull;
        unsigned int it = 0;
        for (unsigned int frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++it) {
            float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll 1
            for (int t = 0; t < kFusedTilesPerFrame; ++t) {
                float c[12];
                mel_tile<kFusedRows, SLABS_PER_STAGE, STAGES, JITTER>(ring, bar_full, bar_empty, row_off, sw, lane, stage, phase,
                                                                      c, jitter_seed, &jitter_counter);
                const unsigned int p = t * kFusedRows + threadIdx.x;                 // pixel within the frame
                const unsigned int dst = flip180 ? (kFramePixels - 1u - p) : p;
                const size_t dst_row = static_cast<size_t>(frame) * kFramePixels + dst;
                if (keep) {
                    float4* o = reinterpret_cast<float4*>(mfcc_out + dst_row * 12u);
                    stg128_hint(o, make_float4(c[0], c[1], c[2], c[3]), keep);
                    stg128_hint(o + 1, make_float4(c[4], c[5], c[6], c[7]), keep);
                    stg128_hint(o + 2, make_float4(c[8], c[9], c[10], c[11]), keep);
                } else {
                    store_cepstra(mfcc_out, dst_row, c);
                }
#pragma unroll
                for (int m = 0; m < 12; ++m) { mn = fminf(mn, c[m]); mx = fmaxf(mx, c[m]); }
            }
            mn = warp_min(mn);
            mx = warp_max(mx);
            const int slot = it & 1;
            if (JITTER) jitter_spin(jitter_seed, 4u, jitter_counter);
            mbar_wait(frame_empty + 8 * slot, ((it >> 1) & 1u) ^ 1u);    // slot consumed by the energy warps
            if (lane == 0) { sh.minmax[slot][warp][0] = mn; sh.minmax[slot][warp][1] = mx; }
            __syncwarp();
            if (JITTER) jitter_spin(jitter_seed, 5u, jitter_counter);
            mbar_arrive(frame_full + 8 * slot);                           // release: MFCC rows + min/max visible
        }
        return;
    }

    // ---------------------------------------- energy warps -----------------------------------------
    const int et = threadIdx.x - (kFusedRows + 32);                     // 0..255
    const uint32_t exp_table = smem_u32(sh.tab.exp2);
    unsigned int it = 0;
    HeatSmem hs = {};
    unsigned int chunk_it = 0;
    const int heat_h = HH ? HH : heat_out.out_h, heat_w = HW ? HW : heat_out.out_w;
    if (HEAT_VEC != 0) {
        unsigned char* heat_base = reinterpret_cast<unsigned char*>(&sh) + ((sizeof(FusedShared) + 15) & ~size_t(15));
        hs = heat_smem_carve(heat_base, heat_stream_layout(heat_h, heat_w, false, kFusedEnergyWarps), heat_h, heat_w);
        heat_taps_init<kFusedEnergyThreads>(hs, heat_h, heat_w, et);
        energy_group_sync();
    }
    for (unsigned int frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++it) {
        const int slot = it & 1;
        if (JITTER) jitter_spin(jitter_seed, 6u, jitter_counter);
        mbar_wait(frame_full + 8 * slot, (it >> 1) & 1u);
        float lo = sh.minmax[slot][0][0], hi = sh.minmax[slot][0][1];
#pragma unroll
        for (int w = 1; w < kFusedConsumerWarps; ++w) {
            lo = fminf(lo, sh.minmax[slot][w][0]);
            hi = fmaxf(hi, sh.minmax[slot][w][1]);
        }
        if (JITTER) jitter_spin(jitter_seed, 7u, jitter_counter);
        mbar_arrive(frame_empty + 8 * slot);
        const FrameNormFast norm(lo, __fsub_rn(hi, lo));
        const float* img = mfcc_out + static_cast<size_t>(frame) * kFrameValues;
#pragma unroll 1
        for (int p = et; p < kFramePixels; p += kFusedEnergyThreads) {
            const float4* src = reinterpret_cast<const float4*>(img + p * kMfccNum);
            const float4 a = __ldcg(src), b = __ldcg(src + 1), c = __ldcg(src + 2);   // L2: written by this SM just now
            float x[kMfccNum] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            unsigned int rare;
            double en = pixel_energy(x, normalize_first != 0, norm, sh.tab, exp_table, rare);
            if (rare) {                                            // the values already loaded, not a second (read-only) load
                const float raw[kMfccNum] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
                en = pixel_energy_plain(raw, nullptr, normalize_first != 0, norm.lo, norm.range);
            }
            if (JITTER && (p & 255) == (et & 255)) jitter_spin(jitter_seed, 8u, jitter_counter);
            sh.map[p] = en;
            if (energy_out != nullptr) energy_out[static_cast<size_t>(frame) * kFramePixels + p] = en;
        }
        energy_group_sync();
        if (mask_out != nullptr || mean_out != nullptr) {
            const double mean = frame_mean(sh.map, sh.part, sh.leaf, &sh.mean, et, kFusedEnergyThreads, [] { energy_group_sync(); });
            if (et == 0 && mean_out != nullptr) mean_out[frame] = mean;
            if (mask_out != nullptr) {
                for (int p = et; p < kFramePixels; p += kFusedEnergyThreads)
                    mask_out[static_cast<size_t>(frame) * kFramePixels + p] = sh.map[p] > mean ? 1 : 0;
            }
        }
        energy_group_sync();      // sh.map is rewritten for the next frame
        if (HEAT_VEC != 0) {
            constexpr int kPerThread = HeatPerThread<kFusedEnergyThreads>::value;
            double e[kPerThread];
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                const int p = et + i * kFusedEnergyThreads;
                e[i] = p < kFramePixels ? sh.map[p] : CUDART_NAN;
            }
            // (heat_phase's first barrier comes after these reads, so no thread starts the next frame's map before them)
            heat_phase<kFusedEnergyThreads, HEAT_VEC == 0 ? 2 : HEAT_VEC, HW, HH>(
                e, hs, sh.red64, sh.red32, heat_h, heat_w,
                heat_out.heat + static_cast<size_t>(frame) * heat_h * heat_w, et, chunk_it, [] { energy_group_sync(); }, [] {});
        }
    }
    if (HEAT_VEC != 0 && lane == 0) bulk_wait_all<0>();        // shared memory must outlive the copies that read it
}

}  // namespace aig
