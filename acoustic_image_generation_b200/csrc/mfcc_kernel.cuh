// Stage 1 kernels: pixel power spectra [n_rows, fft_len] f32 -> MFCC rows [n_rows, mfcc_num] f32.
//
// Replaces get_feats + np.float32 (dataloader/outdoor_data_mfcc.py:851-876, :823) and, on store,
// the 180-degree flip of _parse_sequence (:314-315).
//
// mfcc_banded_kernel - the hot kernel, for the reference tables only.
//   HBM-bound streaming scan: 2 KiB read and 48 B written per spectrum, ~2.6 kFLOP of FP32.
//   * Data movement: a producer warp issues 2-D TMA tile loads (cp.async.bulk.tensor, 128-byte
//     swizzle, L2 evict-first) of [ROWS spectra x 32 bins] slabs into a ring of shared-memory
//     stages guarded by full/empty mbarriers, so each SM keeps STAGES * SLABS * ROWS * 128 B of
//     reads in flight with no register staging.
//   * Compute: one consumer thread per spectrum.  The mel filter bank is banded (<= 2 adjacent
//     triangles per bin, 942 non-zeros), so instead of a 512x24 product the thread executes the
//     generated straight-line "mel program" (mel_program_ref.inc): one FFMA-immediate per
//     non-zero weight into two rotating partial sums per triangle; when a triangle closes it is
//     floored, logged and folded into the 12 cepstral accumulators (DCT * mfnorm * lifter folded
//     into one float32 table).  The thread reads its spectrum with LDS.128; the TMA swizzle puts the
//     eight 16-byte chunks of consecutive spectra on distinct bank groups, so the row-strided
//     reads are conflict-free.
//   * NaN/Inf semantics of the dense reference product are kept: a non-finite bin poisons the
//     whole row (the reference's x * 0 = NaN in every column), which the final fix-up zeroes.
//
// mfcc_generic_kernel - any other tables (other fft_len / filter counts): float64, one warp per
//   spectrum, dense product.  Not tuned; it exists so get_feats stays a drop-in for every caller.
#pragma once

#include <math_constants.h>

#include "aig_common.cuh"

namespace aig {

template <int ROWS, int SLABS_PER_STAGE, int STAGES>
struct MfccPipe {
    static_assert(ROWS % 32 == 0 && ROWS <= 256, "ROWS is the TMA box height (<= 256)");
    static_assert(16 % SLABS_PER_STAGE == 0, "a spectrum is 16 slabs of 32 bins");
    static constexpr int kConsumerWarps = ROWS / 32;
    static constexpr int kThreads = ROWS + 32;               // + one producer warp
    static constexpr int kSlabBytes = ROWS * 128;            // [ROWS x 32] f32, 128B-swizzled
    static constexpr int kStageBytes = SLABS_PER_STAGE * kSlabBytes;
    static constexpr int kStagesPerTile = 16 / SLABS_PER_STAGE;
    static constexpr int kSmemBytes = STAGES * kStageBytes + 2 * STAGES * 8 + 1024;  // + align slack
};

#define AIG_MEL_DECL(f) float m##f##_0 = 0.f, m##f##_1 = 0.f;

// One consumer thread's work on one tile: wait for the tile's ring stages, run the straight-line mel
// program on the thread's spectrum, release the stages, and return the 12 cepstra with the reference's
// NaN/Inf -> 0 fix-up applied.  `stage` / `phase` are the thread's view of the ring position.
// JITTER (debug builds of the fused kernel only): pseudo-random delays before every barrier wait / arrive.
template <int ROWS, int SLABS_PER_STAGE, int STAGES, bool JITTER = false>
__device__ __forceinline__ void mel_tile(uint32_t ring, uint32_t bar_full, uint32_t bar_empty, uint32_t row_off,
                                         uint32_t sw, int lane, int& stage, uint32_t& phase, float (&c)[12],
                                         unsigned int jitter_seed = 0, unsigned int* jitter_counter = nullptr) {
    using P = MfccPipe<ROWS, SLABS_PER_STAGE, STAGES>;
    AIG_MEL_DECL(0) AIG_MEL_DECL(1) AIG_MEL_DECL(2) AIG_MEL_DECL(3) AIG_MEL_DECL(4) AIG_MEL_DECL(5)
    AIG_MEL_DECL(6) AIG_MEL_DECL(7) AIG_MEL_DECL(8) AIG_MEL_DECL(9) AIG_MEL_DECL(10) AIG_MEL_DECL(11)
    AIG_MEL_DECL(12) AIG_MEL_DECL(13) AIG_MEL_DECL(14) AIG_MEL_DECL(15) AIG_MEL_DECL(16) AIG_MEL_DECL(17)
    AIG_MEL_DECL(18) AIG_MEL_DECL(19) AIG_MEL_DECL(20) AIG_MEL_DECL(21) AIG_MEL_DECL(22) AIG_MEL_DECL(23)
    // two interleaved sets of cepstral accumulators (even / odd triangles): partial sums stay half as large, which
    // trims the float32 rounding of the 24-term sums when |MFCC| runs into the hundreds
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, c5 = 0.f;
    float c6 = 0.f, c7 = 0.f, c8 = 0.f, c9 = 0.f, c10 = 0.f, c11 = 0.f;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f, d5 = 0.f;
    float d6 = 0.f, d7 = 0.f, d8 = 0.f, d9 = 0.f, d10 = 0.f, d11 = 0.f;
    float poison = 0.f;   // 0, or NaN once any bin of the row is NaN/Inf (reference: x * 0 = NaN)

#define MEL_SLAB_BEGIN(s)                                                                        \
    {                                                                                            \
        if (JITTER && (s) % SLABS_PER_STAGE == 0) jitter_spin(jitter_seed, 1u, *jitter_counter); \
        if ((s) % SLABS_PER_STAGE == 0) mbar_wait(bar_full + 8 * stage, phase);                  \
        const uint32_t slab = ring + stage * P::kStageBytes +                                    \
                              ((s) % SLABS_PER_STAGE) * P::kSlabBytes + row_off;
#define MEL_LOAD(s, j) const float4 v##j = lds128(slab + ((static_cast<uint32_t>(j) << 4) ^ sw));
#define MEL_BIN0(j, c) poison = fmaf(v##j.c, 0.f, poison);
#define MEL_BIN1(j, c, f, p, w) m##f##_##p = fmaf(v##j.c, w, m##f##_##p);
#define MEL_BIN2(j, c, f, p, w, g, q, u) MEL_BIN1(j, c, f, p, w) MEL_BIN1(j, c, g, q, u)
#define MEL_DONE(f, w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11)                            \
        {                                                                                        \
            float e = m##f##_0 + m##f##_1;                                                       \
            /* floor at 0.001 (:858); -Inf would be floored away here but is NaN in the        */\
            /* reference's other columns, so turn it into NaN                                   */\
            e = (e < 0.001f) ? ((e == -CUDART_INF_F) ? CUDART_NAN_F : 0.001f) : e;               \
            const float lg = logf(e);                                                            \
            if ((f) % 2 == 0) {                                                                  \
                c0 = fmaf(lg, w0, c0); c1 = fmaf(lg, w1, c1); c2 = fmaf(lg, w2, c2);            \
                c3 = fmaf(lg, w3, c3); c4 = fmaf(lg, w4, c4); c5 = fmaf(lg, w5, c5);            \
                c6 = fmaf(lg, w6, c6); c7 = fmaf(lg, w7, c7); c8 = fmaf(lg, w8, c8);            \
                c9 = fmaf(lg, w9, c9); c10 = fmaf(lg, w10, c10); c11 = fmaf(lg, w11, c11);      \
            } else {                                                                             \
                d0 = fmaf(lg, w0, d0); d1 = fmaf(lg, w1, d1); d2 = fmaf(lg, w2, d2);            \
                d3 = fmaf(lg, w3, d3); d4 = fmaf(lg, w4, d4); d5 = fmaf(lg, w5, d5);            \
                d6 = fmaf(lg, w6, d6); d7 = fmaf(lg, w7, d7); d8 = fmaf(lg, w8, d8);            \
                d9 = fmaf(lg, w9, d9); d10 = fmaf(lg, w10, d10); d11 = fmaf(lg, w11, d11);      \
            }                                                                                    \
        }
#define MEL_SLAB_END(s)                                                                          \
        if ((s) % SLABS_PER_STAGE == SLABS_PER_STAGE - 1) {                                      \
            if (JITTER) jitter_spin(jitter_seed, 2u, *jitter_counter);                           \
            __syncwarp();                                                                        \
            if (lane == 0) mbar_arrive(bar_empty + 8 * stage);                                   \
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }                                   \
        }                                                                                        \
    }
#include "mel_program_ref.inc"
#undef MEL_SLAB_BEGIN
#undef MEL_LOAD
#undef MEL_BIN0
#undef MEL_BIN1
#undef MEL_BIN2
#undef MEL_DONE
#undef MEL_SLAB_END

    const float raw[12] = {c0 + d0, c1 + d1, c2 + d2, c3 + d3, c4 + d4, c5 + d5,
                           c6 + d6, c7 + d7, c8 + d8, c9 + d9, c10 + d10, c11 + d11};
#pragma unroll
    for (int m = 0; m < 12; ++m) {
        const float v = raw[m] + poison;
        c[m] = (fabsf(v) <= 3.402823466e38f) ? v : 0.f;     // NaN / Inf -> 0 (:871-872)
    }
}

// The producer lane's work on one tile: stream its 16 slabs through the ring.
template <int ROWS, int SLABS_PER_STAGE, int STAGES>
__device__ __forceinline__ void load_tile(const CUtensorMap* tmap, uint32_t ring, uint32_t bar_full,
                                          uint32_t bar_empty, int32_t row0, uint64_t policy, int& stage,
                                          uint32_t& phase, unsigned int jitter_seed = 0,
                                          unsigned int* jitter_counter = nullptr) {
    using P = MfccPipe<ROWS, SLABS_PER_STAGE, STAGES>;
#pragma unroll 1
    for (int kb = 0; kb < P::kStagesPerTile; ++kb) {
        if (jitter_seed != 0u) jitter_spin(jitter_seed, 3u, *jitter_counter);
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * stage, P::kStageBytes);
#pragma unroll
        for (int s = 0; s < SLABS_PER_STAGE; ++s)
            tma_load_2d(ring + stage * P::kStageBytes + s * P::kSlabBytes, tmap, bar_full + 8 * stage,
                        (kb * SLABS_PER_STAGE + s) * 32, row0, policy);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
}

__device__ __forceinline__ void store_cepstra(float* __restrict__ out, size_t dst_row, const float (&c)[12]) {
    float4* o = reinterpret_cast<float4*>(out + dst_row * 12u);
    o[0] = make_float4(c[0], c[1], c[2], c[3]);
    o[1] = make_float4(c[4], c[5], c[6], c[7]);
    o[2] = make_float4(c[8], c[9], c[10], c[11]);
}

template <int ROWS, int SLABS_PER_STAGE, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(ROWS + 32, MIN_CTAS)
mfcc_banded_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ out,
                   unsigned int n_rows, unsigned int n_tiles, int flip180, unsigned int frame_pixels,
                   int l2_evict_first) {
    using P = MfccPipe<ROWS, SLABS_PER_STAGE, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle atoms are 1024 B
    const uint32_t bar_full = ring + STAGES * P::kStageBytes;
    const uint32_t bar_empty = bar_full + STAGES * 8;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);                       // producer's expect_tx arrive
            mbar_init(bar_empty + 8 * s, P::kConsumerWarps);      // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    if (warp == P::kConsumerWarps) {
        // ---------------- producer: one lane streams tiles through the ring ----------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = l2_evict_first ? l2_policy_evict_first() : 0ull;
            for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                load_tile<ROWS, SLABS_PER_STAGE, STAGES>(&tmap, ring, bar_full, bar_empty,
                                                         static_cast<int32_t>(tile * ROWS), policy, stage, phase);
        }
        return;
    }

    // -------------------- consumers: one thread per spectrum of the tile --------------------
    const uint32_t row_off = threadIdx.x * 128u;
    const uint32_t sw = (threadIdx.x & 7u) << 4;    // 128B swizzle: 16-byte chunk index ^= row % 8
    for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float c[12];
        mel_tile<ROWS, SLABS_PER_STAGE, STAGES>(ring, bar_full, bar_empty, row_off, sw, lane, stage, phase, c);
        const unsigned int row = tile * ROWS + threadIdx.x;
        if (row < n_rows) {
            unsigned int dst = row;
            if (flip180) {
                const unsigned int frame = row / frame_pixels;
                dst = frame * frame_pixels + (frame_pixels - 1u - (row - frame * frame_pixels));
            }
            store_cepstra(out, dst, c);
        }
    }
}

#undef AIG_MEL_DECL

// ---------------------------------------------------------------------------------------------
// Generic tables: float64, one warp per spectrum, lanes own filters (and later coefficients).
// bank [fft_len, filter_num], dct [filter_num, mfcc_num], lifter [mfcc_num] in global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kGenericMaxFilters = 64;
constexpr int kGenericWarps = 8;

__global__ void __launch_bounds__(kGenericWarps * 32)
mfcc_generic_kernel(const float* __restrict__ power, long long n_rows, int fft_len, int filter_num,
                    int mfcc_num, const double* __restrict__ bank, const double* __restrict__ dct,
                    const double* __restrict__ lifter, double mfnorm, float* __restrict__ out,
                    int flip180, long long frame_pixels) {
    __shared__ double logmel[kGenericWarps][kGenericMaxFilters];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long warps_total = static_cast<long long>(gridDim.x) * kGenericWarps;
    for (long long row = static_cast<long long>(blockIdx.x) * kGenericWarps + warp; row < n_rows;
         row += warps_total) {
        const float* x = power + row * fft_len;
        double acc0 = 0.0, acc1 = 0.0;
        const bool has0 = lane < filter_num, has1 = lane + 32 < filter_num;
        for (int k = 0; k < fft_len; ++k) {
            const double xv = static_cast<double>(__ldg(x + k));        // warp-uniform broadcast load
            const double* w = bank + static_cast<long long>(k) * filter_num;
            if (has0) acc0 = fma(xv, w[lane], acc0);
            if (has1) acc1 = fma(xv, w[lane + 32], acc1);
        }
        if (has0) logmel[warp][lane] = log(acc0 < 0.001 ? 0.001 : acc0);
        if (has1) logmel[warp][lane + 32] = log(acc1 < 0.001 ? 0.001 : acc1);
        __syncwarp();
        long long dst = row;
        if (flip180) {
            const long long frame = row / frame_pixels;
            dst = frame * frame_pixels + (frame_pixels - 1 - (row - frame * frame_pixels));
        }
        for (int m = lane; m < mfcc_num; m += 32) {
            double c = 0.0;
            for (int f = 0; f < filter_num; ++f) c = fma(logmel[warp][f], dct[f * mfcc_num + m], c);
            c *= mfnorm;
            c *= lifter[m];
            if (!(fabs(c) <= 1.7976931348623157e308)) c = 0.0;
            out[dst * mfcc_num + m] = static_cast<float>(c);
        }
        __syncwarp();
    }
}

}  // namespace aig
