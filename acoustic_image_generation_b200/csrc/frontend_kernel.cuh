// "Next" rows N1 and N2: the callers' arithmetic on either side of the hot path.
//
// N1  spectrum_kernel   front half of _build_spectrograms_function (dataloader/outdoor_data_mfcc.py:796-805):
//                       Tukey window * audio -> 1024-point real FFT -> drop Nyquist -> squared magnitude,
//                       float64 like NumPy, producing the [n, 512] power rows the MFCC kernel consumes.
//     filtfilt_kernel   butter_lowpass_filter (:565-575): scipy.signal.filtfilt of an order-10 Butterworth -
//                       odd extension, forward/backward direct-form-II-transposed IIR with lfilter_zi initial
//                       state, float64, in scipy's operation order (no FMA contraction: the order-10 transfer
//                       function is ill-conditioned and amplifies any reordering).
// N2  normalize_mfcc_kernel   _normalize_mfcc (:696-703): per-vector float32 min-max of the 12 MFCCs.
//     tile_mfcc_kernel        mfccmap = tile(reshape(mfcc, (-1,1,12)), (1, 36*48, 1)) (trainer/mfcctrainer.py:38-40,
//                             iouenergythreshold.py:99-101): the [B,36,48,12] conditioning image of the UNet.
//     split_triplets_kernel, triplet_mse_kernel   the four channel-triplet slices and the five MSE terms of the
//                             trainers (trainer/mfcctrainer.py:103-117).
#pragma once

#include <type_traits>

#include "aig_common.cuh"

namespace aig {

constexpr int kAudioSamples = 1024;     // _NUMBER_OF_SAMPLES (outdoor_data_mfcc.py:9)
constexpr int kSpectrumThreads = 256;

// ---- N1: windowed power spectrum --------------------------------------------------------------------
// twiddle[k] = exp(-2*pi*i*k/1024), k < 512, float64, built on the host.
template <typename In>
__global__ void __launch_bounds__(kSpectrumThreads)
spectrum_kernel(const In* __restrict__ audio, long long n_rows, const double* __restrict__ window,
                const double2* __restrict__ twiddle, float* __restrict__ power) {
    __shared__ double s_re[kAudioSamples];
    __shared__ double s_im[kAudioSamples];
    const int tid = threadIdx.x;
    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        // load in bit-reversed order (decimation in time), real input
        for (int i = tid; i < kAudioSamples; i += kSpectrumThreads) {
            double v = static_cast<double>(audio[row * kAudioSamples + i]);
            if (window != nullptr) v = __dmul_rn(v, window[i]);
            const int r = __brev(static_cast<unsigned>(i)) >> 22;     // 10-bit reversal
            s_re[r] = v;
            s_im[r] = 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int half = 1; half < kAudioSamples; half <<= 1) {        // 10 radix-2 stages
            const int step = (kAudioSamples / 2) / half;              // twiddle stride
            for (int b = tid; b < kAudioSamples / 2; b += kSpectrumThreads) {
                const int j = b & (half - 1);
                const int lo = ((b - j) << 1) + j, hi = lo + half;
                const double2 w = twiddle[j * step];
                const double xr = s_re[hi], xi = s_im[hi];
                const double tr = xr * w.x - xi * w.y, ti = xr * w.y + xi * w.x;
                const double ur = s_re[lo], ui = s_im[lo];
                s_re[lo] = ur + tr; s_im[lo] = ui + ti;
                s_re[hi] = ur - tr; s_im[hi] = ui - ti;
            }
            __syncthreads();
        }
        // np.abs(rfft)[:, :-1] ** 2 (:802-803): bins 0..511
        for (int k = tid; k < kAudioSamples / 2; k += kSpectrumThreads) {
            const double mag = hypot(s_re[k], s_im[k]);
            power[row * (kAudioSamples / 2) + k] = static_cast<float>(__dmul_rn(mag, mag));
        }
    }
}

// ---- N1: zero-phase IIR (scipy.signal.filtfilt, padtype='odd', method='pad') -------------------------
constexpr int kMaxTaps = 16;

// One thread per signal row.  scratch [n_rows, length + 2 * pad] float64 holds the forward pass.
template <typename In>
__global__ void filtfilt_kernel(const In* __restrict__ x, long long n_rows, int length, int ntaps, int pad,
                                const double* __restrict__ coef /* b[ntaps] | a[ntaps] | zi[ntaps-1] */,
                                double* __restrict__ scratch, float* __restrict__ y_out) {
    const long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double b[kMaxTaps], a[kMaxTaps], z[kMaxTaps];
    for (int i = 0; i < kMaxTaps; ++i) { b[i] = 0.0; a[i] = 0.0; z[i] = 0.0; }
    const double a0 = coef[ntaps];
    for (int i = 0; i < ntaps; ++i) { b[i] = __ddiv_rn(coef[i], a0); a[i] = __ddiv_rn(coef[ntaps + i], a0); }
    const In* xr = x + row * length;
    double* fwd = scratch + row * (length + 2 * pad);
    const int total = length + 2 * pad;
    const double first = static_cast<double>(xr[0]), last = static_cast<double>(xr[length - 1]);
    // odd extension: 2*x[0] - x[pad..1], x, 2*x[-1] - x[-2..-(pad+1)].  scipy builds it in the INPUT dtype before the
    // filter promotes to float64: float32 arithmetic for float32 audio, exact integers for the int32 samples of the
    // tfrecords (|x| < 2^30 assumed).
    auto reflect = [&](double end, In v) -> double {
        if (std::is_same<In, float>::value)
            return static_cast<double>(__fsub_rn(__fmul_rn(2.f, static_cast<float>(end)), static_cast<float>(v)));
        return __dadd_rn(__dmul_rn(2.0, end), -static_cast<double>(v));
    };
    auto ext = [&](int i) -> double {
        if (i < pad) return reflect(first, xr[pad - i]);
        if (i < pad + length) return static_cast<double>(xr[i - pad]);
        return reflect(last, xr[length - 2 - (i - pad - length)]);
    };
    // scipy's direct form II transposed step: y = z0 + b0*x; z[i] = z[i+1] + x*b[i+1] - y*a[i+1]; z[last] = x*b[n-1] - y*a[n-1]
    auto step = [&](double xn) -> double {
        const double yn = __dadd_rn(z[0], __dmul_rn(b[0], xn));
#pragma unroll
        for (int i = 0; i < kMaxTaps - 2; ++i)
            if (i < ntaps - 2)
                z[i] = __dadd_rn(__dadd_rn(z[i + 1], __dmul_rn(xn, b[i + 1])), -__dmul_rn(yn, a[i + 1]));
        z[ntaps - 2] = __dadd_rn(__dmul_rn(xn, b[ntaps - 1]), -__dmul_rn(yn, a[ntaps - 1]));
        return yn;
    };
    const double x0 = ext(0);
    for (int i = 0; i < ntaps - 1; ++i) z[i] = __dmul_rn(coef[2 * ntaps + i], x0);      // zi * ext[0]
    for (int i = 0; i < total; ++i) fwd[i] = step(ext(i));
    const double y0 = fwd[total - 1];
    for (int i = 0; i < ntaps - 1; ++i) z[i] = __dmul_rn(coef[2 * ntaps + i], y0);      // zi * y[-1]
    for (int i = total - 1; i >= 0; --i) {
        const double v = step(fwd[i]);
        if (i >= pad && i < pad + length) y_out[row * length + (i - pad)] = static_cast<float>(v);   // np.float32(y) (:574)
    }
}

// ---- N2: per-vector min-max and the tiled conditioning image ------------------------------------------
// tf.reduce_min / reduce_max propagate NaN (outdoor_data_mfcc.py:699,701): one NaN makes the whole vector NaN.
__device__ __forceinline__ void minmax_normalize12(float (&v)[12]) {
    float mn = v[0], mx = v[0];
    bool nan = v[0] != v[0];
#pragma unroll
    for (int m = 1; m < 12; ++m) { mn = fminf(mn, v[m]); mx = fmaxf(mx, v[m]); nan |= v[m] != v[m]; }
    if (nan) mn = CUDART_NAN_F;
    const float range = __fsub_rn(mx, mn);          // max(x - min) == fl(max - min)
#pragma unroll
    for (int m = 0; m < 12; ++m) v[m] = __fdiv_rn(__fsub_rn(v[m], mn), range);
}

__global__ void normalize_mfcc_kernel(const float* __restrict__ mfcc, long long n, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* src = reinterpret_cast<const float4*>(mfcc + i * 12);
    const float4 a = src[0], b = src[1], c = src[2];
    float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    minmax_normalize12(v);
    float4* dst = reinterpret_cast<float4*>(out + i * 12);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    dst[2] = make_float4(v[8], v[9], v[10], v[11]);
}

constexpr int kTileThreads = 192;       // a multiple of 3: every thread always writes the same third of the vector
__global__ void __launch_bounds__(kTileThreads)
tile_mfcc_kernel(const float* __restrict__ mfcc, long long n, int normalize, float* __restrict__ map_out) {
    for (long long f = blockIdx.x; f < n; f += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(mfcc + f * 12);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        if (normalize) minmax_normalize12(v);
        const int third = threadIdx.x % 3;
        const float4 mine = third == 0 ? make_float4(v[0], v[1], v[2], v[3])
                          : third == 1 ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(v[8], v[9], v[10], v[11]);
        float4* dst = reinterpret_cast<float4*>(map_out + f * kFrameValues);
#pragma unroll 3
        for (int i = threadIdx.x; i < kFrameValues / 4; i += kTileThreads) __stcs(dst + i, mine);
    }
}

// (A bulk-copy form - a warp builds 72 repetitions in shared memory and ships the chunk 24 times - was measured against
// this kernel in round 2: identical, 5.1 TB/s at 4096 frames and 6.0 TB/s = 0.95 of the write roof from 16 k frames up,
// tools/tile_probe.py; the stores are not what limits it, the short launch is.)

// ---- N2: channel-triplet slices and their losses (trainer/mfcctrainer.py:103-117) -----------------------
// The trainers cut both the target and the generated [n,36,48,12] image into the four channel triplets
// tf.slice(x, [0,0,0,3t], [-1,36,48,3]) and take tf.losses.mean_squared_error of the whole image and of each pair of
// triplets.  split_triplets_kernel writes the four slices as contiguous [n,36,48,3] tensors (what tf.slice returns);
// triplet_mse_kernel reads both images once and produces all five means.
constexpr int kTripletThreads = 256;
constexpr int kTripletTile = 256;        // pixels per tile
__global__ void __launch_bounds__(kTripletThreads)
split_triplets_kernel(const float* __restrict__ images, long long n_pixels, float* __restrict__ out) {
    // Tile of 256 pixels: coalesced float4 loads into a channel-major shared tile (row stride 257: conflict-free both
    // ways), then each triplet's 768 floats leave as 192 coalesced float4 stores.  n_pixels is a multiple of 1728, so
    // of 4; the last tile may be short.
    __shared__ float s_tile[12][kTripletTile + 1];
    const long long tiles = (n_pixels + kTripletTile - 1) / kTripletTile;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long first = tile * kTripletTile;
        const int count = static_cast<int>(min(static_cast<long long>(kTripletTile), n_pixels - first));
        const float4* src = reinterpret_cast<const float4*>(images + first * 12);
        __syncthreads();
        for (int i = threadIdx.x; i < count * 3; i += kTripletThreads) {
            const float4 q = __ldcs(src + i);
            const int px = i / 3, c = 4 * (i - 3 * px);
            s_tile[c][px] = q.x; s_tile[c + 1][px] = q.y; s_tile[c + 2][px] = q.z; s_tile[c + 3][px] = q.w;
        }
        __syncthreads();
        const int quads = count * 3 / 4;                     // float4 per triplet in this tile
        for (int i = threadIdx.x; i < 4 * quads; i += kTripletThreads) {
            const int t = i / quads, j = i - t * quads;
            float w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o = 4 * j + k, px = o / 3;
                w[k] = s_tile[3 * t + (o - 3 * px)][px];
            }
            __stcs(reinterpret_cast<float4*>(out + (t * n_pixels + first) * 3) + j, make_float4(w[0], w[1], w[2], w[3]));
        }
    }
}

// partial[block][4]: float64 sums of the float32 squared differences per triplet, in a fixed order (deterministic).
__global__ void __launch_bounds__(kTripletThreads)
triplet_mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n_pixels,
                   double* __restrict__ partial) {
    __shared__ double s_sum[kTripletThreads / 32][4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long px = blockIdx.x * static_cast<long long>(kTripletThreads) + threadIdx.x; px < n_pixels;
         px += static_cast<long long>(gridDim.x) * kTripletThreads) {
        const float4* pa = reinterpret_cast<const float4*>(a) + px * 3;
        const float4* pb = reinterpret_cast<const float4*>(b) + px * 3;
        float x[12], y[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 q = __ldcs(pa + i), r = __ldcs(pb + i);
            x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
            y[4 * i] = r.x; y[4 * i + 1] = r.y; y[4 * i + 2] = r.z; y[4 * i + 3] = r.w;
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const float d = __fsub_rn(x[c], y[c]);
            acc[c / 3] += static_cast<double>(__fmul_rn(d, d));      // squared difference in float32, like TF
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        double v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_sum[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int w = 0; w < kTripletThreads / 32; ++w) v += s_sum[w][threadIdx.x];
        partial[blockIdx.x * 4 + threadIdx.x] = v;
    }
}

// mse_out[0] = whole image, mse_out[1 + t] = triplet t
__global__ void triplet_mse_finish_kernel(const double* __restrict__ partial, int n_blocks, long long n_pixels,
                                          double* __restrict__ mse_out) {
    __shared__ double s_total[4];
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int blk = 0; blk < n_blocks; ++blk) v += partial[blk * 4 + threadIdx.x];
        s_total[threadIdx.x] = v;
        mse_out[1 + threadIdx.x] = v / static_cast<double>(n_pixels * 3);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        mse_out[0] = (s_total[0] + s_total[1] + s_total[2] + s_total[3]) / static_cast<double>(n_pixels * 12);
}

// ---- N4: heat-map overlay (showvideo.py:217-233, showimages.py:144-150) ----------------------------------
// gray = cv2.cvtColor(frame, COLOR_BGR2GRAY) (OpenCV 4.x: 15-bit fixed point, (3735 B + 19235 G + 9798 R + 2^14) >> 15), shown with imshow(cmap=gray) - i.e. min/max
// normalised and quantised to 256 levels - then imshow(map, cmap=jet, alpha=0.7): the normalised heat map goes through
// matplotlib's 256-entry jet table and is alpha-blended over the gray image.  One CTA per frame; out is RGB8.
// jet_lut: 256 x 3 uint8.  Without a frame (bgr == nullptr) the colour-mapped heat map alone is written.
// Per CTA: s_jet[j] = jet[j] * alpha (or jet[j] itself without a frame), eight conflict-free copies; per frame: the gray
// level of a luma is one multiply-high by the frame's reciprocal (GrayLevels) - the blend of a pixel is then one table row
// plus one product, three adds, three roundings, with the reference's operation order (products rounded separately, then added).
// VEC = 4: four pixels per thread (float4 heat, 3 x 32-bit BGR in, 3 x 32-bit RGB out); needs n_pixels % 4 == 0 and
// 16 / 4 / 4-byte aligned heat / bgr / out.  VEC = 1: any geometry.
__device__ __forceinline__ int luma15(unsigned b, unsigned g, unsigned r) {
    return static_cast<int>((b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
}

// imshow(gray, cmap=gray): level = min(int(fl(fl((y - lo) / (hi - lo)) * 256)), 255).  With a = y - lo <= b = hi - lo <= 255 the
// float32 quotient can only land on the other side of a multiple of 1/256 if 256 a / b is within 2^-24 of an integer without
// being one, and its distance to the next integer is at least 1 / 255: the level IS the integer quotient 256 a / b, which
// one multiply-high by a per-frame reciprocal returns exactly (n = 256 a <= 65 280, M = floor(2^32 / b) + 1 overshoots
// 2^32 / b by e / b with e <= b, and n e < 2^32).  b = 1: M = 2^32 - 1 gives 0 and 255.  b = 0 (constant frame): level 0.
struct GrayLevels {
    unsigned lo8, m;          // lo << 8, reciprocal
    float keep;               // 1 - alpha
    __device__ __forceinline__ GrayLevels(int lo, int hi, float keep_) : lo8(static_cast<unsigned>(lo) << 8), keep(keep_) {
        const unsigned b = static_cast<unsigned>(hi - lo);
        m = b == 0u ? 0u : b == 1u ? 0xffffffffu : 0xffffffffu / b + 1u;      // a power of two gets 2^32 / b itself (e = 0), every other b floor(2^32 / b) + 1
    }
    // y8 = luma << 8; returns fl(level * (1 - alpha))
    __device__ __forceinline__ float operator()(unsigned y8) const {
        return __fmul_rn(static_cast<float>(min(__umulhi(y8 - lo8, m), 255u)), keep);
    }
};

// Pixel quads ahead (per thread, in strides of the CTA) whose lines are asked into L2 (prefetch.global.L2) while the current
// ones are worked on; the first kOverlayPrefetch quads of the heat map are asked for at the end of the luma pass.  The kernel
// was short of bytes in flight, not of bandwidth or issue slots (ncu: DRAM 55 %, issue slots 54 %, L1 49 %, the same time
// with and without the second BGR read): 7.3 -> 9.3 M frames/s at 224 x 298 (8 192 frames; 4 / 8 / 16 ahead: 9.20 / 9.26 / 8.95).
constexpr int kOverlayPrefetch = 8;

// the low bytes of four words as one word (three byte permutes)
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

constexpr size_t kOverlayLumaMaxBytes = 80 * 1024;    // luma plane + 32 KB of colour-table copies + 1 KB reserved, twice, in 228 KB
constexpr int kOverlayThreads = 512;     // 2 CTAs per SM

// LUMA (VEC = 4 with a frame only): the first pass leaves the four lumas of every pixel quad as one 32-bit word in dynamic
// shared memory (n_pixels bytes: 65 KB at 224 x 298, two CTAs per SM; up to 80 KB), so the second pass neither reads the BGR frame
// again (ncu, round 2: 867 KB of DRAM traffic per frame against 667 KB algorithmic - the re-read missed L2) nor repeats
// the luma arithmetic.
template <int VEC, bool LUMA = false>
__global__ void __launch_bounds__(kOverlayThreads, 2)
overlay_kernel(const float* __restrict__ heat, const uint8_t* __restrict__ bgr, long long n_frames, int n_pixels,
               float alpha, const uint8_t* __restrict__ jet_lut, uint8_t* __restrict__ rgb_out) {
    static_assert(!LUMA || VEC == 4, "the luma plane is stored per pixel quad");
    extern __shared__ __align__(16) uint32_t s_luma[];       // LUMA: [n_pixels / 4]
    // eight copies of the colour table, one per lane of a quarter warp: the 128-bit look-ups of a warp are free of bank
    // conflicts whatever the heat values (ncu, round 2, one copy + a 256-entry gray table: the L1 data pipe 87 % busy, 29 k
    // shared-memory wavefronts per frame where 8 k are the minimum - that pipe, not DRAM, was the bound)
    __shared__ float4 s_jet[256][8];
    __shared__ int s_red[2][kOverlayThreads / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, copy = tid & 7;
    for (int i = tid; i < 256 * 8; i += kOverlayThreads) {
        const int j = i >> 3;
        const float r = jet_lut[3 * j], g = jet_lut[3 * j + 1], b = jet_lut[3 * j + 2];
        s_jet[j][i & 7] = bgr ? make_float4(__fmul_rn(r, alpha), __fmul_rn(g, alpha), __fmul_rn(b, alpha), 0.f)
                              : make_float4(r, g, b, 0.f);
    }
    const float keep = __fsub_rn(1.f, alpha);
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        __syncthreads();
        const uint8_t* img = bgr ? bgr + f * n_pixels * 3 : nullptr;
        int lo = 255, hi = 255;
        if (img != nullptr) {
            hi = 0;
            if (VEC == 4) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(img);
#pragma unroll 4
                for (int q = tid; q < n_pixels / 4; q += kOverlayThreads) {
                    if (kOverlayPrefetch && q + kOverlayPrefetch * kOverlayThreads < n_pixels / 4) prefetch_l2(w + 3 * (q + kOverlayPrefetch * kOverlayThreads));
                    const uint32_t w0 = __ldg(w + 3 * q), w1 = __ldg(w + 3 * q + 1), w2 = __ldg(w + 3 * q + 2);
                    const int y0 = luma15(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
                    const int y1 = luma15(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
                    const int y2 = luma15((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
                    const int y3 = luma15((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
                    lo = min(min(lo, min(y0, y1)), min(y2, y3));
                    hi = max(max(hi, max(y0, y1)), max(y2, y3));
                    if (LUMA) s_luma[q] = static_cast<uint32_t>(y0 | (y1 << 8) | (y2 << 16) | (y3 << 24));
                }
            } else {
                for (int p = tid; p < n_pixels; p += kOverlayThreads) {
                    const int y = luma15(img[3 * p], img[3 * p + 1], img[3 * p + 2]);
                    lo = min(lo, y); hi = max(hi, y);
                }
            }
            if (kOverlayPrefetch && VEC == 4) {
                const float4* h4 = reinterpret_cast<const float4*>(heat + f * n_pixels);
#pragma unroll
                for (int u = 0; u < kOverlayPrefetch; ++u)
                    if (tid + u * kOverlayThreads < n_pixels / 4) prefetch_l2(h4 + tid + u * kOverlayThreads);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
            __syncthreads();
            lo = s_red[0][0]; hi = s_red[1][0];
#pragma unroll
            for (int w = 1; w < kOverlayThreads / 32; ++w) { lo = min(lo, s_red[0][w]); hi = max(hi, s_red[1][w]); }
        }
        const GrayLevels gray(lo, hi, keep);
        const float* hp = heat + f * n_pixels;
        uint8_t* op = rgb_out + f * n_pixels * 3;
        if (VEC == 4) {
            const float4* h4 = reinterpret_cast<const float4*>(hp);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(img);
            uint32_t* o = reinterpret_cast<uint32_t*>(op);
            const int nq = n_pixels / 4;
            // Four pixel quads per thread and round, all their loads issued before the first is used: with one quad per round
            // the kernel sat on global-memory latency (ncu, round 2: 20 of 26 warp-cycles per issue on the long scoreboard,
            // 0.55 of the DRAM roof with 55 resident warps per SM).
            constexpr int U = 4;                  // (six quads for the luma form, whose quads need one word instead of three: 6.61 M frames/s against 6.89 M)
            for (int q0 = tid; q0 < nq; q0 += U * kOverlayThreads) {
                float4 hv[U];
                uint32_t wv[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = q0 + u * kOverlayThreads;
                    if (kOverlayPrefetch && q + kOverlayPrefetch * kOverlayThreads < nq) prefetch_l2(h4 + q + kOverlayPrefetch * kOverlayThreads);
                    if (q < nq) {
                        hv[u] = __ldcs(h4 + q);
                        if (LUMA) wv[u][0] = s_luma[q];
                        else if (img != nullptr) { wv[u][0] = __ldg(w + 3 * q); wv[u][1] = __ldg(w + 3 * q + 1); wv[u][2] = __ldg(w + 3 * q + 2); }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = q0 + u * kOverlayThreads;
                    if (q >= nq) break;
                    const float hs[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
                    float gk[4] = {0.f, 0.f, 0.f, 0.f};
                    if (LUMA) {
                        const uint32_t y = wv[u][0];
                        gk[0] = gray((y << 8) & 0xff00u); gk[1] = gray(y & 0xff00u); gk[2] = gray((y >> 8) & 0xff00u); gk[3] = gray((y >> 16) & 0xff00u);
                    } else if (img != nullptr) {
                        const uint32_t w0 = wv[u][0], w1 = wv[u][1], w2 = wv[u][2];
                        gk[0] = gray(static_cast<unsigned>(luma15(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u)) << 8);
                        gk[1] = gray(static_cast<unsigned>(luma15(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u)) << 8);
                        gk[2] = gray(static_cast<unsigned>(luma15((w1 >> 16) & 255u, w1 >> 24, w2 & 255u)) << 8);
                        gk[3] = gray(static_cast<unsigned>(luma15((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24)) << 8);
                    }
                    uint32_t c[12];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ji = min(max(static_cast<int>(__fmul_rn(hs[k], 256.f)), 0), 255);   // Colormap: int(x * N)
                        const float4 jet = s_jet[ji][copy];
                        // rint of a sum in [0, 255] without the conversion unit: + 2^23 leaves it, rounded to nearest even, in the low mantissa byte
                        c[3 * k] = __float_as_uint(__fadd_rn(__fadd_rn(jet.x, gk[k]), 8388608.f));
                        c[3 * k + 1] = __float_as_uint(__fadd_rn(__fadd_rn(jet.y, gk[k]), 8388608.f));
                        c[3 * k + 2] = __float_as_uint(__fadd_rn(__fadd_rn(jet.z, gk[k]), 8388608.f));
                    }
                    __stcs(o + 3 * q, pack_low_bytes(c[0], c[1], c[2], c[3]));
                    __stcs(o + 3 * q + 1, pack_low_bytes(c[4], c[5], c[6], c[7]));
                    __stcs(o + 3 * q + 2, pack_low_bytes(c[8], c[9], c[10], c[11]));
                }
            }
        } else {
            for (int p = tid; p < n_pixels; p += kOverlayThreads) {
                const int ji = min(max(static_cast<int>(__fmul_rn(hp[p], 256.f)), 0), 255);
                const float4 jet = s_jet[ji][copy];
                const float gk = img ? gray(static_cast<unsigned>(luma15(img[3 * p], img[3 * p + 1], img[3 * p + 2])) << 8) : 0.f;
                op[3 * p] = static_cast<uint8_t>(__float2int_rn(__fadd_rn(jet.x, gk)));
                op[3 * p + 1] = static_cast<uint8_t>(__float2int_rn(__fadd_rn(jet.y, gk)));
                op[3 * p + 2] = static_cast<uint8_t>(__float2int_rn(__fadd_rn(jet.z, gk)));
            }
        }
    }
}

}  // namespace aig
