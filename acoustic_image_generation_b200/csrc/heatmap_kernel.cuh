// Heat-map kernels (F6) and the exact-integer mask up-sampling (part of F9).
//
// heatmap_kernel       cv2.resize bilinear + implicit Normalize (showimages.py:147-148) as a float64 replica of the oracle
//                      ("heatmap_exact" mode: bit-equal float32 result).
// heat_stream_kernel   the default float32 path.  <false>: energy maps in, heat maps out (aig_heatmap).  <true>: the whole
//                      per-frame body of showvideo.py:226-228 / showimages.py:146-148 in ONE launch (aig_energy_heatmap):
//                      find_logen (+ mean mask) with the two-threads-per-pixel float64 code of energy_kernel.cuh, then
//                      up-sampling and normalisation, the energy map never leaving shared memory.  Output rows are staged
//                      in shared memory and leave the SM as bulk asynchronous copies (cp.async.bulk shared -> global, SASS
//                      UBLKCP): one instruction per 2.4 KB instead of one st.global per 8 bytes, which is what took the
//                      round-1 kernel off the issue limit (76 % issue-active, 0.58 of the write roof).
// heatmap_fast_kernel  round-1 float32 kernel with per-thread stores, kept for output shapes whose frames are not
//                      16-byte multiples (odd widths).
// resize_mask_kernel   cv2.resize(mask) > 0.5 in exact integers (showimages_bb.py:303-304).
#pragma once

#include <type_traits>

#include "aig_common.cuh"
#include "energy_kernel.cuh"

namespace aig {

// ---- bilinear taps (cv2.resize INTER_LINEAR: half-pixel centres, border clamp) -------------------
// float64 taps exactly as the oracle forms them: pos = (d + 0.5) * (n_src / n_dst) - 0.5.
__device__ __forceinline__ void linear_tap(int d, int n_src, int n_dst, int* i0, int* i1, double* w1) {
    const double scale = __ddiv_rn(static_cast<double>(n_src), static_cast<double>(n_dst));
    const double pos = __dadd_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), -0.5);
    int lo = static_cast<int>(floor(pos));
    double w = __dadd_rn(pos, -static_cast<double>(lo));
    if (lo < 0) { lo = 0; w = 0.0; }
    if (lo >= n_src - 1) { lo = n_src - 1; w = 0.0; }
    *i0 = lo;
    *i1 = min(lo + 1, n_src - 1);
    *w1 = w;
}
// Integer taps: pos = ((2d+1) * n_src - n_dst) / (2 * n_dst); weight numerator over den = 2 * n_dst.
__device__ __forceinline__ void linear_tap_exact(int d, int n_src, int n_dst, int* i0, int* i1, int* num) {
    const int den = 2 * n_dst;
    const int t = (2 * d + 1) * n_src - n_dst;
    int lo = (t >= 0) ? t / den : -((-t + den - 1) / den);
    int r = t - lo * den;
    if (lo < 0) { lo = 0; r = 0; }
    if (lo >= n_src - 1) { lo = n_src - 1; r = 0; }
    *i0 = lo;
    *i1 = min(lo + 1, n_src - 1);
    *num = r;
}

constexpr int kHeatThreads = 256;
constexpr int kMaxOut = 2048;   // out_h, out_w <= 2048

// energy [n, 36, 48] f64 -> heat [n, out_h, out_w] f32 = (up - min(up)) / (max(up) - min(up)).
// Dynamic shared memory: out_w * (2 int + 1 double) + out_h * (2 int + 1 double).
__global__ void __launch_bounds__(kHeatThreads)
heatmap_kernel(const double* __restrict__ energy, long long n_frames, int out_h, int out_w,
               float* __restrict__ heat) {
    extern __shared__ double s_dyn[];
    __shared__ double s_map[kFramePixels];
    __shared__ double s_red[2][kHeatThreads / 32];
    double* s_wx = s_dyn;                      // [out_w]
    double* s_wy = s_wx + out_w;               // [out_h]
    int* s_x0 = reinterpret_cast<int*>(s_wy + out_h);   // [out_w] x0 | x1 << 16
    int* s_y0 = s_x0 + out_w;                  // [out_h]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int d = tid; d < out_w; d += kHeatThreads) {
        int i0, i1; double w;
        linear_tap(d, kFrameW, out_w, &i0, &i1, &w);
        s_x0[d] = i0 | (i1 << 16); s_wx[d] = w;
    }
    for (int d = tid; d < out_h; d += kHeatThreads) {
        int i0, i1; double w;
        linear_tap(d, kFrameH, out_h, &i0, &i1, &w);
        s_y0[d] = i0 | (i1 << 16); s_wy[d] = w;
    }
    const int n_out = out_h * out_w;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        __syncthreads();
        for (int p = tid; p < kFramePixels; p += kHeatThreads) s_map[p] = energy[frame * kFramePixels + p];
        __syncthreads();
        auto sample = [&](int idx) -> double {
            const int y = idx / out_w, x = idx - y * out_w;
            const int xi = s_x0[x], yi = s_y0[y];
            const int x0 = xi & 0xffff, x1 = xi >> 16, y0 = yi & 0xffff, y1 = yi >> 16;
            const double wx = s_wx[x], wy = s_wy[y];
            const double ux = __dadd_rn(1.0, -wx), uy = __dadd_rn(1.0, -wy);
            // horizontal pass on the two source rows, then the vertical pass (no FMA contraction)
            const double top = __dadd_rn(__dmul_rn(s_map[y0 * kFrameW + x0], ux), __dmul_rn(s_map[y0 * kFrameW + x1], wx));
            const double bot = __dadd_rn(__dmul_rn(s_map[y1 * kFrameW + x0], ux), __dmul_rn(s_map[y1 * kFrameW + x1], wx));
            return __dadd_rn(__dmul_rn(top, uy), __dmul_rn(bot, wy));
        };
        double mn = CUDART_INF, mx = -CUDART_INF;
        for (int idx = tid; idx < n_out; idx += kHeatThreads) {
            const double v = sample(idx);
            mn = fmin(mn, v); mx = fmax(mx, v);
        }
        mn = warp_min(mn); mx = warp_max(mx);
        if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; }
        __syncthreads();
        mn = s_red[0][0]; mx = s_red[1][0];
#pragma unroll
        for (int w = 1; w < kHeatThreads / 32; ++w) { mn = fmin(mn, s_red[0][w]); mx = fmax(mx, s_red[1][w]); }
        const double range = __dadd_rn(mx, -mn);
        float* dst = heat + frame * n_out;
        for (int idx = tid; idx < n_out; idx += kHeatThreads)
            dst[idx] = __double2float_rn(__ddiv_rn(__dadd_rn(sample(idx), -mn), range));
    }
}

// Fast path (default): the same map in float32.  Bilinear interpolation commutes with the affine map
// t = (e - min e) / (max e - min e), so the frame is first normalised to [0, 1] in float64 (1728 values) and everything
// per output pixel - separable bilinear, min/max of the up-sampled image, final normalisation - runs in float32 on
// values of order one: error ~1e-7 of the output range however flat the raw energies are.  The horizontal pass is done
// once into shared memory (36 x out_w), so an output pixel costs two shared loads and three FMAs per pass and the
// kernel approaches the HBM write rate (out_h * out_w * 4 B per frame).
// Dynamic shared memory: 36 * out_w floats (rows) + out_w * (int + float) + out_h * (int + float).
// VEC = pixels per thread per step (4 when out_w % 4 == 0, 2 when even, else 1): the kernel is issue-bound, so the
// vertical passes use 8/16-byte shared loads and global stores.
// Packed float32 pairs (fma.rn.f32x2 / add / mul, sm_100): two IEEE round-to-nearest operations per instruction, bit for
// bit what the scalar forms give - the streaming heat-map kernel is issue-bound, not FLOP-bound.
__device__ __forceinline__ unsigned long long pack2(float2 v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long r) {
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
// ((b - a) * wy + a - mn) * inv per component: the same four roundings as (fmaf(b - a, wy, a) - mn) * inv
__device__ __forceinline__ float2 lerp_norm2(float2 a, float2 b, float wy, float mn, float inv) {
    const unsigned long long pa = pack2(a), pb = pack2(b);
    const unsigned long long neg1 = pack2(make_float2(-1.f, -1.f)), w2 = pack2(make_float2(wy, wy));
    const unsigned long long nmn = pack2(make_float2(-mn, -mn)), i2 = pack2(make_float2(inv, inv));
    unsigned long long d, v, t, o;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pa), "l"(neg1), "l"(pb));       // b - a
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(d), "l"(w2), "l"(pa));          // (b - a) * wy + a
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(v), "l"(nmn));                      // - mn
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(t), "l"(i2));                       // * inv
    return unpack2(o);
}

template <int VEC>
struct HeatVec;
template <>
struct HeatVec<1> {
    using T = float;
    static __device__ __forceinline__ void lerp_minmax(const float* r0, const float* r1, int x, float wy, float& mn, float& mx) {
        const float v = fmaf(r1[x] - r0[x], wy, r0[x]);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    static __device__ __forceinline__ void lerp_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        __stcs(o + x, (fmaf(r1[x] - r0[x], wy, r0[x]) - mn) * inv);
    }
    static __device__ __forceinline__ void norm_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        o[x] = (fmaf(r1[x] - r0[x], wy, r0[x]) - mn) * inv;
    }
};
template <>
struct HeatVec<2> {
    static __device__ __forceinline__ void lerp_minmax(const float* r0, const float* r1, int x, float wy, float& mn, float& mx) {
        const float2 a = *reinterpret_cast<const float2*>(r0 + x), b = *reinterpret_cast<const float2*>(r1 + x);
        const float v0 = fmaf(b.x - a.x, wy, a.x), v1 = fmaf(b.y - a.y, wy, a.y);
        mn = fminf(mn, fminf(v0, v1)); mx = fmaxf(mx, fmaxf(v0, v1));
    }
    static __device__ __forceinline__ void lerp_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        const float2 a = *reinterpret_cast<const float2*>(r0 + x), b = *reinterpret_cast<const float2*>(r1 + x);
        __stcs(reinterpret_cast<float2*>(o + x), make_float2((fmaf(b.x - a.x, wy, a.x) - mn) * inv, (fmaf(b.y - a.y, wy, a.y) - mn) * inv));
    }
    static __device__ __forceinline__ void norm_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        const float2 a = *reinterpret_cast<const float2*>(r0 + x), b = *reinterpret_cast<const float2*>(r1 + x);
        *reinterpret_cast<float2*>(o + x) = lerp_norm2(a, b, wy, mn, inv);
    }
};
template <>
struct HeatVec<4> {
    static __device__ __forceinline__ void lerp_minmax(const float* r0, const float* r1, int x, float wy, float& mn, float& mx) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + x), b = *reinterpret_cast<const float4*>(r1 + x);
        const float v0 = fmaf(b.x - a.x, wy, a.x), v1 = fmaf(b.y - a.y, wy, a.y);
        const float v2 = fmaf(b.z - a.z, wy, a.z), v3 = fmaf(b.w - a.w, wy, a.w);
        mn = fminf(fminf(mn, fminf(v0, v1)), fminf(v2, v3)); mx = fmaxf(fmaxf(mx, fmaxf(v0, v1)), fmaxf(v2, v3));
    }
    static __device__ __forceinline__ void lerp_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + x), b = *reinterpret_cast<const float4*>(r1 + x);
        __stcs(reinterpret_cast<float4*>(o + x),
               make_float4((fmaf(b.x - a.x, wy, a.x) - mn) * inv, (fmaf(b.y - a.y, wy, a.y) - mn) * inv,
                           (fmaf(b.z - a.z, wy, a.z) - mn) * inv, (fmaf(b.w - a.w, wy, a.w) - mn) * inv));
    }
    static __device__ __forceinline__ void norm_store(const float* r0, const float* r1, int x, float wy, float mn, float inv, float* o) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + x), b = *reinterpret_cast<const float4*>(r1 + x);
        const float2 lo = lerp_norm2(make_float2(a.x, a.y), make_float2(b.x, b.y), wy, mn, inv);
        const float2 hi = lerp_norm2(make_float2(a.z, a.w), make_float2(b.z, b.w), wy, mn, inv);
        *reinterpret_cast<float4*>(o + x) = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
};

template <int VEC>
__global__ void __launch_bounds__(kHeatThreads)
heatmap_fast_kernel(const double* __restrict__ energy, long long n_frames, int out_h, int out_w,
                    float* __restrict__ heat) {
    extern __shared__ __align__(16) float s_fast[];
    __shared__ float s_t[kFramePixels];
    __shared__ double s_red64[2][kHeatThreads / 32];
    __shared__ float s_red32[2][kHeatThreads / 32];
    float* s_rows = s_fast;                                         // [36][out_w]
    float* s_wx = s_rows + kFrameH * out_w;                         // [out_w]
    float* s_wy = s_wx + out_w;                                     // [out_h]
    int* s_x0 = reinterpret_cast<int*>(s_wy + out_h);               // [out_w] x0 | x1 << 16
    int* s_y0 = s_x0 + out_w;                                       // [out_h]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int d = tid; d < out_w; d += kHeatThreads) {
        int i0, i1; double w;
        linear_tap(d, kFrameW, out_w, &i0, &i1, &w);
        s_x0[d] = i0 | (i1 << 16); s_wx[d] = static_cast<float>(w);
    }
    for (int d = tid; d < out_h; d += kHeatThreads) {
        int i0, i1; double w;
        linear_tap(d, kFrameH, out_h, &i0, &i1, &w);
        // Bit 31 marks the first and last output row of every source-row pair.  fmaf(r1 - r0, wy, r0) is monotonic in
        // wy, and wy grows with y inside a pair, so the image's min / max are attained on the marked rows: pass 1 visits
        // only those (about 2 * 37 of out_h rows) and finds exactly the values a full pass would.
        int p0, p1, n0, n1; double wn;
        linear_tap(max(d - 1, 0), kFrameH, out_h, &p0, &p1, &wn);
        linear_tap(min(d + 1, out_h - 1), kFrameH, out_h, &n0, &n1, &wn);
        const bool edge = d == 0 || d == out_h - 1 || p0 != i0 || p1 != i1 || n0 != i0 || n1 != i1;
        s_y0[d] = i0 | (i1 << 16) | (edge ? 0x80000000 : 0); s_wy[d] = static_cast<float>(w);
    }
    constexpr int kPerThread = (kFramePixels + kHeatThreads - 1) / kHeatThreads;
    double e[kPerThread];                                          // this thread's energies of the frame being started
    auto fetch = [&](long long frame) {
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int p = tid + i * kHeatThreads;
            e[i] = (frame < n_frames && p < kFramePixels) ? __ldcs(energy + frame * kFramePixels + p) : CUDART_NAN;
        }
    };
    fetch(blockIdx.x);
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        __syncthreads();
        // frame min / max in float64, then t = (e - min) / (max - min) as float32
        double lo = CUDART_INF, hi = -CUDART_INF;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) { lo = fmin(lo, e[i]); hi = fmax(hi, e[i]); }
        lo = warp_min(lo); hi = warp_max(hi);
        if (lane == 0) { s_red64[0][warp] = lo; s_red64[1][warp] = hi; }
        __syncthreads();
        lo = s_red64[0][0]; hi = s_red64[1][0];
#pragma unroll
        for (int w = 1; w < kHeatThreads / 32; ++w) { lo = fmin(lo, s_red64[0][w]); hi = fmax(hi, s_red64[1][w]); }
        const double span = hi - lo;
        const double inv_span = 1.0 / span;        // one division per frame: t = (e - min) * (1 / span), rounded to float32
#pragma unroll
        for (int i = 0; i < (kFramePixels + kHeatThreads - 1) / kHeatThreads; ++i) {
            const int p = tid + i * kHeatThreads;
            if (p < kFramePixels) s_t[p] = span > 0.0 ? static_cast<float>((e[i] - lo) * inv_span) : 0.f;
        }
        fetch(frame + gridDim.x);                                  // next frame's energies arrive during the passes below
        __syncthreads();
        // horizontal pass, once: warps own source rows, lanes walk the output columns
        for (int r = warp; r < kFrameH; r += kHeatThreads / 32) {
            const float* t = s_t + r * kFrameW;
            float* row = s_rows + r * out_w;
            for (int x = lane; x < out_w; x += 32) {
                const int xi = s_x0[x];
                const float a = t[xi & 0xffff], b = t[xi >> 16];
                row[x] = fmaf(b - a, s_wx[x], a);
            }
        }
        __syncthreads();
        // pass 1: min / max of the up-sampled image
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        for (int y = warp; y < out_h; y += kHeatThreads / 32) {
            const int yi = s_y0[y];
            if (yi >= 0) continue;                                   // interior row of its pair: cannot hold an extreme
            const float wy = s_wy[y];
            const float* r0 = s_rows + (yi & 0xffff) * out_w;
            const float* r1 = s_rows + ((yi >> 16) & 0x7fff) * out_w;
            for (int x = lane * VEC; x < out_w; x += 32 * VEC) HeatVec<VEC>::lerp_minmax(r0, r1, x, wy, mn, mx);
        }
        mn = warp_min(mn); mx = warp_max(mx);
        if (lane == 0) { s_red32[0][warp] = mn; s_red32[1][warp] = mx; }
        __syncthreads();
        mn = s_red32[0][0]; mx = s_red32[1][0];
#pragma unroll
        for (int w = 1; w < kHeatThreads / 32; ++w) { mn = fminf(mn, s_red32[0][w]); mx = fmaxf(mx, s_red32[1][w]); }
        // a constant frame gives 0/0 = NaN, like the reference's (x - min) / (max - min)
        const float inv = (span > 0.0 && mx > mn) ? 1.f / (mx - mn) : CUDART_NAN_F;
        float* dst = heat + frame * static_cast<long long>(out_h) * out_w;
        // pass 2: normalise and stream out
        for (int y = warp; y < out_h; y += kHeatThreads / 32) {
            const int yi = s_y0[y];
            const float wy = s_wy[y];
            const float* r0 = s_rows + (yi & 0xffff) * out_w;
            const float* r1 = s_rows + ((yi >> 16) & 0x7fff) * out_w;
            float* o = dst + static_cast<long long>(y) * out_w;
            for (int x = lane * VEC; x < out_w; x += 32 * VEC) HeatVec<VEC>::lerp_store(r0, r1, x, wy, mn, inv, o);
        }
    }
}


// ---- heat_stream_kernel ---------------------------------------------------------------------------------------------
// 384 threads = 12 warps (= 6 warp pairs in the energy phase), two CTAs per SM so that one CTA's float64 phase overlaps the
// other's store phase (24 warps per SM: with 16 the passes were latency-bound, profiles/r02_ncu_heat_stream.csv).  Per frame:
//   1. energies: loaded (and prefetched a frame ahead) or computed in place (FUSED)
//   2. float64 min / max of the 1728 energies, t = (e - min) / (max - min) as float32            (as heatmap_fast_kernel)
//   3. horizontal pass once into 36 rows of out_w floats
//   4. min / max of the up-sampled image from the first and last output row of every source-row pair (monotonic lerp)
//   5. output: warp w owns row pairs w, w + 8, ...; a row is (lerp - min) * inv in packed float32 pairs (fma.rn.f32x2: the
//      same roundings as the scalar form, so the image's minimum is exactly 0), written to the warp's own double-buffered
//      staging slot and shipped with one bulk copy per row pair; bulk-copy groups are per thread, so the output pass
//      needs no CTA-wide barrier at all.
// Requirements (checked by the host): out_w even, out_h * out_w a multiple of 4, heat 16-byte aligned - then every
// row pair starts on a 16-byte boundary and is a 16-byte multiple long.  Other shapes run heatmap_fast_kernel.
constexpr int kStreamThreads = 384;                  // 12 warps: 1728 pixels = 9 rounds of 6 warp pairs, 36 source rows = 3 per warp
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kHeatMaxWarps = 16;                    // size of the reduction scratch of heat_phase

struct EnergyPhaseShared {             // FUSED: lives in the staging area, which is idle while a frame's energies are computed
    double map[kFramePixels];
    double part[16][8];
    double leaf[16];
    double mean;
    float red[3 * kStreamWarps];
    unsigned int rare_bits[kFramePixels / 32];
    EnergyTables tab;                  // reloaded every frame (9 KB from L2): the heat-map rows overwrite it
};

struct HeatStreamLayout {
    int wp;                 // row stride of the 36 blended rows (out_w rounded up to 4)
    unsigned int off_rows, off_stage, off_taps, total;
};
__host__ __device__ inline HeatStreamLayout heat_stream_layout(int out_h, int out_w, bool fused, int warps = kStreamWarps,
                                                              bool t_elsewhere = false) {
    HeatStreamLayout l;
    l.wp = (out_w + 3) & ~3;
    l.off_rows = t_elsewhere ? 0 : kFramePixels * 4;                        // after t[1728] unless the caller places t itself
    l.off_stage = l.off_rows + kFrameH * l.wp * 4;
    unsigned int stage = warps * 2 * 2 * out_w * 4;                         // per warp: 2 buffers of 2 rows
    if (fused && stage < sizeof(EnergyPhaseShared)) stage = sizeof(EnergyPhaseShared);
    l.off_taps = l.off_stage + ((stage + 15u) & ~15u);
    l.total = l.off_taps + (out_w + out_h) * 8;
    return l;
}

struct HeatStreamArgs {
    Stage2Args s2;              // FUSED: slot 0 = images and the optional energy / mask / mean outputs
    const double* energy_in;    // !FUSED
    long long n_frames;
    int out_h, out_w;
    float* heat;
    unsigned int jitter_seed;   // energy_heat_ws_kernel<..., JITTER = true> only ("debug_jitter")
};

// (d = b - a) once per column pair, then per output row v = d * wy + a, (v - mn) * inv: the roundings of lerp_norm2.
__device__ __forceinline__ unsigned long long diff2(unsigned long long pa, unsigned long long pb) {
    unsigned long long d;
    const unsigned long long neg1 = pack2(make_float2(-1.f, -1.f));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pa), "l"(neg1), "l"(pb));
    return d;
}
__device__ __forceinline__ unsigned long long norm2(unsigned long long pa, unsigned long long d, unsigned long long w2,
                                                    unsigned long long nmn, unsigned long long i2) {
    unsigned long long v, t, o;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(d), "l"(w2), "l"(pa));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(v), "l"(nmn));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(t), "l"(i2));
    return o;
}

// ---- the heat-map phase of one frame, shared by heat_stream_kernel and the persistent MFCC + energy + heat-map kernel ----
struct HeatSmem {              // carved out of dynamic shared memory (heat_stream_layout)
    float* t;                  // [1728] frame-normalised energies
    float* rows;               // [36][wp] horizontally blended rows
    float* stage;              // [warps][2 buffers][2 rows][out_w] staging slots of the bulk copies
    float* wx; float* wy;      // [out_w], [out_h] tap weights
    int* x0; int* y0;          // [out_w], [out_h] tap indices (lo | hi << 16; y0 bit 31 = edge row of its source-row pair)
    int wp;
};
__device__ __forceinline__ HeatSmem heat_smem_carve(unsigned char* base, const HeatStreamLayout& lay, int out_h, int out_w) {
    HeatSmem s;
    s.t = reinterpret_cast<float*>(base);
    s.rows = reinterpret_cast<float*>(base + lay.off_rows);
    s.stage = reinterpret_cast<float*>(base + lay.off_stage);
    s.wx = reinterpret_cast<float*>(base + lay.off_taps);
    s.wy = s.wx + out_w;
    s.x0 = reinterpret_cast<int*>(s.wy + out_h);
    s.y0 = s.x0 + out_w;
    s.wp = lay.wp;
    return s;
}
// Called once per CTA by the THREADS threads that run the heat phase (index tid); synchronise before first use.
template <int THREADS>
__device__ __forceinline__ void heat_taps_init(const HeatSmem& s, int out_h, int out_w, int tid) {
    for (int d = tid; d < out_w; d += THREADS) {
        int i0, i1; double w;
        linear_tap(d, kFrameW, out_w, &i0, &i1, &w);
        s.x0[d] = i0 | (i1 << 16); s.wx[d] = static_cast<float>(w);
    }
    for (int d = tid; d < out_h; d += THREADS) {
        int i0, i1; double w;
        linear_tap(d, kFrameH, out_h, &i0, &i1, &w);
        int p0, p1, n0, n1; double wn;                  // bit 31: first / last output row of its source-row pair (see heatmap_fast_kernel)
        linear_tap(max(d - 1, 0), kFrameH, out_h, &p0, &p1, &wn);
        linear_tap(min(d + 1, out_h - 1), kFrameH, out_h, &n0, &n1, &wn);
        const bool edge = d == 0 || d == out_h - 1 || p0 != i0 || p1 != i1 || n0 != i0 || n1 != i1;
        s.y0[d] = i0 | (i1 << 16) | (edge ? 0x80000000 : 0); s.wy[d] = static_cast<float>(w);
    }
}

// pixels of the frame held by each of THREADS threads (thread tid owns pixels tid + THREADS * i)
template <int THREADS>
struct HeatPerThread { static constexpr int value = (kFramePixels + THREADS - 1) / THREADS; };

// Steps 2-5 of the list above for one frame whose 1728 energies the THREADS threads hold in e[] (thread tid owns
// pixels tid + THREADS i).  W, H: the output size as template constants (0 = run-time size): the reference's two sizes,
// 298 x 224 (showimages.py:147) and 224 x 224 (BASELINE configs[2]), are instantiated with constants, every column loop
// then unrolls completely with immediate offsets - the run-time-size build spent half of its instructions on loop
// bounds and address arithmetic (profiles/r02_ncu_stage2_kernels.csv).  `sync` is the barrier of the participating
// threads, `after_t` runs once the energies have been consumed (the stand-alone kernel prefetches the next frame's there).
// chunk_it is the calling warp's running chunk counter (selects its staging buffer; survives across frames).
struct HeatNoHook { __device__ __forceinline__ void operator()() const {} };
template <int THREADS, int VEC, int W, int H, typename Sync, typename AfterT, typename AfterRows = HeatNoHook>
__device__ __forceinline__ void heat_phase(const double (&e)[HeatPerThread<THREADS>::value], const HeatSmem& s,
                                           double (*red64)[kHeatMaxWarps], float (*red32)[kHeatMaxWarps], int out_h_rt,
                                           int out_w_rt, float* dst, int tid, unsigned int& chunk_it, Sync sync, AfterT after_t,
                                           AfterRows after_rows = AfterRows{}) {
    constexpr int kWarps = THREADS / 32, kPerThread = HeatPerThread<THREADS>::value;
    static_assert(kWarps <= kHeatMaxWarps, "reduction scratch too small");
    const int out_h = H ? H : out_h_rt, out_w = W ? W : out_w_rt;
    const int warp = tid >> 5, lane = tid & 31;
    const int wp = s.wp;
    // column loops: x = (lane + 32 k) * V for k = 0 .. ; fully unrolled when the width is a template constant
    auto columns = [&](auto vec_tag, auto body) {
        constexpr int V = decltype(vec_tag)::value;
        if constexpr (W != 0) {
            constexpr int kIters = (W / V + 31) / 32;
#pragma unroll
            for (int k = 0; k < kIters; ++k) {
                const int x = (lane + 32 * k) * V;
                if (k + 1 < kIters || x < W) body(x);
            }
        } else {
            for (int x = lane * V; x < out_w; x += 32 * V) body(x);
        }
    };
    // 2. frame min / max in float64, t = (e - min) / (max - min) as float32
    double lo = CUDART_INF, hi = -CUDART_INF;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) { lo = fmin(lo, e[i]); hi = fmax(hi, e[i]); }
    lo = warp_min(lo); hi = warp_max(hi);
    sync();                               // the previous frame's readers of red64 / t / rows are done
    if (lane == 0) { red64[0][warp] = lo; red64[1][warp] = hi; }
    sync();
    lo = red64[0][0]; hi = red64[1][0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) { lo = fmin(lo, red64[0][w]); hi = fmax(hi, red64[1][w]); }
    const double span = hi - lo;
    const double inv_span = 1.0 / span;            // one division per frame, not one per energy
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int p = tid + i * THREADS;
        if (p < kFramePixels) s.t[p] = span > 0.0 ? static_cast<float>((e[i] - lo) * inv_span) : 0.f;
    }
    after_t();
    sync();
    // 3. horizontal pass: a column's taps are loaded once and applied to this warp's source rows (warp, warp + 8, ...)
    columns(std::integral_constant<int, 1>{}, [&](int x) {
        const int xi = s.x0[x];
        const float wx = s.wx[x];
        const int c0 = xi & 0xffff, c1 = xi >> 16;
#pragma unroll
        for (int r = warp; r < kFrameH; r += kWarps) {
            const float u = s.t[r * kFrameW + c0], v = s.t[r * kFrameW + c1];
            s.rows[r * wp + x] = fmaf(v - u, wx, u);
        }
    });
    sync();
    after_rows();                         // every participating thread is done with s.t
    // 4. min / max of the up-sampled image
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int y = warp; y < out_h; y += kWarps) {
        const int yi = s.y0[y];
        if (yi >= 0) continue;                                // interior row of its pair: cannot hold an extreme
        const float wy = s.wy[y];
        const float* r0 = s.rows + (yi & 0xffff) * wp;
        const float* r1 = s.rows + ((yi >> 16) & 0x7fff) * wp;
        columns(std::integral_constant<int, VEC>{}, [&](int x) { HeatVec<VEC>::lerp_minmax(r0, r1, x, wy, mn, mx); });
    }
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0) { red32[0][warp] = mn; red32[1][warp] = mx; }
    sync();
    mn = red32[0][0]; mx = red32[1][0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) { mn = fminf(mn, red32[0][w]); mx = fmaxf(mx, red32[1][w]); }
    // a constant frame gives 0/0 = NaN, like the reference's (x - min) / (max - min)
    const float inv = (span > 0.0 && mx > mn) ? 1.f / (mx - mn) : CUDART_NAN_F;
    const unsigned long long nmn2 = pack2(make_float2(-mn, -mn)), inv2 = pack2(make_float2(inv, inv));
    // 5. output: row pairs through this warp's staging slot, one bulk copy each
    float* my_stage = s.stage + warp * 4 * out_w;                 // [2 buffers][2 rows][out_w]
    const uint32_t my_stage_addr = smem_u32(my_stage);
    const int n_pairs = (out_h + 1) / 2;
    for (int q = warp; q < n_pairs; q += kWarps, ++chunk_it) {
        const unsigned int buf = chunk_it & 1u;
        if (chunk_it >= 2u) {                                 // the copy issued two chunks ago has finished reading this buffer
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
        }
        float* stage = my_stage + buf * 2 * out_w;
        const int y0 = 2 * q;
        const bool two = (H != 0 && H % 2 == 0) || y0 + 1 < out_h;
        const int ya = s.y0[y0], yb = s.y0[two ? y0 + 1 : y0];
        const float wa = s.wy[y0], wb = s.wy[two ? y0 + 1 : y0];
        const unsigned long long wa2 = pack2(make_float2(wa, wa)), wb2 = pack2(make_float2(wb, wb));
        const float* a0 = s.rows + (ya & 0xffff) * wp;
        const float* a1 = s.rows + ((ya >> 16) & 0x7fff) * wp;
        if (two && ((ya ^ yb) & 0x7fffffff) == 0) {
            // both output rows interpolate between the same two source rows (five times out of six at 36 -> 224):
            // one pair of loads and one difference serve both
            columns(std::integral_constant<int, VEC>{}, [&](int x) {
#pragma unroll
                for (int h = 0; h < VEC; h += 2) {
                    const unsigned long long pa = *reinterpret_cast<const unsigned long long*>(a0 + x + h);
                    const unsigned long long pb = *reinterpret_cast<const unsigned long long*>(a1 + x + h);
                    const unsigned long long d = diff2(pa, pb);
                    *reinterpret_cast<unsigned long long*>(stage + x + h) = norm2(pa, d, wa2, nmn2, inv2);
                    *reinterpret_cast<unsigned long long*>(stage + out_w + x + h) = norm2(pa, d, wb2, nmn2, inv2);
                }
            });
        } else {
            const float* b0 = s.rows + (yb & 0xffff) * wp;
            const float* b1 = s.rows + ((yb >> 16) & 0x7fff) * wp;
            columns(std::integral_constant<int, VEC>{}, [&](int x) {
#pragma unroll
                for (int h = 0; h < VEC; h += 2) {
                    const unsigned long long pa = *reinterpret_cast<const unsigned long long*>(a0 + x + h);
                    const unsigned long long pb = *reinterpret_cast<const unsigned long long*>(a1 + x + h);
                    *reinterpret_cast<unsigned long long*>(stage + x + h) = norm2(pa, diff2(pa, pb), wa2, nmn2, inv2);
                    if (two) {
                        const unsigned long long qa = *reinterpret_cast<const unsigned long long*>(b0 + x + h);
                        const unsigned long long qb = *reinterpret_cast<const unsigned long long*>(b1 + x + h);
                        *reinterpret_cast<unsigned long long*>(stage + out_w + x + h) = norm2(qa, diff2(qa, qb), wb2, nmn2, inv2);
                    }
                }
            });
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store_s2g(dst + static_cast<long long>(y0) * out_w, my_stage_addr + buf * 2 * out_w * 4,
                           static_cast<uint32_t>((two ? 2 : 1) * out_w * 4));
            bulk_commit();
        }
    }
}

template <bool FUSED, int VEC, int W, int H>
__global__ void __launch_bounds__(kStreamThreads, 2)
heat_stream_kernel(const __grid_constant__ HeatStreamArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ double s_red64[2][kHeatMaxWarps];
    __shared__ float s_red32[2][kHeatMaxWarps];
    const int out_h = H ? H : a.out_h, out_w = W ? W : a.out_w;
    const HeatStreamLayout lay = heat_stream_layout(out_h, out_w, FUSED);
    const HeatSmem hs = heat_smem_carve(s_raw, lay, out_h, out_w);
    EnergyPhaseShared& eph = *reinterpret_cast<EnergyPhaseShared*>(s_raw + lay.off_stage);
    const int tid = threadIdx.x, lane = tid & 31;
    heat_taps_init<kStreamThreads>(hs, out_h, out_w, tid);
    constexpr int kPerThread = HeatPerThread<kStreamThreads>::value;
    double e[kPerThread];
    auto fetch = [&](long long frame) {
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int p = tid + i * kStreamThreads;
            e[i] = (frame < a.n_frames && p < kFramePixels) ? __ldcs(a.energy_in + frame * kFramePixels + p) : CUDART_NAN;
        }
    };
    if (!FUSED) fetch(blockIdx.x);
    unsigned int chunk_it = 0;                                    // this warp's chunk counter: selects the staging buffer
    const long long frame_values = static_cast<long long>(out_h) * out_w;
    __syncthreads();

    for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        if (FUSED) {
            // the staging area doubles as the energy phase's scratch: the previous frame's copies must have left it
            if (lane == 0) bulk_wait_read<0>();
            __syncthreads();
            const Stage2Args& s = a.s2;
            const float* img = s.img[0] + frame * kFrameValues;
            float lo = 0.f, hi = 1.f;
            if (s.normalize_first) group_minmax(img, kFrameValues / 4, tid, kStreamThreads, eph.red, [] { __syncthreads(); }, lo, hi);
            const FrameNormFast norm(lo, __fsub_rn(hi, lo));
            double* energy = s.energy[0] ? s.energy[0] + frame * kFramePixels : nullptr;
            if (tid < kFramePixels / 32) eph.rare_bits[tid] = 0u;       // the staging area held heat-map rows a moment ago
            load_energy_tables(eph.tab, tid, kStreamThreads);
            __syncthreads();
            frame_energy_pixels<kStreamThreads, false>(img, 0, kFramePixels, s.normalize_first != 0, norm, nullptr, energy, eph.map,
                                                       eph.rare_bits, eph.tab, nullptr, tid);
            __syncthreads();
            if (frame_energy_fixup(img, 0, kFramePixels, s.normalize_first != 0, norm, nullptr, energy, eph.map, eph.rare_bits,
                                   tid, kStreamThreads))
                __syncthreads();
            if (s.mask[0] != nullptr || s.mean[0] != nullptr) {
                const double mean = frame_mean(eph.map, eph.part, eph.leaf, &eph.mean, tid, kStreamThreads, [] { __syncthreads(); });
                if (tid == 0 && s.mean[0] != nullptr) s.mean[0][frame] = mean;
                if (s.mask[0] != nullptr)
                    for (int p = tid; p < kFramePixels; p += kStreamThreads)
                        s.mask[0][frame * kFramePixels + p] = eph.map[p] > mean ? 1 : 0;
            }
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                const int p = tid + i * kStreamThreads;
                e[i] = p < kFramePixels ? eph.map[p] : CUDART_NAN;
            }
        }
        // (FUSED: heat_phase's first barrier also orders the reads of eph.map above before the staging area is written)
        heat_phase<kStreamThreads, VEC, W, H>(e, hs, s_red64, s_red32, out_h, out_w, a.heat + frame * frame_values, tid, chunk_it,
                              [] { __syncthreads(); },
                              [&] { if (!FUSED) fetch(frame + gridDim.x); });   // next frame's energies arrive during the passes
    }
    if (lane == 0) bulk_wait_all<0>();        // shared memory must outlive the copies that read it
}

// ---- energy + heat map, warp-specialised (aig_energy_heatmap) --------------------------------------------------------
// heat_stream_kernel<true> runs find_logen and the heat-map phase one after the other in the same warps: ncu (round 2)
// shows the two phases simply add up - 55 % of the samples in the float64 pixel loop, 25 % in the up-sampling passes, the
// FP64 pipe busy 29 % of the time against 50-58 % in the stand-alone energy kernel - because two CTAs per SM drift into
// the same phase and a lone CTA's 12 warps cannot fill the pipe.  Here the two stages are different warps of a CTA and
// overlap by construction:
//   ET threads          find_logen of frame i + 1 (+ optional min-max, mean, mask, energy output) into map[slot]
//   HW warps            heat_phase() of frame i: normalise, up-sample, stage rows, bulk copies
// The map is handed over through a full / empty mbarrier pair per slot, as the persistent MFCC + energy kernel hands its
// frames to the energy warps.  The shape in use (WsTwin): two CTAs of 352 + 160 threads per SM - two independent
// pipelines whose min-max / pixel-loop / mean phases interleave, 4.9 pixel rounds per frame - with ONE map slot whose
// first half doubles as the heat-map phase's float32 t[] (the energies are in the heat-map threads' registers by then;
// the slot goes back to the float64 warps after the horizontal pass), which is what lets two CTAs fit in 227 KB.
// How the 512 threads (64 registers each: the whole register file with two CTAs) are split was measured at 224 x 298
// (M frames/s, 8192 frames; heat_stream_kernel<true>: 7.03): float64 + heat-map threads 256 + 256: 7.70, 288 + 224: 7.95,
// 320 + 192: 7.83, 352 + 160: 8.48, 352 + 128: 8.24, 384 + 128: 8.37 (8.14 with two map slots), 416 + 96: 7.56,
// 448 + 64: 5.88; one CTA of 512 + 512 per SM with two map slots: 6.85.  The float64 side is the critical path and is
// latency-bound at this occupancy (ncu, 256 + 256: FP64 pipe 31 %, issue slots 52 %, 5.5 warps per issue waiting on the
// long scoreboard), so it gets the warps; five heat-map warps per CTA still keep up (20 M frames/s with 24 per SM alone).
template <int ET_, int HW_, int CTAS_, int SLOTS_>
struct WsConfig {
    static constexpr int ET = ET_, HW = HW_, CTAS = CTAS_, SLOTS = SLOTS_;
    static constexpr int HT = HW * 32, THREADS = ET + HT;
    static_assert(HW <= kHeatMaxWarps, "reduction scratch of heat_phase");
    static_assert(SLOTS == 1 || SLOTS == 2, "map slots");
};
using WsTwin = WsConfig<352, 5, 2, 1>;

template <typename C>
struct WsShared {                      // behind the heat-map phase's rows / staging slots / taps in dynamic shared memory
    EnergyTables tab;
    double map[C::SLOTS][kFramePixels];
    double part[16][8];
    double leaf[16];
    double mean;
    double red64[2][kHeatMaxWarps];
    float red32[2][kHeatMaxWarps];
    float red[3 * (C::ET / 32)];
    unsigned int rare_bits[kFramePixels / 32];
    unsigned long long bar[4];         // full[2], empty[2]
};
template <typename C>
__host__ __device__ inline size_t energy_heat_ws_smem(int out_h, int out_w) {
    return ((heat_stream_layout(out_h, out_w, false, C::HW, C::SLOTS == 1).total + 15u) & ~15u) + sizeof(WsShared<C>) + 16;
}

// JITTER = true is the "debug_jitter" build: both roles spin for pseudo-random times around their barrier operations
// (aig_common.cuh: jitter_spin), which walks the hand-over through every relative order of the two sides; its results
// must still equal the sequential kernel's (tests: test_race_stress_jittered_energy_heat_ws_kernel).
template <typename C, int VEC, int W, int H, bool JITTER = false>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
energy_heat_ws_kernel(const __grid_constant__ HeatStreamArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr bool kOneSlot = C::SLOTS == 1;
    const int out_h = H ? H : a.out_h, out_w = W ? W : a.out_w;
    const HeatStreamLayout lay = heat_stream_layout(out_h, out_w, false, C::HW, kOneSlot);
    HeatSmem hs = heat_smem_carve(s_raw, lay, out_h, out_w);
    WsShared<C>& ws = *reinterpret_cast<WsShared<C>*>(s_raw + ((lay.total + 15u) & ~15u));
    if (kOneSlot) hs.t = reinterpret_cast<float*>(ws.map[0]);
    const uint32_t full = smem_u32(&ws.bar[0]), empty = smem_u32(&ws.bar[2]);        // + 8 * slot
    const int tid = threadIdx.x;
    unsigned int jitter_counter = 0;
    if (tid == 0) {
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full + 8 * s, C::ET);                  // every energy thread arrives (release of its map values)
            mbar_init(empty + 8 * s, C::HT);                 // every heat-map thread arrives when it no longer needs the slot
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid < C::ET) {
        // ---------------------------------------- float64 warps ----------------------------------------
        const int et = tid;
        auto group_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(C::ET) : "memory"); };
        load_energy_tables(ws.tab, et, C::ET);
        if (et < kFramePixels / 32) ws.rare_bits[et] = 0u;
        group_sync();
        const Stage2Args& s = a.s2;
        unsigned int it = 0;
        for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x, ++it) {
            const unsigned int slot = kOneSlot ? 0u : it & 1u, use = kOneSlot ? it : it >> 1;
            const float* img = s.img[0] + frame * kFrameValues;
            float lo = 0.f, hi = 1.f;
            if (s.normalize_first) group_minmax(img, kFrameValues / 4, et, C::ET, ws.red, group_sync, lo, hi);
            const FrameNormFast norm(lo, __fsub_rn(hi, lo));
            double* energy = s.energy[0] ? s.energy[0] + frame * kFramePixels : nullptr;
            double* map = ws.map[slot];
            if (JITTER) jitter_spin(a.jitter_seed, 11u, jitter_counter);
            mbar_wait(empty + 8 * slot, (use & 1u) ^ 1u);                      // the heat-map warps are done with this slot's last frame
            if (JITTER) jitter_spin(a.jitter_seed, 12u, jitter_counter);
            // (the cp.async input staging of stage2_kernel was measured here too: 8.18 M frames/s against 8.48 M without)
            frame_energy_pixels<C::ET, false>(img, 0, kFramePixels, s.normalize_first != 0, norm, nullptr, energy, map,
                                              ws.rare_bits, ws.tab, nullptr, et);
            group_sync();
            if (frame_energy_fixup(img, 0, kFramePixels, s.normalize_first != 0, norm, nullptr, energy, map, ws.rare_bits, et,
                                   C::ET)) {
                group_sync();
                if (et < kFramePixels / 32) ws.rare_bits[et] = 0u;
                group_sync();
            }
            if (s.mask[0] != nullptr || s.mean[0] != nullptr) {
                const double mean = frame_mean(map, ws.part, ws.leaf, &ws.mean, et, C::ET, group_sync);
                if (et == 0 && s.mean[0] != nullptr) s.mean[0][frame] = mean;
                if (s.mask[0] != nullptr)
                    for (int p = et; p < kFramePixels; p += C::ET)
                        s.mask[0][frame * kFramePixels + p] = map[p] > mean ? 1 : 0;
            }
            if (JITTER) jitter_spin(a.jitter_seed, 13u, jitter_counter);
            mbar_arrive(full + 8 * slot);                                      // release: this thread's map values are visible
        }
        return;
    }

    // -------------------------------------------- heat-map warps --------------------------------------------
    const int ht = tid - C::ET;
    auto heat_sync = [] { asm volatile("bar.sync 2, %0;" ::"n"(C::HT) : "memory"); };
    heat_taps_init<C::HT>(hs, out_h, out_w, ht);
    heat_sync();
    constexpr int kPerThread = HeatPerThread<C::HT>::value;
    unsigned int chunk_it = 0, it = 0;
    const long long frame_values = static_cast<long long>(out_h) * out_w;
    for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x, ++it) {
        const unsigned int slot = kOneSlot ? 0u : it & 1u, use = kOneSlot ? it : it >> 1;
        if (JITTER) jitter_spin(a.jitter_seed, 14u, jitter_counter);
        mbar_wait_relaxed(full + 8 * slot, use & 1u, 256);                     // usually microseconds: the float64 warps are the slower side
        double e[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int p = ht + i * C::HT;
            e[i] = p < kFramePixels ? ws.map[slot][p] : CUDART_NAN;
        }
        // two slots: the energies are in registers, the slot is free.  One slot: t[] lives in it until the horizontal pass
        // is over (heat_phase's first barrier orders every thread's reads of the map above before anyone writes t).
        if (JITTER) jitter_spin(a.jitter_seed, 15u, jitter_counter);
        if (!kOneSlot) mbar_arrive(empty + 8 * slot);
        heat_phase<C::HT, VEC, W, H>(e, hs, ws.red64, ws.red32, out_h, out_w, a.heat + frame * frame_values, ht, chunk_it,
                                     heat_sync, [] {}, [&] {
                                         if (JITTER) jitter_spin(a.jitter_seed, 16u, jitter_counter);
                                         if (kOneSlot) mbar_arrive(empty);
                                     });
    }
    if ((ht & 31) == 0) bulk_wait_all<0>();       // shared memory must outlive the copies that read it
}

// mask [n, 36, 48] u8 -> mask_up [n, out_h, out_w] u8, value 1 iff bilinear(mask != 0) > 1/2 exactly.
// The bilinear value is a ratio of integers: with tap numerators xn / (2 out_w) and yn / (2 out_h),
//   val = [m00 (xd - xn) + m01 xn] (yd - yn) + [m10 (xd - xn) + m11 xn] yn   over   xd yd   (xd = 2 out_w, yd = 2 out_h)
// so "> 0.5" is 2 val > xd yd with no rounding.  Separable: the 36 source rows are blended horizontally once per
// frame (values 0..xd <= 4096, uint16), every output pixel then needs two shared loads and two integer multiply-adds
// (val <= 2^24).
struct MaskTaps {
    int* x0;            // [out_w] x0 | x1 << 16
    int* xn;            // [out_w]
    int* y0;            // [out_h] y0 | y1 << 16
    int* yn;            // [out_h]
    uint16_t* rows;     // [36][out_w] horizontally blended source rows of the current frame
    __device__ __forceinline__ void carve(int* base, int out_h, int out_w) {
        x0 = base; xn = x0 + out_w; y0 = xn + out_w; yn = y0 + out_h;
        rows = reinterpret_cast<uint16_t*>(yn + out_h);
    }
    static __host__ __device__ size_t bytes(int out_h, int out_w) {
        return static_cast<size_t>(out_w + out_h) * 2 * sizeof(int) + static_cast<size_t>(kFrameH) * out_w * sizeof(uint16_t);
    }
    __device__ __forceinline__ void build_taps(int out_h, int out_w, int tid, int n_threads) const {
        for (int d = tid; d < out_w; d += n_threads) {
            int i0, i1, r; linear_tap_exact(d, kFrameW, out_w, &i0, &i1, &r);
            x0[d] = i0 | (i1 << 16); xn[d] = r;
        }
        for (int d = tid; d < out_h; d += n_threads) {
            int i0, i1, r; linear_tap_exact(d, kFrameH, out_h, &i0, &i1, &r);
            y0[d] = i0 | (i1 << 16); yn[d] = r;
        }
    }
    // warp per source row, lanes over output columns
    __device__ __forceinline__ void blend_rows(const uint8_t* s_mask, int out_w, int warp, int lane, int n_warps) const {
        const int xd = 2 * out_w;
        for (int ys = warp; ys < kFrameH; ys += n_warps) {
            const uint8_t* m = s_mask + ys * kFrameW;
            uint16_t* dst = rows + ys * out_w;
            for (int x = lane; x < out_w; x += 32) {
                const int xi = x0[x], n = xn[x];
                dst[x] = static_cast<uint16_t>(m[xi & 0xffff] * (xd - n) + m[xi >> 16] * n);
            }
        }
    }
};

__global__ void __launch_bounds__(kHeatThreads)
resize_mask_kernel(const uint8_t* __restrict__ mask, long long n_frames, int out_h, int out_w,
                   uint8_t* __restrict__ mask_up) {
    extern __shared__ int s_taps[];
    __shared__ uint8_t s_mask[kFramePixels];
    MaskTaps t;
    t.carve(s_taps, out_h, out_w);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    t.build_taps(out_h, out_w, tid, kHeatThreads);
    const int yd = 2 * out_h, half = 2 * out_w * out_h;       // xd * yd / 2
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        __syncthreads();
        for (int p = tid; p < kFramePixels; p += kHeatThreads) s_mask[p] = mask[frame * kFramePixels + p] != 0;
        __syncthreads();
        t.blend_rows(s_mask, out_w, warp, lane, kHeatThreads / 32);
        __syncthreads();
        uint8_t* dst = mask_up + frame * static_cast<long long>(out_h) * out_w;
        for (int y = warp; y < out_h; y += kHeatThreads / 32) {
            const int yi = t.y0[y], n = t.yn[y];
            const uint16_t* r0 = t.rows + (yi & 0xffff) * out_w;
            const uint16_t* r1 = t.rows + (yi >> 16) * out_w;
            uint8_t* o = dst + static_cast<long long>(y) * out_w;
            for (int x = lane; x < out_w; x += 32) o[x] = static_cast<uint8_t>(r0[x] * (yd - n) + r1[x] * n > half);
        }
    }
}

}  // namespace aig
