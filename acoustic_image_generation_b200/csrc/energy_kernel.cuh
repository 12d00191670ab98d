// Stage 2 kernels: 12-channel acoustic image -> energy map, mean mask, up-sampled heat map.
//
// energy_kernel      F4 + F5 + F7  per-frame min-max (outdoor_data_mfcc.py:672-679), find_logen
//                    (iouenergythreshold.py:294-323) and the mean-threshold mask (:217-219).
// heatmap_kernel     F6            cv2.resize bilinear + implicit Normalize (showimages.py:147-148).
// resize_mask_kernel F9 (part)     cv2.resize(mask) > 0.5 in exact integers (showimages_bb.py:303-304).
//
// One CTA per frame: the per-frame reductions (min, max, mean) are CTA-local.  The energy is
// computed in float64 like the reference (float32-stored in-place scaling, then a float64
// 12x24 projection, exp, band sum in NumPy's pairwise order, reciprocal) so that the
// `map > mean(map)` decision is reproduced bit for bit up to the last-ulp differences of exp()
// and of the BLAS summation order; the float64 mean follows NumPy's pairwise summation tree
// for 1728 elements exactly.  FP64 work per frame is ~2 MFLOP, far below the HBM time of the
// stage-1 stream it follows, and runs on the otherwise idle FP64 pipe.
#pragma once

#include "aig_common.cuh"
#include "mel_tables_ref.inc"

namespace aig {

// find_logen's constants (iouenergythreshold.py:304-308), identical to the MFCC constants.
__constant__ double c_dct[kFilterNum * kMfccNum] = AIG_REF_DCT;      // [24][12]
__constant__ double c_lifter[kMfccNum] = AIG_REF_LIFTER;
__constant__ double c_mfnorm = AIG_REF_MFNORM;

// 128 threads = one warp per SM sub-partition, so four CTAs per SM may use 128 registers each (the per-pixel float64
// chain needs ~120; with 192-thread CTAs the uneven warp split capped it at 96 and it spilled).  13.5 pixels per thread:
// in the 14th round the upper two warps idle.
constexpr int kEnergyThreads = 128;
constexpr int kPixelsPerThread = (kFramePixels + kEnergyThreads - 1) / kEnergyThreads;
static_assert(kFramePixels % 64 == 0, "whole warps drop out of the last round");

// NumPy pairwise-sum leaves for n = 1728: 1728 -> 864 -> 432 -> 216 -> (104, 112); every leaf is
// summed with 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).
__device__ __forceinline__ int leaf_start(int leaf) { return 216 * (leaf >> 1) + ((leaf & 1) ? 104 : 0); }
__device__ __forceinline__ int leaf_len(int leaf) { return (leaf & 1) ? 112 : 104; }

__constant__ double c_inv_lifter[kMfccNum] = AIG_REF_INV_LIFTER;     // RN(1 / lifter[m])
__constant__ double c_exp2_table[64] = AIG_EXP2_TABLE;               // RN(2^(j/64)); kernels copy it to shared memory

// v / L for a constant L with r = RN(1 / L): q0 = RN(v * r), e = v - q0 * L (exact in one FMA), q = RN(q0 + e * r).
// By Markstein's theorem q is the correctly rounded quotient for finite v; tests/test_gpu_parity.py checks it against
// __ddiv_rn for every finite float32 v and all twelve lifter constants.  Non-finite inputs never reach it: pixel_energy
// sends such pixels down the plain path.
__device__ __forceinline__ double div_by_lifter(double v, int m) {
    const double r = c_inv_lifter[m];
    const double q0 = __dmul_rn(v, r);
    const double e = __fma_rn(-q0, c_lifter[m], v);
    return __fma_rn(e, r, q0);
}

// The per-frame min-max normalisation (x - lo) / range in float32 (:672-679): one IEEE division per value, like the
// reference.  (A shared-reciprocal form with one or two residual corrections was tried and rejected: against __fdiv_rn on
// 2^34 pseudo-random triples it left 7e-5 of the quotients one ulp off - a 24-bit reciprocal is not enough - so unlike
// the float64 division by the lifter constants it cannot replace the division exactly.)
struct FrameNorm {
    float lo, range;
    __device__ __forceinline__ FrameNorm(float lo_, float range_) : lo(lo_), range(range_) {}
    __device__ __forceinline__ float apply(float x) const { return __fdiv_rn(__fsub_rn(x, lo), range); }
};

// exp(x) by table: k = rint(x * 64 / ln2), r = x - k * ln2 / 64 (two-part constant), exp(x) =
// 2^(k >> 6) * T[k & 63] * (1 + p(r)) with a degree-6 polynomial on |r| <= ln2 / 128 (truncation 3e-20).  Worst-case
// error just under 1 ulp (table entry + final rounding), the same class as CUDA's and NumPy's exp; 11 FP64 operations
// instead of ~17 plus the special-case branches.  Valid for |x| <= 700 (|k| < 2^16); `out_of_range` collects the
// violation so that the caller can redo the pixel with exp() - NaN propagates by itself.
// The constants live in __constant__ memory so that each FMA takes its coefficient as a constant-bank operand (as
// immediates every 64-bit coefficient costs two extra UMOVs per use).
__constant__ double c_exp_k[9] = {AIG_EXP_64_OVER_LN2, 6755399441055744.0 /* 1.5 * 2^52: rint by addition */,
                                  -(AIG_EXP_LN2_64_HEAD), -(AIG_EXP_LN2_64_TAIL),
                                  1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

__device__ __forceinline__ double exp_table64(double x, const double* __restrict__ table, unsigned int& out_of_range) {
    const double t = __fma_rn(x, c_exp_k[0], c_exp_k[1]);
    const int k = __double2loint(t);
    out_of_range |= static_cast<unsigned int>(k + 65536) >> 17;       // non-zero iff |k| >= 2^16 (|x| > ~709)
    const double kd = __dadd_rn(t, -c_exp_k[1]);
    double r = __fma_rn(kd, c_exp_k[2], x);
    r = __fma_rn(kd, c_exp_k[3], r);
    double p = __fma_rn(r, c_exp_k[4], c_exp_k[5]);
    p = __fma_rn(p, r, c_exp_k[6]);
    p = __fma_rn(p, r, c_exp_k[7]);
    p = __fma_rn(p, r, c_exp_k[8]);
    p = __fma_rn(__dmul_rn(r, r), p, r);                              // e^r - 1
    const double tj = table[k & 63];
    const double y = __fma_rn(tj, p, tj);
    return __hiloint2double(__double2hiint(y) + ((k >> 6) << 20), __double2loint(y));   // * 2^(k >> 6)
}

// The plain form of one pixel (IEEE divisions, exp(), straight 12-term dot products), kept out of line: it serves the
// pixels the fast path cannot (non-finite inputs, |mel| > 700) and is what the fast path is checked against.
// `src` / `scaled_dst` are global-memory pointers (scaled_dst may be null).
__device__ __noinline__ double pixel_energy_plain(const float* __restrict__ src, float* __restrict__ scaled_dst, bool normalize,
                                                  float lo, float range) {
    double z[kMfccNum];
    for (int m = 0; m < kMfccNum; ++m) {
        float v = src[m];
        if (normalize) v = __fdiv_rn(__fsub_rn(v, lo), range);
        v = __double2float_rn(__ddiv_rn(static_cast<double>(v), c_lifter[m]));
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), c_mfnorm));
        if (scaled_dst != nullptr) scaled_dst[m] = v;
        z[m] = static_cast<double>(v);
    }
    double r[8];
    for (int j = 0; j < kFilterNum; ++j) {
        double mel = 0.0;
        for (int m = 0; m < kMfccNum; ++m) mel = fma(z[m], c_dct[j * kMfccNum + m], mel);
        const double e = exp(mel);
        r[j & 7] = (j < 8) ? e : __dadd_rn(r[j & 7], e);
    }
    const double total = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                   __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    return __ddiv_rn(1.0, total);
}

// One pixel of find_logen: optional float32 min-max normalisation (:672-679), the float64-compute /
// float32-store scaling `mfcc /= lifter; mfcc *= mfnorm` (:310-311), the float64 projection on dct_base^T,
// exp, the band sum in NumPy's order for n = 24 (r[k] = e[k] + e[k+8] + e[k+16], then the balanced tree
// over r[0..7]) and the reciprocal (:313-321).  x[] is left holding the scaled float32 values.
//
// FP64 work is what this stage costs (and, on a board that sits at its power cap, what it costs the HBM stream next to
// it), so the projection uses the symmetry of the basis: cos((m+1) pi (23-j+0.5) / 24) = (-1)^(m+1) cos((m+1) pi (j+0.5) / 24),
// i.e. mel[j] = A_j + B_j and mel[23-j] = A_j - B_j with A over odd m and B over even m - 144 FMAs instead of 288.
// The result differs from a straight 12-term dot product only in the last ulp, like one BLAS differs from another.
// `rare` comes back non-zero when the pixel needs the plain path instead (the caller redoes it with pixel_energy_plain).
__device__ __forceinline__ double pixel_energy(float (&x)[kMfccNum], bool normalize, const FrameNorm& norm,
                                               const double* __restrict__ exp_table, unsigned int& rare) {
    double z[kMfccNum];
    rare = 0;
#pragma unroll
    for (int m = 0; m < kMfccNum; ++m) {
        float v = x[m];
        if (normalize) v = norm.apply(v);                                     // float32, as TF
        rare |= (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u;            // NaN / Inf: plain path
        v = __double2float_rn(div_by_lifter(static_cast<double>(v), m));
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), c_mfnorm));
        x[m] = v;
        z[m] = static_cast<double>(v);
    }
    double r[8], third[8];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < kMfccNum; m += 2) {
            b = fma(z[m], c_dct[j * kMfccNum + m], b);                        // m + 1 odd: antisymmetric in j <-> 23 - j
            a = fma(z[m + 1], c_dct[j * kMfccNum + m + 1], a);                // m + 1 even: symmetric
        }
        const double e_lo = exp_table64(__dadd_rn(a, b), exp_table, rare);   // band j
        const double e_hi = exp_table64(__dadd_rn(a, -b), exp_table, rare);  // band 23 - j
        if (j < 8) {
            r[j] = e_lo;                      // first term of r[j]
            third[7 - j] = e_hi;              // band 23 - j = 16 + (7 - j): third term of r[7 - j]
        } else {
            r[j - 8] = __dadd_rn(r[j - 8], e_lo);          // band j = 8 + (j - 8): second term of r[j - 8]
            r[15 - j] = __dadd_rn(r[15 - j], e_hi);        // band 23 - j = 8 + (15 - j): second term of r[15 - j]
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], third[k]);
    const double total = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                   __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    return __ddiv_rn(1.0, total);
}

// ---- self-tests of the two arithmetic shortcuts above (aig_selftest) -----------------------------------------------
// out[0]: float32 bit patterns v (all 2^32) x 12 lifters where div_by_lifter(v) != __ddiv_rn(v, lifter) as values
//         (NaN == NaN, -0 == +0); out[1]: of those, how many differ after the float32 store the reference applies.
__global__ void selftest_division_kernel(unsigned long long* out) {
    unsigned long long bad = 0, bad_after_store = 0;
    const unsigned long long total = 1ull << 32;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const double v = static_cast<double>(__uint_as_float(static_cast<unsigned int>(i)));
#pragma unroll
        for (int m = 0; m < kMfccNum; ++m) {
            if (!(fabs(v) <= 3.402823466e38)) continue;          // non-finite inputs take the plain path
            const double fast = div_by_lifter(v, m), exact = __ddiv_rn(v, c_lifter[m]);
            const bool same = (fast == exact) || (fast != fast && exact != exact);
            if (!same) {
                ++bad;
                const float a = __double2float_rn(fast), b = __double2float_rn(exact);
                if (!((a == b) || (a != a && b != b))) ++bad_after_store;
            }
        }
    }
    if (bad) atomicAdd(out, bad);
    if (bad_after_store) atomicAdd(out + 1, bad_after_store);
}

// exp_table64 against CUDA's exp() (itself <= 1 ulp) on n points spread over [-700, 700] plus a dense sweep of
// [-12, 12], the range find_logen's inputs produce: out[0] = points differing, out[1] = max difference in ulps,
// out[2] = points compared.
__global__ void selftest_exp_kernel(unsigned long long n, unsigned long long* out) {
    __shared__ double s_exp[64];
    if (threadIdx.x < 64) s_exp[threadIdx.x] = c_exp2_table[threadIdx.x];
    __syncthreads();
    unsigned long long differ = 0, max_ulps = 0, count = 0;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const double u = static_cast<double>(i % n) / static_cast<double>(n);           // [0, 1)
        const double x = (i < n) ? (u * 1400.0 - 700.0) : (u * 24.0 - 12.0);
        unsigned int oob = 0;
        const long long a = __double_as_longlong(exp_table64(x, s_exp, oob)), b = __double_as_longlong(exp(x));
        if (oob) { differ += 1ull << 40; }                      // must never trigger on [-700, 700]
        const unsigned long long d = static_cast<unsigned long long>(a > b ? a - b : b - a);
        differ += d != 0;
        max_ulps = d > max_ulps ? d : max_ulps;
        ++count;
    }
    atomicAdd(out, differ);
    atomicMax(out + 1, max_ulps);
    atomicAdd(out + 2, count);
}

// np.mean over the 1728 doubles of s_map, bit-compatible with NumPy's pairwise summation.  Called by a
// thread group of at least 128 threads (index t within the group); `sync` is the group's barrier.
// Returns the mean in every thread of the group.
template <typename Sync>
__device__ __forceinline__ double frame_mean(const double* s_map, double (*s_part)[8], double* s_leaf,
                                             double* s_mean, int t, Sync sync) {
    if (t < 128) {
        const int leaf = t >> 3, k = t & 7;
        const double* a = s_map + leaf_start(leaf);
        const int len = leaf_len(leaf);
        double r = a[k];
        for (int i = 8; i < len; i += 8) r = __dadd_rn(r, a[i + k]);
        s_part[leaf][k] = r;
    }
    sync();
    if (t < 16) {
        const double* r = s_part[t];
        s_leaf[t] = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                              __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    }
    sync();
    if (t == 0) {
        double s8[8], s4[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s8[i] = __dadd_rn(s_leaf[2 * i], s_leaf[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] = __dadd_rn(s8[2 * i], s8[2 * i + 1]);
        const double sum = __dadd_rn(__dadd_rn(s4[0], s4[1]), __dadd_rn(s4[2], s4[3]));
        *s_mean = __ddiv_rn(sum, static_cast<double>(kFramePixels));
    }
    sync();
    return *s_mean;
}

__global__ void __launch_bounds__(kEnergyThreads, 4)
energy_kernel(const float* __restrict__ images, long long n_frames, int normalize_first,
              float* __restrict__ scaled_out, double* __restrict__ energy_out,
              uint8_t* __restrict__ mask_out, double* __restrict__ mean_out) {
    __shared__ double s_map[kFramePixels];
    __shared__ double s_part[16][8];
    __shared__ double s_leaf[16];
    __shared__ float s_red[2][kEnergyThreads / 32];
    __shared__ double s_mean;
    __shared__ double s_exp[64];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid < 64) s_exp[tid] = c_exp2_table[tid];
    __syncthreads();

    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float* img = images + frame * kFrameValues;
        float lo = 0.f, range = 1.f;
        if (normalize_first) {
            // min and max of the frame; max(x - min) == fl(max - min) because rounding is monotonic
            float mn = CUDART_INF_F, mx = -CUDART_INF_F;
            const float4* img4 = reinterpret_cast<const float4*>(img);
            for (int i = tid; i < kFrameValues / 4; i += kEnergyThreads) {
                const float4 v = __ldg(img4 + i);
                mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
                mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
            }
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; }
            __syncthreads();
            mn = s_red[0][0]; mx = s_red[1][0];
#pragma unroll
            for (int w = 1; w < kEnergyThreads / 32; ++w) {
                mn = fminf(mn, s_red[0][w]);
                mx = fmaxf(mx, s_red[1][w]);
            }
            lo = mn;
            range = __fsub_rn(mx, mn);
        }

        const FrameNorm norm(lo, range);
        // software pipeline: pixel i + 1 is in flight while pixel i goes through its ~600 float64 operations
        const float4* first = reinterpret_cast<const float4*>(img + tid * kMfccNum);
        float4 na = __ldg(first), nb = __ldg(first + 1), nc = __ldg(first + 2);
#pragma unroll 1
        for (int i = 0; i < kPixelsPerThread; ++i) {
            const int p = tid + i * kEnergyThreads;
            if (p >= kFramePixels) break;
            const float4 a = na, b = nb, c = nc;
            if (p + kEnergyThreads < kFramePixels) {
                const float4* src = reinterpret_cast<const float4*>(img + (p + kEnergyThreads) * kMfccNum);
                na = __ldg(src); nb = __ldg(src + 1); nc = __ldg(src + 2);
            }
            float x[kMfccNum] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
            unsigned int rare;
            double en = pixel_energy(x, normalize_first != 0, norm, s_exp, rare);
            float* scaled_dst = scaled_out != nullptr ? scaled_out + frame * kFrameValues + p * kMfccNum : nullptr;
            if (rare) {
                en = pixel_energy_plain(img + p * kMfccNum, scaled_dst, normalize_first != 0, lo, range);
            } else if (scaled_dst != nullptr) {
                float4* dst = reinterpret_cast<float4*>(scaled_dst);
                dst[0] = make_float4(x[0], x[1], x[2], x[3]);
                dst[1] = make_float4(x[4], x[5], x[6], x[7]);
                dst[2] = make_float4(x[8], x[9], x[10], x[11]);
            }
            s_map[p] = en;
            if (energy_out != nullptr) energy_out[frame * kFramePixels + p] = en;
        }
        __syncthreads();

        if (mask_out != nullptr || mean_out != nullptr) {
            const double mean = frame_mean(s_map, s_part, s_leaf, &s_mean, tid, [] { __syncthreads(); });
            if (tid == 0 && mean_out != nullptr) mean_out[frame] = mean;
            if (mask_out != nullptr) {
                for (int p = tid; p < kFramePixels; p += kEnergyThreads)
                    mask_out[frame * kFramePixels + p] = s_map[p] > mean ? 1 : 0;
            }
        }
        __syncthreads();   // s_map / s_red are reused by the next frame
    }
}

// F4 alone: _normalize_acoustic_images_rescaled over frames (float32).
__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ images, long long n_frames, float* __restrict__ out) {
    __shared__ float s_red[2][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float4* img4 = reinterpret_cast<const float4*>(images + frame * kFrameValues);
        float4* out4 = reinterpret_cast<float4*>(out + frame * kFrameValues);
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        for (int i = tid; i < kFrameValues / 4; i += 256) {
            const float4 v = img4[i];
            mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
            mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; }
        __syncthreads();
        mn = s_red[0][0]; mx = s_red[1][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_red[0][w]); mx = fmaxf(mx, s_red[1][w]); }
        const float range = __fsub_rn(mx, mn);
        for (int i = tid; i < kFrameValues / 4; i += 256) {
            float4 v = img4[i];
            v.x = __fdiv_rn(__fsub_rn(v.x, mn), range);
            v.y = __fdiv_rn(__fsub_rn(v.y, mn), range);
            v.z = __fdiv_rn(__fsub_rn(v.z, mn), range);
            v.w = __fdiv_rn(__fsub_rn(v.w, mn), range);
            out4[i] = v;
        }
        __syncthreads();
    }
}

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
#pragma unroll
        for (int i = 0; i < (kFramePixels + kHeatThreads - 1) / kHeatThreads; ++i) {
            const int p = tid + i * kHeatThreads;
            if (p < kFramePixels) s_t[p] = span > 0.0 ? static_cast<float>((e[i] - lo) / span) : 0.f;
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
