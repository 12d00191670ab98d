// Stage 2 kernels: 12-channel acoustic image -> energy map and mean mask (the heat map is heatmap_kernel.cuh).
//
// energy_kernel          F4 + F5 + F7  per-frame min-max (outdoor_data_mfcc.py:672-679), find_logen
//                        (iouenergythreshold.py:294-323) and the mean-threshold mask (:217-219).
// energy_cluster_kernel  the same for small batches: one frame per thread-block cluster of 8 CTAs.
// acivw_kernel           the reference's whole evaluation step (iouenergythreshold.py:213-229) in one launch: real and
// acivw_cluster_kernel   reconstructed image -> two energy maps -> two masks -> (I, U) -> success counts.
//
// One CTA (or cluster) per frame: the per-frame reductions (min, max, mean) are CTA-local.  The energy is
// computed in float64 like the reference (float32-stored in-place scaling, then a float64
// 12x24 projection, exp, band sum in NumPy's pairwise order, reciprocal) so that the
// `map > mean(map)` decision is reproduced bit for bit up to the last-ulp differences of exp()
// and of the BLAS summation order; the float64 mean follows NumPy's pairwise summation tree
// for 1728 elements exactly.  FP64 work per frame is ~2 MFLOP, far below the HBM time of the
// stage-1 stream it follows, and runs on the otherwise idle FP64 pipe.
#pragma once

#include <cooperative_groups.h>

#include "aig_common.cuh"
#include "mel_tables_ref.inc"

namespace aig {

// find_logen's constants (iouenergythreshold.py:304-308), identical to the MFCC constants.
__constant__ double c_dct[kFilterNum * kMfccNum] = AIG_REF_DCT;      // [24][12]
__constant__ double c_lifter[kMfccNum] = AIG_REF_LIFTER;
__constant__ double c_mfnorm = AIG_REF_MFNORM;

static_assert(kFramePixels % 64 == 0, "whole warp pairs drop out of the last round");

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

// The same without the range bookkeeping: the pair kernels bound every exponent once per pixel instead
// (|mel_j| <= sum_m |z_m| because |cos| <= 1; half_pixel sends the pixel to the plain path when that sum exceeds 700).
__device__ __forceinline__ double exp_table64_inrange(double x, const double* __restrict__ table) {
    const double t = __fma_rn(x, c_exp_k[0], c_exp_k[1]);
    const int k = __double2loint(t);
    const double kd = __dadd_rn(t, -c_exp_k[1]);
    double r = __fma_rn(kd, c_exp_k[2], x);
    r = __fma_rn(kd, c_exp_k[3], r);
    double p = __fma_rn(r, c_exp_k[4], c_exp_k[5]);
    p = __fma_rn(p, r, c_exp_k[6]);
    p = __fma_rn(p, r, c_exp_k[7]);
    p = __fma_rn(p, r, c_exp_k[8]);
    p = __fma_rn(__dmul_rn(r, r), p, r);
    const double tj = table[k & 63];
    const double y = __fma_rn(tj, p, tj);
    return __hiloint2double(__double2hiint(y) + ((k >> 6) << 20), __double2loint(y));
}

// The plain form of one pixel (IEEE divisions, exp(), straight 12-term dot products), kept out of line: it serves the
// pixels the fast path cannot (non-finite inputs, |mel| > 700) and is what the fast path is checked against.
// `raw` holds the pixel's 12 input values, already in registers: the image may alias `scaled_dst` (find_logen scales its
// argument in place) or have been written by other warps of the same kernel (fused kernel), so it is never re-read
// here through a read-only path.  scaled_dst may be null.
__device__ __noinline__ double pixel_energy_plain(const float (&raw)[kMfccNum], float* scaled_dst, bool normalize,
                                                  float lo, float range) {
    double z[kMfccNum];
    for (int m = 0; m < kMfccNum; ++m) {
        float v = raw[m];
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
    float sum_abs = 0.f;                // NaN / Inf in, NaN / Inf out
#pragma unroll
    for (int m = 0; m < kMfccNum; ++m) {
        float v = x[m];
        if (normalize) v = norm.apply(v);                                     // float32, as TF
        v = __double2float_rn(div_by_lifter(static_cast<double>(v), m));
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), c_mfnorm));
        x[m] = v;
        sum_abs += fabsf(v);
        z[m] = static_cast<double>(v);
    }
    // Every |mel_j| <= sum_m |z_m| (|cos| <= 1): at most 700 keeps all 24 table exponentials in range; anything else -
    // huge values, Inf, NaN (the comparison is false for NaN) - takes the plain path.  One float32 test per pixel
    // instead of range bookkeeping in each of the 24 exponentials; the pair kernels below use the same criterion, so
    // every kernel sends exactly the same pixels down the plain path.
    rare = !(sum_abs <= 700.f);
    double r[8], third[8];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < kMfccNum; m += 2) {
            b = fma(z[m], c_dct[j * kMfccNum + m], b);                        // m + 1 odd: antisymmetric in j <-> 23 - j
            a = fma(z[m + 1], c_dct[j * kMfccNum + m + 1], a);                // m + 1 even: symmetric
        }
        const double e_lo = exp_table64_inrange(__dadd_rn(a, b), exp_table);   // band j
        const double e_hi = exp_table64_inrange(__dadd_rn(a, -b), exp_table);  // band 23 - j
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


// =====================================================================================================================
// Two threads per pixel (energy_kernel, stage2_kernel, stage2_cluster_kernel)
// =====================================================================================================================
// One thread per pixel needs ~120 registers for the float64 chain of find_logen (12 scaled values, 24 exponentials in
// flight, the NumPy-ordered band sums): 16 warps per SM, a dependent DFMA chain in each, and the FP64 pipe sat at 43 %
// (profiles/r01_ncu_secondary_kernels.csv) with 116 bytes of spills.  Here a pixel is shared by the same lane of two
// adjacent warps ("halves" of a warp pair).  The 24 mel bands fall into four groups of three (j, 23 - j) basis pairs,
//     group g = pairs {g, 7 - g, 8 + g}  ->  bands {g, 23-g, 7-g, 16+g, 8+g, 15-g}  =  all terms of r[g] and r[7-g]
// (r[k] = (e[k] + e[k+8]) + e[k+16] is NumPy's strided partial sum for n = 24), so half 0 (groups 0, 1) produces
// r0, r1, r6, r7 and half 1 (groups 2, 3) r2..r5 without exchanging a single exponential; only the 6 + 6 scaled float32
// channel values and two float64 partial sums (r2 + r3, r4 + r5) cross the pair, through shared memory around two
// 64-thread named barriers.  Every operation keeps the order of pixel_energy(), so the result is bit-identical to the
// one-thread form the fused kernel's energy warps run.  Which half a warp is decides its constants at compile time
// (template parameter, warp-uniform branch): every DCT coefficient stays a constant-bank operand.

// Named barriers with IMMEDIATE ids.  With the id in a register ptxas must assume all 16 hardware barriers are in use
// ("used 16 barriers"), and an SM only has 64: four CTAs per SM however few registers and shared memory they need - the
// first build of these kernels ran at half its intended occupancy for exactly that reason.  NamedBar<1, MAX_ID, T>::sync(id)
// is an if-chain over the ids 1 .. MAX_ID a kernel really uses, each branch a `bar.sync <imm>, <imm>`.
template <int ID, int MAX_ID, int THREADS>
struct NamedBar {
    static __device__ __forceinline__ void sync(int id) {
        if (id == ID) asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory");
        else NamedBar<ID + 1, MAX_ID, THREADS>::sync(id);
    }
};
template <int MAX_ID, int THREADS>
struct NamedBar<MAX_ID, MAX_ID, THREADS> {
    static __device__ __forceinline__ void sync(int) {
        asm volatile("bar.sync %0, %1;" ::"n"(MAX_ID), "n"(THREADS) : "memory");
    }
};

// The per-frame min-max normalisation with the reciprocal hoisted out of the per-value division.  With r = RN(1 / range),
// q0 = RN(d * r), rem = d - q0 * range (exact in one FMA), q = RN(q0 + rem * r) is the correctly rounded quotient
// (Markstein) as long as nothing over- or underflows: taken when range and d are within 2^+-60, IEEE division otherwise
// (zero, denormal, huge, non-finite).  aig_selftest(2) compares it with __fdiv_rn on 2^32 (d, range) pairs.
struct FrameNormFast {
    float lo, range, r;
    bool fast;
    __device__ __forceinline__ FrameNormFast(float lo_, float range_) : lo(lo_), range(range_) {
        const unsigned int e = (__float_as_uint(range_) >> 23) & 0xffu;
        fast = (e - 67u) <= 120u && range_ > 0.f;
        r = fast ? __frcp_rn(range_) : 0.f;
    }
    __device__ __forceinline__ float apply(float x) const {
        const float d = __fsub_rn(x, lo);
        const float q0 = __fmul_rn(d, r);
        // fast: range (hence r) within 2^+-60, so q0 in [2^-40, 2^40] means d within 2^+-100: nothing under- or overflows.
        // Zero, tiny, huge and non-finite differences (comparison false for NaN) take IEEE division.
        if (fast && q0 >= 9.094947e-13f && q0 <= 1.0995116e12f) {
            const float rem = __fmaf_rn(-q0, range, d);
            return __fmaf_rn(rem, r, q0);
        }
        return __fdiv_rn(d, range);
    }
};

// out[0]: (d, range) pairs where FrameNormFast::apply differs from __fdiv_rn(d, range) (as bit patterns, NaN == NaN);
// out[1]: pairs compared; out[2]: pairs that took the fast path.  Ranges: 2^12 values spread over float32's exponent
// range (dense around 2^-4 .. 2^8, where MFCC frames live); d: 2^20 values per range, half of them in [0, range].
__global__ void selftest_norm_kernel(unsigned long long* out) {
    unsigned long long bad = 0, count = 0, fast_taken = 0;
    const unsigned long long total = 1ull << 32;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const unsigned int ri = static_cast<unsigned int>(i >> 20), di = static_cast<unsigned int>(i & 0xfffffu);
        unsigned int h = ri * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // exponent: three quarters of the ranges in [2^-4, 2^8), the rest anywhere (including denormal and huge)
        const unsigned int ex = (ri & 3u) ? (123u + (h >> 8) % 12u) : ((h >> 8) % 255u);
        const float range = __uint_as_float((ex << 23) | (h & 0x7fffffu));
        unsigned int g = (di + 1u) * 2246822519u + ri * 3266489917u;
        g ^= g >> 16; g *= 2654435761u; g ^= g >> 13;
        float d;
        if (di & 1u) d = range * (static_cast<float>(g >> 8) * (1.0f / 16777216.0f));     // in [0, range)
        else d = __uint_as_float(g);                                                       // any bit pattern
        const FrameNormFast norm(0.f, range);
        const float a = norm.apply(d), b = __fdiv_rn(__fsub_rn(d, 0.f), range);
        const float q0 = __fmul_rn(d, norm.r);
        fast_taken += norm.fast && q0 >= 9.094947e-13f && q0 <= 1.0995116e12f;
        bad += !((__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b));
        ++count;
    }
    if (bad) atomicAdd(out, bad);
    atomicAdd(out + 1, count);
    atomicAdd(out + 2, fast_taken);
}

// The constants of the pair kernels live in SHARED memory, not in the constant bank: ptxas hoists loop-invariant
// constant-bank loads out of the pixel loop, and ~100 float64 coefficients overflow the 63 uniform registers into
// spilled vector registers (500 bytes of local-memory traffic per pixel in the first build of these kernels).  Shared
// loads cannot move across the pair barriers inside the loop; all lanes read the same address (one wavefront), two
// coefficients per LDS.128.
struct __align__(16) EnergyTables {
    double dct[12 * kMfccNum];         // rows 0..11 of dct_base^T (the other twelve follow from the symmetry)
    double lifter[kMfccNum];
    double inv_lifter[kMfccNum];
    double exp2[64];                   // 2^(j/64)
    double mfnorm;
};
__device__ __forceinline__ void load_energy_tables(EnergyTables& t, int tid, int threads) {
    for (int i = tid; i < 12 * kMfccNum; i += threads) t.dct[i] = c_dct[i];
    for (int i = tid; i < 64; i += threads) t.exp2[i] = c_exp2_table[i];
    if (tid < kMfccNum) { t.lifter[tid] = c_lifter[tid]; t.inv_lifter[tid] = c_inv_lifter[tid]; }
    if (tid == 0) t.mfnorm = c_mfnorm;
}

// One (j, 23 - j) basis pair: the two exponentials exp(A + B), exp(A - B) of pixel_energy(), same operation order.
template <int J>
__device__ __forceinline__ void band_pair(const double (&z)[kMfccNum], const EnergyTables& tab, double& e_lo, double& e_hi) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int m = 0; m < kMfccNum; m += 2) {
        b = fma(z[m], tab.dct[J * kMfccNum + m], b);
        a = fma(z[m + 1], tab.dct[J * kMfccNum + m + 1], a);
    }
    e_lo = exp_table64_inrange(__dadd_rn(a, b), tab.exp2);       // band J
    e_hi = exp_table64_inrange(__dadd_rn(a, -b), tab.exp2);      // band 23 - J
}

// Group G: r[G] = (e[G] + e[G+8]) + e[G+16] and r[7-G] = (e[7-G] + e[15-G]) + e[23-G].
template <int G>
__device__ __forceinline__ void band_group(const double (&z)[kMfccNum], const EnergyTables& tab, double& r_g, double& r_7g) {
    double lo1, hi1, lo2, hi2, lo3, hi3;
    band_pair<G>(z, tab, lo1, hi1);          // bands G,     23 - G
    band_pair<8 + G>(z, tab, lo3, hi3);      // bands 8 + G, 15 - G
    r_g = __dadd_rn(lo1, lo3);
    band_pair<7 - G>(z, tab, lo2, hi2);      // bands 7 - G, 16 + G
    r_g = __dadd_rn(r_g, hi2);
    r_7g = __dadd_rn(__dadd_rn(lo2, hi3), hi1);
}

struct PairExchange {                  // one per warp pair, in shared memory
    float x[2][6][32];                 // each half's six scaled channel values
    double sum[2][32];                 // half 1's r2 + r3 and r4 + r5
};

// This half's share of one pixel.  raw: channels 6 * HALF .. 6 * HALF + 5.  Returns the energy in half 0 (0 in half 1);
// `scaled` receives this half's float32-stored scaled values (find_logen's side effect), `rare` the pair's combined flag.
template <int HALF, int MAX_BAR>
__device__ __forceinline__ double half_pixel(const float (&raw)[6], bool normalize, const FrameNormFast& norm,
                                             const EnergyTables& tab, PairExchange& ex, int lane, int bar_id,
                                             float (&scaled)[6], unsigned int& rare_out) {
    double z[kMfccNum];
    float own_abs = 0.f;                // sum of |scaled value|: NaN / Inf in, NaN / Inf out
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int m = 6 * HALF + i;
        float v = raw[i];
        if (normalize) v = norm.apply(v);                                     // float32, as TF
        {                                                                     // div_by_lifter with the shared-memory constants
            const double d = static_cast<double>(v), r = tab.inv_lifter[m];
            const double q0 = __dmul_rn(d, r);
            v = __double2float_rn(__fma_rn(__fma_rn(-q0, tab.lifter[m], d), r, q0));
        }
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), tab.mfnorm));
        scaled[i] = v;
        ex.x[HALF][i][lane] = v;
        own_abs += fabsf(v);
        z[m] = static_cast<double>(v);
    }
    NamedBar<1, MAX_BAR, 64>::sync(bar_id);
    float all_abs = own_abs;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float v = ex.x[1 - HALF][i][lane];
        all_abs += fabsf(v);
        z[6 * (1 - HALF) + i] = static_cast<double>(v);
    }
    // Every |mel_j| <= sum_m |z_m| (|cos| <= 1): at most 700 keeps all 24 table exponentials in range; anything else -
    // huge values, Inf, NaN (the comparison is false for NaN) - goes to the plain path.  Both halves see the same 12
    // values, so the flag needs no exchange.
    const unsigned int rare = !(all_abs <= 700.f);
    double ra, rb, rc, rd;
    band_group<2 * HALF>(z, tab, ra, rb);          // half 0: r0, r7     half 1: r2, r5
    band_group<2 * HALF + 1>(z, tab, rc, rd);      // half 0: r1, r6     half 1: r3, r4
    const double s_lo = __dadd_rn(ra, rc);                     // half 0: r0 + r1    half 1: r2 + r3
    const double s_hi = __dadd_rn(rd, rb);                     // half 0: r6 + r7    half 1: r4 + r5
    if (HALF == 1) { ex.sum[0][lane] = s_lo; ex.sum[1][lane] = s_hi; }
    NamedBar<1, MAX_BAR, 64>::sync(bar_id);
    rare_out = rare;
    if (HALF == 1) return 0.0;
    const double total = __dadd_rn(__dadd_rn(s_lo, ex.sum[0][lane]), __dadd_rn(ex.sum[1][lane], s_hi));
    return __ddiv_rn(1.0, total);
}

// Pixels [p_begin, p_end) of one frame by a group of PAIRS warp pairs (thread index gt within the group; named barriers
// bar_base .. bar_base + PAIRS - 1 <= MAX_BAR belong to the pairs).  img / scaled / energy point at the frame; img may alias scaled
// (find_logen's in-place scaling), so neither is read through a read-only path and a pixel's raw values are loaded before
// anything of that pixel is stored.  Leaves map[p - p_begin] = energy.  Pixels the fast path cannot serve (non-finite
// input, |mel| > 700) only get their bit set in rare_bits (zero on entry); the caller runs frame_energy_fixup after a
// group barrier - the out-of-line plain path stays out of this loop and so do the register spills around its call.
template <int PAIRS, int MAX_BAR>
__device__ __forceinline__ void frame_energy_range(const float* img, int p_begin, int p_end, bool normalize,
                                                   const FrameNormFast& norm, float* scaled, double* energy, double* map,
                                                   PairExchange* ex, unsigned int* rare_bits, const EnergyTables& tab, int gt,
                                                   int bar_base) {
    const int warp = gt >> 5, lane = gt & 31, pair = warp >> 1, half = warp & 1;
    constexpr int kStep = 32 * PAIRS;
    const int bar_id = bar_base + pair;
    auto load6 = [&](int p, float (&r)[6]) {
        const float2* s = reinterpret_cast<const float2*>(img + p * kMfccNum + 6 * half);
        const float2 a = s[0], b = s[1], c = s[2];
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = c.x; r[5] = c.y;
    };
#pragma unroll 1
    for (int base = p_begin + pair * 32; base < p_end; base += kStep) {          // pair-uniform trip count
        const int p = base + lane;
        const bool active = p < p_end;
        float raw[6];
        load6(min(p, p_end - 1), raw);
        // next round's line on its way while this one computes (a prefetch, not a register prefetch: the float64 chain
        // needs every register the 8-CTAs-per-SM budget has)
        if (base + kStep < p_end)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(img + min(p + kStep, p_end - 1) * kMfccNum + 6 * half));
        float sc[6];
        unsigned int rare;
        double en;
        if (half == 0) en = half_pixel<0, MAX_BAR>(raw, normalize, norm, tab, ex[pair], lane, bar_id, sc, rare);
        else en = half_pixel<1, MAX_BAR>(raw, normalize, norm, tab, ex[pair], lane, bar_id, sc, rare);
        if (rare) {
            // neither half stores anything of this pixel: frame_energy_fixup redoes it from its raw values
            if (half == 0 && active) atomicOr(&rare_bits[(p - p_begin) >> 5], 1u << ((p - p_begin) & 31));
        } else if (scaled != nullptr && active) {
            float2* dst = reinterpret_cast<float2*>(scaled + p * kMfccNum + 6 * half);
            dst[0] = make_float2(sc[0], sc[1]);
            dst[1] = make_float2(sc[2], sc[3]);
            dst[2] = make_float2(sc[4], sc[5]);
        }
        if (half == 0 && active) {
            map[p - p_begin] = en;
            if (energy != nullptr) energy[p] = en;
        }
    }
}

// Second pass for the flagged pixels: the plain path (IEEE divisions, exp()) from the raw values, which are still intact
// because neither half stored the pixel.  Called by the whole group after a barrier.  Returns (group-uniformly) whether
// any bit was set; the caller then synchronises and clears the words before the next frame.
__device__ __forceinline__ bool frame_energy_fixup(const float* img, int p_begin, int p_end, bool normalize,
                                                   const FrameNormFast& norm, float* scaled, double* energy, double* map,
                                                   unsigned int* rare_bits, int gt, int threads) {
    const int words = (p_end - p_begin + 31) >> 5;
    bool any = false;
    for (int w = 0; w < words; ++w) any |= rare_bits[w] != 0u;
    if (!any) return false;                                        // group-uniform: every thread reads the same words
    for (int i = gt; i < p_end - p_begin; i += threads) {
        if (!((rare_bits[i >> 5] >> (i & 31)) & 1u)) continue;
        const int p = p_begin + i;
        float all[kMfccNum];
#pragma unroll
        for (int m = 0; m < kMfccNum; ++m) all[m] = img[p * kMfccNum + m];
        const double en = pixel_energy_plain(all, scaled != nullptr ? scaled + p * kMfccNum : nullptr, normalize, norm.lo, norm.range);
        map[i] = en;
        if (energy != nullptr) energy[p] = en;
    }
    return true;
}

// min / max of n4 float4 values by a group of `threads` threads, NaN-propagating like tf.reduce_min / reduce_max
// (outdoor_data_mfcc.py:674,677): any NaN in the frame makes both NaN, hence the whole normalised frame.
// red: float [3][threads / 32] scratch.  Result in every thread of the group.
template <typename Sync>
__device__ __forceinline__ void group_minmax(const float* values, int n4, int gt, int threads, float* red, Sync sync,
                                             float& lo, float& hi) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    bool nan = false;
    const float4* v4 = reinterpret_cast<const float4*>(values);
    for (int i = gt; i < n4; i += threads) {
        const float4 v = v4[i];
        mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
        mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    nan = __any_sync(0xffffffffu, nan);
    const int warps = threads >> 5;
    if ((gt & 31) == 0) { red[gt >> 5] = mn; red[warps + (gt >> 5)] = mx; red[2 * warps + (gt >> 5)] = nan ? 1.f : 0.f; }
    sync();
    mn = red[0]; mx = red[warps];
    float bad = red[2 * warps];
    for (int w = 1; w < warps; ++w) { mn = fminf(mn, red[w]); mx = fmaxf(mx, red[warps + w]); bad += red[2 * warps + w]; }
    sync();                                                     // red may be reused right away
    lo = bad != 0.f ? CUDART_NAN_F : mn;
    hi = bad != 0.f ? CUDART_NAN_F : mx;
}

constexpr int kEnergyPairs = 2;                        // warp pairs per frame in the batch kernels: 64 pixels per round, 27 rounds
constexpr int kEnergyThreads = kEnergyPairs * 64;
constexpr int kScoreThresholdsMax = 1024;              // == kMaxThresholds of score_kernel.cuh

template <int PAIRS>
struct EnergyGroupShared {
    double map[kFramePixels];
    double part[16][8];
    double leaf[16];
    double mean;
    PairExchange ex[PAIRS];
    float red[3 * PAIRS * 2];
    unsigned int rare_bits[kFramePixels / 32];
};

// Arguments of the stage-2 kernels.  Slot 0 is the image of aig_energy, or the real image of aig_acivw_batch; slot 1 the
// reconstructed image (GROUPS == 2 only).  Every output pointer is nullable.
struct Stage2Args {
    const float* img[2];
    float* scaled[2];
    double* energy[2];
    uint8_t* mask[2];
    double* mean[2];
    long long n_frames;
    int normalize_first;
    // scoring, GROUPS == 2 (iouenergythreshold.py:224-229)
    const double* thr;
    int k_thr;
    long long* inter;
    long long* uni;
    unsigned long long* pos;
    unsigned long long* num;
};

// GROUPS == 1: energy_kernel (aig_energy).  GROUPS == 2: the ACIVW evaluation step (aig_acivw_batch): threads 0-127 take
// the real image, 128-255 the reconstructed one, both energy maps stay in shared memory, and the masks, their
// intersection / union counts, the IoU and the per-threshold success counts follow in the same CTA - no mask ever
// travels through HBM unless the caller asks for it.
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * kEnergyThreads, 8 / GROUPS)
stage2_kernel(const __grid_constant__ Stage2Args a) {
    __shared__ EnergyGroupShared<kEnergyPairs> sh[GROUPS];
    __shared__ EnergyTables s_tab;
    __shared__ unsigned int s_pos[GROUPS == 2 ? kScoreThresholdsMax : 1];
    __shared__ int s_iu[2];
    __shared__ double s_iou;
    const int tid = threadIdx.x;
    const int group = tid / kEnergyThreads, gt = tid % kEnergyThreads;
    // barriers: GROUPS == 1: 1, 2 = the pairs, 0 = the group; GROUPS == 2: 1, 2 | 3 and 4, 5 | 6 = pairs | group of each image
    constexpr int kMaxBar = GROUPS == 1 ? kEnergyPairs : GROUPS * (kEnergyPairs + 1);
    const int bar_base = 1 + group * (kEnergyPairs + 1), bar_group = bar_base + kEnergyPairs;
    auto group_sync = [&] { if (GROUPS == 1) __syncthreads(); else NamedBar<1, kMaxBar, kEnergyThreads>::sync(bar_group); };
    load_energy_tables(s_tab, tid, GROUPS * kEnergyThreads);
    if (gt < kFramePixels / 32) sh[group].rare_bits[gt] = 0u;
    if (GROUPS == 2) {
        for (int k = tid; k < a.k_thr; k += GROUPS * kEnergyThreads) s_pos[k] = 0;
        if (blockIdx.x == 0 && tid == 0) atomicAdd(a.num, static_cast<unsigned long long>(a.n_frames));   // num += 1 per frame (:229)
    }
    __syncthreads();
    EnergyGroupShared<kEnergyPairs>& g = sh[group];

    for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        const float* img = a.img[group] + frame * kFrameValues;
        float lo = 0.f, hi = 1.f;
        if (a.normalize_first) group_minmax(img, kFrameValues / 4, gt, kEnergyThreads, g.red, group_sync, lo, hi);
        const FrameNormFast norm(lo, __fsub_rn(hi, lo));        // max(x - min) == fl(max - min): rounding is monotonic
        float* scaled = a.scaled[group] ? a.scaled[group] + frame * kFrameValues : nullptr;
        double* energy = a.energy[group] ? a.energy[group] + frame * kFramePixels : nullptr;
        frame_energy_range<kEnergyPairs, kMaxBar>(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy, g.map, g.ex,
                                         g.rare_bits, s_tab, gt, bar_base);
        group_sync();
        if (frame_energy_fixup(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy, g.map, g.rare_bits, gt,
                               kEnergyThreads)) {
            group_sync();
            if (gt < kFramePixels / 32) g.rare_bits[gt] = 0u;
            group_sync();
        }
        if (GROUPS == 2 || a.mask[group] != nullptr || a.mean[group] != nullptr) {
            const double mean = frame_mean(g.map, g.part, g.leaf, &g.mean, gt, group_sync);
            if (gt == 0 && a.mean[group] != nullptr) a.mean[group][frame] = mean;
            if (a.mask[group] != nullptr) {
                for (int p = gt; p < kFramePixels; p += kEnergyThreads)
                    a.mask[group][frame * kFramePixels + p] = g.map[p] > mean ? 1 : 0;
            }
        }
        if (GROUPS == 2) {
            if (tid < 2) s_iu[tid] = 0;
            __syncthreads();
            const double mean_a = sh[0].mean, mean_b = sh[GROUPS - 1].mean;
            int inter = 0, uni = 0;
            for (int p = tid; p < kFramePixels; p += GROUPS * kEnergyThreads) {       // 1728 = 6.75 * 256: whole warps only
                const bool ma = sh[0].map[p] > mean_a, mb = sh[GROUPS - 1].map[p] > mean_b;
                inter += __popc(__ballot_sync(0xffffffffu, ma && mb));
                uni += __popc(__ballot_sync(0xffffffffu, ma || mb));
            }
            if ((tid & 31) == 0) { atomicAdd(&s_iu[0], inter); atomicAdd(&s_iu[1], uni); }
            __syncthreads();
            if (tid == 0) {
                if (a.inter != nullptr) a.inter[frame] = s_iu[0];
                if (a.uni != nullptr) a.uni[frame] = s_iu[1];
                s_iou = __ddiv_rn(static_cast<double>(s_iu[0]), static_cast<double>(s_iu[1]));   // 0 / 0 = NaN: never counts
            }
            __syncthreads();
            const double iou = s_iou;
            for (int k = tid; k < a.k_thr; k += GROUPS * kEnergyThreads)
                if (iou > a.thr[k]) s_pos[k] += 1u;                                   // thread k owns s_pos[k]
        }
        __syncthreads();     // maps, s_iu and s_iou are reused by the next frame
    }
    if (GROUPS == 2) {
        for (int k = tid; k < a.k_thr; k += GROUPS * kEnergyThreads)
            if (s_pos[k] != 0) atomicAdd(a.pos + k, static_cast<unsigned long long>(s_pos[k]));
    }
}

// ---- small batches: one frame (pair) per thread-block cluster -----------------------------------------------------
// Below ~150 frames one CTA per frame leaves most of the 148 SMs idle and a call costs a whole frame's 27 rounds
// (45 us for the reference's batches of 2-16).  Here a cluster of 8 CTAs shares a frame: CTA r takes pixels
// [216 r, 216 r + 216), which are exactly leaves 2r (104 values) and 2r + 1 (112 values) of NumPy's pairwise-sum tree for
// n = 1728, so each CTA reduces its own two leaves and only eight float64 partial sums (plus, when normalising, eight
// min / max pairs, and for the ACIVW step eight (I, U) pairs) cross the cluster through distributed shared memory.
// Every CTA then walks the top of the tree itself, in NumPy's order: the mean - and with it the mask - is bit-identical
// to the one-CTA kernels.
constexpr int kClusterSize = 8;
constexpr int kSlicePixels = kFramePixels / kClusterSize;          // 216
constexpr int kClusterPairs = 4;                                   // 128 pixels per round: two rounds per slice
constexpr int kClusterGroupThreads = kClusterPairs * 64;           // 256
static_assert(kSlicePixels == 104 + 112, "a slice is two leaves of the pairwise tree");

struct ClusterGroupShared {
    double map[kSlicePixels];
    double r8[2][8];
    double leaf[2];
    double mean;
    float lo, hi;
    PairExchange ex[kClusterPairs];
    float red[3 * kClusterPairs * 2];
    unsigned int rare_bits[(kSlicePixels + 31) / 32];
};
struct ClusterMail {               // what the other CTAs of the cluster read
    float mn[2], mx[2];
    double s8[2];
    int inter, uni;
};

template <int GROUPS>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(GROUPS * kClusterGroupThreads, 1)
stage2_cluster_kernel(const __grid_constant__ Stage2Args a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ ClusterGroupShared sh[GROUPS];
    __shared__ ClusterMail mail;
    __shared__ EnergyTables s_tab;
    __shared__ int s_iu[2];
    const int tid = threadIdx.x;
    const int group = tid / kClusterGroupThreads, gt = tid % kClusterGroupThreads;
    const int lane = tid & 31;
    constexpr int kMaxBar = GROUPS == 1 ? kClusterPairs : GROUPS * (kClusterPairs + 1);
    const int bar_base = 1 + group * (kClusterPairs + 1), bar_group = bar_base + kClusterPairs;
    auto group_sync = [&] { if (GROUPS == 1) __syncthreads(); else NamedBar<1, kMaxBar, kClusterGroupThreads>::sync(bar_group); };
    const unsigned int rank = cluster.block_rank();
    const long long n_clusters = gridDim.x / kClusterSize;
    load_energy_tables(s_tab, tid, GROUPS * kClusterGroupThreads);
    if (gt < (kSlicePixels + 31) / 32) sh[group].rare_bits[gt] = 0u;
    if (GROUPS == 2 && blockIdx.x == 0 && tid == 0) atomicAdd(a.num, static_cast<unsigned long long>(a.n_frames));
    __syncthreads();
    ClusterGroupShared& g = sh[group];
    const int p0 = static_cast<int>(rank) * kSlicePixels;

    for (long long frame = blockIdx.x / kClusterSize; frame < a.n_frames; frame += n_clusters) {
        const float* img = a.img[group] + frame * kFrameValues;
        float lo = 0.f, hi = 1.f;
        if (a.normalize_first) {
            group_minmax(img + p0 * kMfccNum, kSlicePixels * kMfccNum / 4, gt, kClusterGroupThreads, g.red, group_sync, lo, hi);
            if (gt == 0) { mail.mn[group] = lo; mail.mx[group] = hi; }
            cluster.sync();
            if (gt < 32) {                                     // lanes 0-7 fetch the eight slices' extremes, NaN-propagating
                float mn = CUDART_INF_F, mx = -CUDART_INF_F;
                bool nan = false;
                if (lane < kClusterSize) {
                    const ClusterMail* remote = cluster.map_shared_rank(&mail, lane);
                    mn = remote->mn[group]; mx = remote->mx[group];
                    nan = (mn != mn) || (mx != mx);
                }
                nan = __any_sync(0xffffffffu, nan);
                mn = warp_min(mn); mx = warp_max(mx);
                if (lane == 0) { g.lo = nan ? CUDART_NAN_F : mn; g.hi = nan ? CUDART_NAN_F : mx; }
            }
            group_sync();
            lo = g.lo; hi = g.hi;
        }
        const FrameNormFast norm(lo, __fsub_rn(hi, lo));
        float* scaled = a.scaled[group] ? a.scaled[group] + frame * kFrameValues : nullptr;
        double* energy = a.energy[group] ? a.energy[group] + frame * kFramePixels : nullptr;
        frame_energy_range<kClusterPairs, kMaxBar>(img, p0, p0 + kSlicePixels, a.normalize_first != 0, norm, scaled, energy, g.map,
                                          g.ex, g.rare_bits, s_tab, gt, bar_base);
        group_sync();
        if (frame_energy_fixup(img, p0, p0 + kSlicePixels, a.normalize_first != 0, norm, scaled, energy, g.map, g.rare_bits,
                               gt, kClusterGroupThreads)) {
            group_sync();
            if (gt < (kSlicePixels + 31) / 32) g.rare_bits[gt] = 0u;
            group_sync();
        }
        // this slice's two leaves of the pairwise tree, then their sum (one of the eight s8 terms of frame_mean)
        if (gt < 16) {
            const int leaf = gt >> 3, k = gt & 7;
            const double* v = g.map + (leaf ? 104 : 0);
            const int len = leaf ? 112 : 104;
            double r = v[k];
            for (int i = 8; i < len; i += 8) r = __dadd_rn(r, v[i + k]);
            g.r8[leaf][k] = r;
        }
        group_sync();
        if (gt < 2) {
            const double* r = g.r8[gt];
            g.leaf[gt] = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                   __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        }
        group_sync();
        if (gt == 0) mail.s8[group] = __dadd_rn(g.leaf[0], g.leaf[1]);
        cluster.sync();
        if (gt < 32) {
            double v = 0.0;
            if (lane < kClusterSize) v = cluster.map_shared_rank(&mail, lane)->s8[group];
            double s8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s8[i] = __shfl_sync(0xffffffffu, v, i);
            if (lane == 0) {
                double s4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] = __dadd_rn(s8[2 * i], s8[2 * i + 1]);
                const double sum = __dadd_rn(__dadd_rn(s4[0], s4[1]), __dadd_rn(s4[2], s4[3]));
                g.mean = __ddiv_rn(sum, static_cast<double>(kFramePixels));
            }
        }
        group_sync();
        const double mean = g.mean;
        if (rank == 0 && gt == 0 && a.mean[group] != nullptr) a.mean[group][frame] = mean;
        if (a.mask[group] != nullptr) {
            for (int p = gt; p < kSlicePixels; p += kClusterGroupThreads)
                a.mask[group][frame * kFramePixels + p0 + p] = g.map[p] > mean ? 1 : 0;
        }
        if (GROUPS == 2) {
            if (tid < 2) s_iu[tid] = 0;
            __syncthreads();
            const double mean_a = sh[0].mean, mean_b = sh[GROUPS - 1].mean;
            if (tid < 224) {                                   // 216 pixels: seven whole warps
                const bool in = tid < kSlicePixels;
                const bool ma = in && sh[0].map[in ? tid : 0] > mean_a, mb = in && sh[GROUPS - 1].map[in ? tid : 0] > mean_b;
                const int inter = __popc(__ballot_sync(0xffffffffu, ma && mb)), uni = __popc(__ballot_sync(0xffffffffu, ma || mb));
                if (lane == 0) { atomicAdd(&s_iu[0], inter); atomicAdd(&s_iu[1], uni); }
            }
            __syncthreads();
            if (tid == 0) { mail.inter = s_iu[0]; mail.uni = s_iu[1]; }
            cluster.sync();
            if (rank == 0 && tid < 32) {
                int inter = 0, uni = 0;
                if (lane < kClusterSize) {
                    const ClusterMail* remote = cluster.map_shared_rank(&mail, lane);
                    inter = remote->inter; uni = remote->uni;
                }
                inter = warp_sum(inter); uni = warp_sum(uni);
                if (lane == 0) {
                    if (a.inter != nullptr) a.inter[frame] = inter;
                    if (a.uni != nullptr) a.uni[frame] = uni;
                }
                const double iou = __ddiv_rn(static_cast<double>(inter), static_cast<double>(uni));
                for (int k = lane; k < a.k_thr; k += 32)
                    if (iou > a.thr[k]) atomicAdd(a.pos + k, 1ull);
            }
        }
        cluster.sync();      // nobody leaves (or overwrites its mail) while a neighbour may still be reading it
    }
}

// F4 alone: _normalize_acoustic_images_rescaled over frames (float32); `images` may alias `out`.
__global__ void __launch_bounds__(256)
normalize_kernel(const float* images, long long n_frames, float* out) {
    __shared__ float s_red[3 * 8];
    const int tid = threadIdx.x;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        const float4* img4 = reinterpret_cast<const float4*>(images + frame * kFrameValues);
        float4* out4 = reinterpret_cast<float4*>(out + frame * kFrameValues);
        float mn, mx;
        group_minmax(images + frame * kFrameValues, kFrameValues / 4, tid, 256, s_red, [] { __syncthreads(); }, mn, mx);
        const float range = __fsub_rn(mx, mn);
        for (int i = tid; i < kFrameValues / 4; i += 256) {
            float4 v = img4[i];
            v.x = __fdiv_rn(__fsub_rn(v.x, mn), range);
            v.y = __fdiv_rn(__fsub_rn(v.y, mn), range);
            v.z = __fdiv_rn(__fsub_rn(v.z, mn), range);
            v.w = __fdiv_rn(__fsub_rn(v.w, mn), range);
            out4[i] = v;
        }
    }
}

}  // namespace aig
