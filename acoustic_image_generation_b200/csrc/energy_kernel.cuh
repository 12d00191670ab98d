// Stage 2 kernels: 12-channel acoustic image -> energy map and mean mask (the heat map is heatmap_kernel.cuh).
//
// stage2_kernel<1>          F4 + F5 + F7  per-frame min-max (outdoor_data_mfcc.py:672-679), find_logen
//                           (iouenergythreshold.py:294-323) and the mean-threshold mask (:217-219).
// stage2_cluster_kernel<1>  the same for small batches: one frame per thread-block cluster of 8 CTAs.
// stage2_kernel<2>          the reference's whole evaluation step (iouenergythreshold.py:213-229) in one launch: real and
// stage2_cluster_kernel<2>  reconstructed image -> two energy maps -> two masks -> (I, U) -> success counts.
//
// One CTA (or cluster) per frame: the per-frame reductions (min, max, mean) are CTA-local.  The energy is
// computed in float64 like the reference (float32-stored in-place scaling, then a float64
// 12x24 projection, exp, band sum in NumPy's pairwise order, reciprocal) so that the
// `map > mean(map)` decision is reproduced bit for bit up to the last-ulp differences of exp()
// and of the BLAS summation order; the float64 mean follows NumPy's pairwise summation tree
// for 1728 elements exactly.
//
// Round 2, second pass: ONE THREAD PER PIXEL again, with leaner arithmetic.  tools/energy_lab.cu / energy_lab2.cu
// measured the candidates on a B200 (profiles/r02_energy_lab.txt): two threads per pixel exchanging through shared memory and
// named barriers (round 2, first pass) 10.7 M frames/s with this arithmetic, the two halves in one warp exchanging by
// shuffles 11.5 M, one thread per pixel 13.7 M, and 14.2 M with the input staged by cp.async - against 9.6 M for the
// pair kernel with the old arithmetic.  The FP64 pipe is the bound (tools/fp64_peak.cu: 56 DFMA/clk/SM at best, 0.87 of
// nominal; it needs >= 12 independent operations in flight per scheduler), so what pays is fewer FP64 instructions:
//   * the projection uses both symmetries of the basis: cos((m+1) pi (j+1/2) / 24) is (anti)symmetric under j -> 23 - j
//     by the parity of m + 1 and, for even m + 1, again under j -> 11 - j: 24 band sums from six "couples" (j, 11 - j),
//     j < 6, at 24 operations per couple = 144 (288 for the plain product, 168 in the first pass);
//   * the coefficients carry the factor 2^10 / ln2, so a band sum x is already in units of the exp table's step: its
//     integer part k (by the 1.5 * 2^52 trick) is table index and binary exponent, the remainder r = x - k is exact,
//     and 2^(r / 1024) - 1 needs four Taylor terms with a 1024-entry table of 2^(i / 1024): 8 FP64 operations per
//     exponential (11 in the first pass, ~17 + branches for exp()).  Same error class as before: table entry and the
//     final FMA are rounded once each (tests/test_energy_tables_cpu.py bounds the whole thing in exact arithmetic,
//     aig_selftest(1) compares it with CUDA's exp on the device).
// Per pixel: 412 FP64 instructions (500 before) out of 770 (1310 before).
#pragma once

#include <cooperative_groups.h>

#include "aig_common.cuh"
#include "mel_tables_ref.inc"

namespace aig {

// find_logen's constants (iouenergythreshold.py:304-308), identical to the MFCC constants.
__constant__ double c_dct[kFilterNum * kMfccNum] = AIG_REF_DCT;      // [24][12], the plain path
__constant__ double c_lifter[kMfccNum] = AIG_REF_LIFTER;
__constant__ double c_mfnorm = AIG_REF_MFNORM;
__constant__ double c_inv_lifter[kMfccNum] = AIG_REF_INV_LIFTER;     // RN(1 / lifter[m])

// NumPy pairwise-sum leaves for n = 1728: 1728 -> 864 -> 432 -> 216 -> (104, 112); every leaf is
// summed with 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).
__device__ __forceinline__ int leaf_start(int leaf) { return 216 * (leaf >> 1) + ((leaf & 1) ? 104 : 0); }
__device__ __forceinline__ int leaf_len(int leaf) { return (leaf & 1) ? 112 : 104; }

// v / L for a constant L with r = RN(1 / L): q0 = RN(v * r), e = v - q0 * L (exact in one FMA), q = RN(q0 + e * r).
// By Markstein's theorem q is the correctly rounded quotient for finite v; aig_selftest(0) checks it against
// __ddiv_rn for every finite float32 v and all twelve lifter constants.  Non-finite inputs never reach it: pixel_energy
// sends such pixels down the plain path.
__device__ __forceinline__ double div_by_lifter(double v, int m) {
    const double r = c_inv_lifter[m];
    const double q0 = __dmul_rn(v, r);
    const double e = __fma_rn(-q0, c_lifter[m], v);
    return __fma_rn(e, r, q0);
}

// ---- the tables of the pixel arithmetic --------------------------------------------------------------------------
constexpr int kExpLog2 = AIG_EXP_TABLE_LOG2;                          // 1024 entries
constexpr int kExpEntries = 1 << kExpLog2;
constexpr double kExpMagic = 6755399441055744.0;                     // 1.5 * 2^52: rint by addition
static_assert(kExpLog2 == 10, "the Taylor degree below is chosen for a 1024-entry table");

// Couple j < 6 (bands j, 23 - j, 11 - j, 12 + j): AA over channels m = 3, 7, 11, AB over m = 1, 5, 9, B_j and B_(11-j) over
// m = 0, 2 .. 10; all times 2^10 / ln2 (tools/gen_mel_tables.py, rounded once from the reference's float64 cosines).
struct CoupleCoef { double aa[3], ab[3], b0[6], b1[6]; };
__device__ const double g_couple_coef[6 * 18] = AIG_COUPLE_COEF;
__constant__ double c_exp_poly[4] = AIG_EXP_POLY;                     // (ln2 / 1024)^n / n!, n = 1 .. 4 (uniform-register operands)
__device__ const double g_exp2_table[kExpEntries] = AIG_EXP2_TABLE;   // RN(2^(i / 1024))

// The tables live in SHARED memory in every kernel: the exp table is indexed by data, and shared loads of the projection
// coefficients cannot be hoisted out of the pixel loop (loop-invariant constant-bank loads were, in the first build of
// the pair kernels: ~100 float64 coefficients overflow the uniform registers into spilled vector registers).  All lanes
// read the same coefficient address (one wavefront), two coefficients per LDS.128.
struct __align__(16) EnergyTables {
    CoupleCoef couple[6];
    double lift[kMfccNum][2];          // {1 / lifter[m], lifter[m]}
    double mfnorm, pad;
    double exp2[kExpEntries];
};
static_assert(sizeof(CoupleCoef) == 18 * sizeof(double), "g_couple_coef is copied as is");
__device__ __forceinline__ void load_energy_tables(EnergyTables& t, int tid, int threads) {
    double* couple = reinterpret_cast<double*>(t.couple);
    for (int i = tid; i < 6 * 18; i += threads) couple[i] = g_couple_coef[i];
    for (int i = tid; i < kExpEntries; i += threads) t.exp2[i] = g_exp2_table[i];
    if (tid < kMfccNum) { t.lift[tid][0] = c_inv_lifter[tid]; t.lift[tid][1] = c_lifter[tid]; }
    if (tid == 0) { t.mfnorm = c_mfnorm; t.pad = 0.0; }
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));     // the table is read-only after the kernel's first barrier
    return v;
}

// exp(x * ln2 / 1024) for x in table-step units, |x| < 2^20 (|x ln2 / 1024| <= 700 and more; pixel_energy guarantees it):
// k = rint(x), r = x - k exact, result = 2^(k >> 10) * T[k & 1023] * (1 + p(r)) with p the degree-4 Taylor polynomial of
// 2^(r / 1024) - 1 (truncation 3.7e-20).  `table` is the shared-memory byte address of EnergyTables::exp2.
// EXP_STRIDE: bytes between consecutive table entries (8: EnergyTables::exp2; 64: one of the eight interleaved copies of
// stage2_wide_kernel, `table` then pointing at this lane's copy of entry 0).
template <int EXP_STRIDE = 8>
__device__ __forceinline__ double exp_units(double x, uint32_t table) {
    const double t = __dadd_rn(x, kExpMagic);
    const int k = __double2loint(t);
    const double r = __dadd_rn(x, -__dadd_rn(t, -kExpMagic));
    double p = __fma_rn(c_exp_poly[3], r, c_exp_poly[2]);
    p = __fma_rn(p, r, c_exp_poly[1]);
    p = __fma_rn(p, r, c_exp_poly[0]);
    const double q = __dmul_rn(p, r);                                  // 2^(r / 1024) - 1
    const int idx = k & (kExpEntries - 1);
    const double tj = lds_f64(table + static_cast<uint32_t>(idx) * static_cast<uint32_t>(EXP_STRIDE));
    const double y = __fma_rn(tj, q, tj);
    int hi;                                                            // * 2^(k >> 10): (k - idx) << 10 added to the high word
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(k - idx), "n"(1 << (20 - kExpLog2)), "r"(__double2hiint(y)));
    return __hiloint2double(hi, __double2loint(y));
}

// The plain form of one pixel (IEEE divisions, exp(), straight 12-term dot products), kept out of line: it serves the
// pixels the fast path cannot (non-finite inputs, huge values) and is what the fast path is checked against.
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

// The per-frame min-max normalisation with the reciprocal hoisted out of the per-value division.  With r = RN(1 / range),
// q0 = RN(d * r), rem = d - q0 * range (exact in one FMA), q = RN(q0 + rem * r) is the correctly rounded quotient
// (Markstein) as long as nothing over- or underflows: proven for range within 2^+-60 and q0 in [2^-40, 2^40] or d = 0
// (aig_selftest(2) compares it with __fdiv_rn on 2^32 (d, range) pairs), IEEE division otherwise.
//   mode 2  every value of the frame is inside that domain, no per-value test: d = x - lo lies in {0} u [2^-25 |lo|, range]
//           (a non-zero difference of two floats is at least half an ulp of the smaller one, and lo is the frame's
//           minimum), so |lo| >= 2^-14 range gives q0 >= 2^-40, and range >= 2^-30 keeps rem a normal number;
//   mode 1  sane range but lo too close to zero for that guarantee (an already normalised frame): per-value test;
//   mode 0  zero, denormal, huge or non-finite range: IEEE division (fast == false in apply()).
struct FrameNormFast {
    float lo, range, r;
    int mode;
    bool fast;
    __device__ __forceinline__ FrameNormFast(float lo_, float range_) : lo(lo_), range(range_) {
        const unsigned int e = (__float_as_uint(range_) >> 23) & 0xffu;
        fast = (e - 67u) <= 120u && range_ > 0.f;
        r = fast ? __frcp_rn(range_) : 0.f;
        mode = !fast ? 0 : ((e >= 97u && fabsf(lo_) >= __fmul_rn(range_, 6.103515625e-05f)) ? 2 : 1);
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
    __device__ __forceinline__ float apply_mode2(float x) const {
        const float d = __fsub_rn(x, lo);
        const float q0 = __fmul_rn(d, r);
        return __fmaf_rn(__fmaf_rn(-q0, range, d), r, q0);
    }
};

__device__ __forceinline__ float max_nan_abs(float m, float a) {     // max(m, |a|), NaN if either is
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(m), "f"(fabsf(a)));
    return r;
}

// The four exponentials of couple j: bands j, 23 - j (p) and 11 - j, 12 + j (q), "lo" the first of each pair.
struct CoupleExp { double lo_p, hi_p, lo_q, hi_q; };
template <int EXP_STRIDE = 8>
__device__ __forceinline__ CoupleExp couple_exp(const double (&z)[kMfccNum], const CoupleCoef& c, uint32_t exp_table) {
    double aa = __dmul_rn(z[3], c.aa[0]);                 // m + 1 = 4, 8, 12: symmetric under j -> 23 - j and j -> 11 - j
    aa = __fma_rn(z[7], c.aa[1], aa);
    aa = __fma_rn(z[11], c.aa[2], aa);
    double ab = __dmul_rn(z[1], c.ab[0]);                 // m + 1 = 2, 6, 10: symmetric under j -> 23 - j, antisymmetric under j -> 11 - j
    ab = __fma_rn(z[5], c.ab[1], ab);
    ab = __fma_rn(z[9], c.ab[2], ab);
    double b0 = __dmul_rn(z[0], c.b0[0]), b1 = __dmul_rn(z[0], c.b1[0]);   // odd m + 1: antisymmetric under j -> 23 - j
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        b0 = __fma_rn(z[2 * i], c.b0[i], b0);
        b1 = __fma_rn(z[2 * i], c.b1[i], b1);
    }
    const double a0 = __dadd_rn(aa, ab), a1 = __dadd_rn(aa, -ab);
    CoupleExp e;
    e.lo_p = exp_units<EXP_STRIDE>(__dadd_rn(a0, b0), exp_table);           // band j
    e.hi_p = exp_units<EXP_STRIDE>(__dadd_rn(a0, -b0), exp_table);          // band 23 - j
    e.lo_q = exp_units<EXP_STRIDE>(__dadd_rn(a1, b1), exp_table);           // band 11 - j
    e.hi_q = exp_units<EXP_STRIDE>(__dadd_rn(a1, -b1), exp_table);          // band 12 + j
    return e;
}

// One pixel of find_logen: optional float32 min-max normalisation (:672-679), the float64-compute /
// float32-store scaling `mfcc /= lifter; mfcc *= mfnorm` (:310-311), the float64 projection on dct_base^T,
// exp, the band sum in NumPy's order for n = 24 (r[k] = (e[k] + e[k+8]) + e[k+16], then the balanced tree
// over r[0..7]) and the reciprocal (:313-321).  x[] is left holding the scaled float32 values.
// `rare` comes back non-zero when the pixel needs the plain path instead (the caller redoes it with pixel_energy_plain):
// every |band sum| <= sum_m |z_m| <= 12 max_m |z_m| (|cos| <= 1), so max |z_m| <= 58 keeps all 24 table exponentials in
// range; anything else - huge values, Inf, NaN (max.NaN propagates it, the comparison is then false) - is rare.  Every
// kernel uses this one function, so all of them send exactly the same pixels down the plain path and agree bit for bit.
template <int EXP_STRIDE = 8>
__device__ __forceinline__ double pixel_energy(float (&x)[kMfccNum], bool normalize, const FrameNormFast& norm,
                                               const EnergyTables& tab, uint32_t exp_table, unsigned int& rare) {
    if (normalize) {                                      // float32, as TF; a frame-uniform choice of the division's form
        if (norm.mode == 2) {
#pragma unroll
            for (int m = 0; m < kMfccNum; ++m) x[m] = norm.apply_mode2(x[m]);
        } else {
#pragma unroll
            for (int m = 0; m < kMfccNum; ++m) x[m] = norm.apply(x[m]);
        }
    }
    double z[kMfccNum];
    float big = 0.f;
    const double mfnorm = tab.mfnorm;
#pragma unroll
    for (int m = 0; m < kMfccNum; ++m) {
        const double d = static_cast<double>(x[m]), r = tab.lift[m][0];
        const double q0 = __dmul_rn(d, r);                                                  // div_by_lifter
        float v = __double2float_rn(__fma_rn(__fma_rn(-q0, tab.lift[m][1], d), r, q0));
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), mfnorm));
        x[m] = v;
        big = max_nan_abs(big, v);
        z[m] = static_cast<double>(v);
    }
    rare = !(big <= 58.f);
    // couple j holds e[j], e[23-j], e[11-j], e[12+j];  r[k] = (e[k] + e[k+8]) + e[k+16]
    const CoupleExp c0 = couple_exp<EXP_STRIDE>(z, tab.couple[0], exp_table);   // e0  e23 e11 e12
    const CoupleExp c3 = couple_exp<EXP_STRIDE>(z, tab.couple[3], exp_table);   // e3  e20 e8  e15
    const CoupleExp c4 = couple_exp<EXP_STRIDE>(z, tab.couple[4], exp_table);   // e4  e19 e7  e16
    const double r0 = __dadd_rn(__dadd_rn(c0.lo_p, c3.lo_q), c4.hi_q);    // (e0 + e8)  + e16
    const double r7 = __dadd_rn(__dadd_rn(c4.lo_q, c3.hi_q), c0.hi_p);    // (e7 + e15) + e23
    const double r3 = __dadd_rn(__dadd_rn(c3.lo_p, c0.lo_q), c4.hi_p);    // (e3 + e11) + e19
    const double r4 = __dadd_rn(__dadd_rn(c4.lo_p, c0.hi_q), c3.hi_p);    // (e4 + e12) + e20
    const CoupleExp c1 = couple_exp<EXP_STRIDE>(z, tab.couple[1], exp_table);   // e1  e22 e10 e13
    const CoupleExp c2 = couple_exp<EXP_STRIDE>(z, tab.couple[2], exp_table);   // e2  e21 e9  e14
    const CoupleExp c5 = couple_exp<EXP_STRIDE>(z, tab.couple[5], exp_table);   // e5  e18 e6  e17
    const double r1 = __dadd_rn(__dadd_rn(c1.lo_p, c2.lo_q), c5.hi_q);    // (e1 + e9)  + e17
    const double r6 = __dadd_rn(__dadd_rn(c5.lo_q, c2.hi_q), c1.hi_p);    // (e6 + e14) + e22
    const double r2 = __dadd_rn(__dadd_rn(c2.lo_p, c1.lo_q), c5.hi_p);    // (e2 + e10) + e18
    const double r5 = __dadd_rn(__dadd_rn(c5.lo_p, c1.hi_q), c2.hi_p);    // (e5 + e13) + e21
    const double total = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                                   __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
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

// exp_units against CUDA's exp() (itself <= 1 ulp) on n points spread over [-700, 700] plus a dense sweep of
// [-12, 12], the range find_logen's inputs produce.  The argument is given in table-step units x; its natural value
// u = x ln2 / 1024 is not a double, so the reference is exp(u_hi) (1 + u_lo) with u = u_hi + u_lo from the double-double
// form of ln2 / 1024.  out[0] = points differing, out[1] = max difference in ulps, out[2] = points compared.
__global__ void selftest_exp_kernel(unsigned long long n, unsigned long long* out) {
    __shared__ double s_exp[kExpEntries];
    for (int i = threadIdx.x; i < kExpEntries; i += blockDim.x) s_exp[i] = g_exp2_table[i];
    __syncthreads();
    const uint32_t table = smem_u32(s_exp);
    const double step_hi = AIG_EXP_STEP_HI, step_lo = AIG_EXP_STEP_LO;
    unsigned long long differ = 0, max_ulps = 0, count = 0;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const double f = static_cast<double>(i % n) / static_cast<double>(n);           // [0, 1)
        const double nat = (i < n) ? (f * 1400.0 - 700.0) : (f * 24.0 - 12.0);
        const double x = nat / step_hi;                                                  // some double in table-step units
        const double u_hi = __dmul_rn(x, step_hi);
        const double u_lo = __fma_rn(x, step_lo, __fma_rn(x, step_hi, -u_hi));
        const double e = exp(u_hi);
        const long long a = __double_as_longlong(exp_units(x, table)), b = __double_as_longlong(__fma_rn(e, u_lo, e));
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
// thread group of `threads` threads (index t within the group); `sync` is the group's barrier.
// Returns the mean in every thread of the group.
template <typename Sync>
__device__ __forceinline__ double frame_mean(const double* s_map, double (*s_part)[8], double* s_leaf,
                                             double* s_mean, int t, int threads, Sync sync) {
    for (int u = t; u < 128; u += threads) {
        const int leaf = u >> 3, k = u & 7;
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
// The pixel loops and kernels
// =====================================================================================================================

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

// Pixels p_begin + gt, + THREADS, ... < p_end of one frame, one thread per pixel (gt = thread index within the group of
// THREADS).  img / scaled / energy point at the frame; img may alias scaled (find_logen's in-place scaling), so neither
// is read through a read-only path and a pixel's raw values are loaded before anything of that pixel is stored.  Leaves
// map[p - p_begin] = energy.  Pixels the fast path cannot serve (non-finite input, huge values) only get their bit set in
// rare_bits (zero on entry); the caller runs frame_energy_fixup after a group barrier - the out-of-line plain path stays
// out of this loop and so do the register spills around its call.
// STAGED: the next pixel's 48 bytes travel into this thread's own shared-memory slots (cp.async) while the current pixel
// is computed - a plain load at the top of the round leaves 13 % of the warps' time waiting for L2 (ncu, energy_lab2),
// registers for a software prefetch do not exist (122 are in use), and a prefetch hint only helps half way.
// EXP_STRIDE / exp_table: see exp_units (0 = the table of `tab`).
template <int THREADS, bool STAGED, int EXP_STRIDE = 8>
__device__ __forceinline__ void frame_energy_pixels(const float* img, int p_begin, int p_end, bool normalize,
                                                    const FrameNormFast& norm, float* scaled, double* energy, double* map,
                                                    unsigned int* rare_bits, const EnergyTables& tab, float4 (*stage)[THREADS],
                                                    int gt, uint32_t exp_table = 0) {
    if (exp_table == 0) exp_table = smem_u32(tab.exp2);
    auto stage_in = [&](int p) {
        const float* src = img + p * kMfccNum;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&stage[c][gt])), "l"(src + 4 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (STAGED && p_begin + gt < p_end) stage_in(p_begin + gt);
#pragma unroll 1
    for (int p = p_begin + gt; p < p_end; p += THREADS) {
        float4 a, b, c;
        if (STAGED) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            a = stage[0][gt]; b = stage[1][gt]; c = stage[2][gt];
            if (p + THREADS < p_end) stage_in(p + THREADS);
        } else {
            const float4* src = reinterpret_cast<const float4*>(img + p * kMfccNum);
            a = src[0]; b = src[1]; c = src[2];
            // (also a compiler barrier: without one the coefficient loads are hoisted out of the loop into 700 bytes of spills)
            if (p + THREADS < p_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(img + (p + THREADS) * kMfccNum) : "memory");
            else asm volatile("" ::: "memory");
        }
        float x[kMfccNum] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        unsigned int rare;
        const double en = pixel_energy<EXP_STRIDE>(x, normalize, norm, tab, exp_table, rare);
        if (rare) {
            // nothing of this pixel is stored: frame_energy_fixup redoes it from its raw values
            atomicOr(&rare_bits[(p - p_begin) >> 5], 1u << ((p - p_begin) & 31));
            continue;
        }
        if (scaled != nullptr) {
            float4* dst = reinterpret_cast<float4*>(scaled + p * kMfccNum);
            dst[0] = make_float4(x[0], x[1], x[2], x[3]);
            dst[1] = make_float4(x[4], x[5], x[6], x[7]);
            dst[2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        map[p - p_begin] = en;
        if (energy != nullptr) energy[p] = en;
    }
}

// Second pass for the flagged pixels: the plain path (IEEE divisions, exp()) from the raw values, which are still intact
// because nothing of the pixel was stored.  Called by the whole group after a barrier.  Returns (group-uniformly) whether
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
// (outdoor_data_mfcc.py:674,677): any NaN in the frame makes both NaN, hence the whole normalised frame.  min.NaN /
// max.NaN carry the NaN through the reduction themselves (no separate flag: half the instructions of the pass).
// red: float [2][threads / 32] scratch.  Result in every thread of the group.
__device__ __forceinline__ float min_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float max_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
template <typename Sync>
__device__ __forceinline__ void group_minmax(const float* values, int n4, int gt, int threads, float* red, Sync sync,
                                             float& lo, float& hi) {
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    const float4* v4 = reinterpret_cast<const float4*>(values);
    auto take = [&](const float4& v) {
        mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
        mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
    };
    int i = gt;
    for (; i + 7 * threads < n4; i += 8 * threads) {       // eight loads in flight: a 64-thread group would otherwise wait 81 times
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = v4[i + u * threads];
#pragma unroll
        for (int u = 0; u < 8; ++u) take(v[u]);
    }
    for (; i < n4; i += threads) take(v4[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warps = threads >> 5;
    if ((gt & 31) == 0) { red[gt >> 5] = mn; red[warps + (gt >> 5)] = mx; }
    sync();
    mn = red[0]; mx = red[warps];
    for (int w = 1; w < warps; ++w) { mn = min_nan(mn, red[w]); mx = max_nan(mx, red[warps + w]); }
    sync();                                                     // red may be reused right away
    lo = mn;
    hi = mx;
}

constexpr int kEnergyThreads = 64;                     // threads per frame in the batch kernels: 27 rounds of 64 pixels
constexpr int kScoreThresholdsMax = 1024;              // == kMaxThresholds of score_kernel.cuh
static_assert(kFramePixels % kEnergyThreads == 0, "whole rounds only");

struct EnergyGroupShared {
    double map[kFramePixels];
    union {                            // the staging slots are idle while the mean is formed
        float4 stage[3][kEnergyThreads];
        struct { double part[16][8]; double leaf[16]; } sum;
    };
    double mean;
    float red[3 * (kEnergyThreads / 32)];
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

// GROUPS == 1: aig_energy, 64 threads per CTA, 8 CTAs per SM (26 KB of shared memory and 122 registers each).
// GROUPS == 2: the ACIVW evaluation step (aig_acivw_batch): threads 0-63 take the real image, 64-127 the reconstructed
// one, both energy maps stay in shared memory, and the masks, their intersection / union counts, the IoU and the
// per-threshold success counts follow in the same CTA - no mask ever travels through HBM unless the caller asks for it.
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * kEnergyThreads, 8 / GROUPS)
stage2_kernel(const __grid_constant__ Stage2Args a) {
    __shared__ EnergyGroupShared sh[GROUPS];
    __shared__ EnergyTables s_tab;
    __shared__ unsigned int s_pos[GROUPS == 2 ? kScoreThresholdsMax : 1];
    __shared__ int s_iu[2];
    __shared__ double s_iou;
    const int tid = threadIdx.x;
    const int group = tid / kEnergyThreads, gt = tid % kEnergyThreads;
    // barriers: 0 = the CTA; GROUPS == 2: 1, 2 = the two images' groups
    auto group_sync = [&] { if (GROUPS == 1) __syncthreads(); else NamedBar<1, GROUPS, kEnergyThreads>::sync(1 + group); };
    load_energy_tables(s_tab, tid, GROUPS * kEnergyThreads);
    if (gt < kFramePixels / 32) sh[group].rare_bits[gt] = 0u;
    if (GROUPS == 2) {
        for (int k = tid; k < a.k_thr; k += GROUPS * kEnergyThreads) s_pos[k] = 0;
        if (blockIdx.x == 0 && tid == 0) atomicAdd(a.num, static_cast<unsigned long long>(a.n_frames));   // num += 1 per frame (:229)
    }
    __syncthreads();
    EnergyGroupShared& g = sh[group];

    for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        const float* img = a.img[group] + frame * kFrameValues;
        // (asking this CTA's next image into L2 here - prefetch.global.L2, 10 lines per thread - was measured: 14.05 M frames/s
        // against 14.5 M, and 8.15 M against 8.48 M in the warp-specialised energy + heat-map kernel: these kernels do not wait
        // on DRAM, the float64 pipe is their bound)
        float lo = 0.f, hi = 1.f;
        if (a.normalize_first) group_minmax(img, kFrameValues / 4, gt, kEnergyThreads, g.red, group_sync, lo, hi);
        const FrameNormFast norm(lo, __fsub_rn(hi, lo));        // max(x - min) == fl(max - min): rounding is monotonic
        float* scaled = a.scaled[group] ? a.scaled[group] + frame * kFrameValues : nullptr;
        double* energy = a.energy[group] ? a.energy[group] + frame * kFramePixels : nullptr;
        frame_energy_pixels<kEnergyThreads, true>(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy, g.map,
                                                  g.rare_bits, s_tab, g.stage, gt);
        group_sync();
        if (frame_energy_fixup(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy, g.map, g.rare_bits, gt,
                               kEnergyThreads)) {
            group_sync();
            if (gt < kFramePixels / 32) g.rare_bits[gt] = 0u;
            group_sync();
        }
        if (GROUPS == 2 || a.mask[group] != nullptr || a.mean[group] != nullptr) {
            const double mean = frame_mean(g.map, g.sum.part, g.sum.leaf, &g.mean, gt, kEnergyThreads, group_sync);
            if (gt == 0 && a.mean[group] != nullptr) a.mean[group][frame] = mean;
            if (a.mask[group] != nullptr) {
                uint8_t* mask = a.mask[group] + frame * kFramePixels;
                if ((reinterpret_cast<uintptr_t>(mask) & 3u) == 0) {                  // four pixels per store
                    uint32_t* dst = reinterpret_cast<uint32_t*>(mask);
                    for (int q = gt; q < kFramePixels / 4; q += kEnergyThreads) {
                        const double* e = g.map + 4 * q;
                        dst[q] = (e[0] > mean ? 1u : 0u) | (e[1] > mean ? 0x100u : 0u) | (e[2] > mean ? 0x10000u : 0u) |
                                 (e[3] > mean ? 0x1000000u : 0u);
                    }
                } else {
                    for (int p = gt; p < kFramePixels; p += kEnergyThreads) mask[p] = g.map[p] > mean ? 1 : 0;
                }
            }
        }
        if (GROUPS == 2) {
            if (tid < 2) s_iu[tid] = 0;
            __syncthreads();
            const double mean_a = sh[0].mean, mean_b = sh[GROUPS - 1].mean;
            int inter = 0, uni = 0;
            for (int p = tid; p < kFramePixels; p += GROUPS * kEnergyThreads) {       // 1728 = 13.5 * 128: whole warps only
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
                if (iou > a.thr[k]) s_pos[k] += 1u;                                   // thread k % 128 owns s_pos[k]
        }
        __syncthreads();     // maps, s_iu and s_iou are reused by the next frame
    }
    if (GROUPS == 2) {
        for (int k = tid; k < a.k_thr; k += GROUPS * kEnergyThreads)
            if (s_pos[k] != 0) atomicAdd(a.pos + k, static_cast<unsigned long long>(s_pos[k]));
    }
}

// ---- aig_energy, large batches: eight frames per CTA around one conflict-free exponential table ----------------------
// ncu on stage2_kernel<1> (round 2): FP64 pipe 50 %, but the shared-memory pipe 66 % of its peak over the whole launch -
// 287 wavefronts per warp and pixel against 206 cycles of float64 pipe: 132 for the 66 broadcast coefficient loads, 48
// for the 24 exponential-table look-ups and 79 more for their bank conflicts (16 lanes of a half warp picking 16 random
// doubles collide ~3-fold).  The shared-memory pipe, not the float64 pipe, is what the kernel runs into.  Eight
// interleaved copies of the table (entry i of copy c at (8 i + c) doubles; a lane reads copy lane % 8, so the 16 lanes
// of a half warp meet at most in pairs) take the conflicts away, but cost 64 KB: affordable once per SM, not once per
// 64-thread CTA.  Hence one CTA of 512 threads per SM: eight groups of 64 threads, each with its own frame, energy map,
// staging slots and named barrier - the same 16 warps at 128 registers as eight CTAs of stage2_kernel<1> - and one table.
constexpr int kWideGroups = 8;
constexpr int kExpCopies = 8;
template <int GROUPS>
struct WideShared {
    EnergyTables tab;                                  // coefficients; its one-copy exponential table serves the plain path only
    double exp2x[kExpEntries][kExpCopies];
    EnergyGroupShared group[kWideGroups];
    unsigned int pos[GROUPS == 2 ? kScoreThresholdsMax : 1];
    int iu[kWideGroups / GROUPS][2];
    double iou[kWideGroups / GROUPS];
};
// GROUPS == 1: aig_energy, a unit is a frame (eight per CTA).  GROUPS == 2: aig_acivw_batch, a unit is a pair of images -
// real and reconstructed, 128 threads, four pairs per CTA - scored in the same CTA as in stage2_kernel<2>; the
// per-threshold success counts of the four pairs meet in one shared vector (shared-memory atomics).
template <int GROUPS>
__global__ void __launch_bounds__(kWideGroups * kEnergyThreads, 1)
stage2_wide_kernel(const __grid_constant__ Stage2Args a) {
    extern __shared__ __align__(16) unsigned char s_wide_raw[];
    WideShared<GROUPS>& ws = *reinterpret_cast<WideShared<GROUPS>*>(s_wide_raw);
    constexpr int kUnits = kWideGroups / GROUPS, kUnitThreads = GROUPS * kEnergyThreads;
    const int tid = threadIdx.x;
    const int slot = tid / kEnergyThreads, gt = tid % kEnergyThreads;     // slot: which of the CTA's eight 64-thread groups
    const int unit = slot / GROUPS, group = slot % GROUPS, ut = tid % kUnitThreads;
    // barriers: 0 = the CTA, 1 .. 8 = the groups, 9 .. 12 = the pairs (GROUPS == 2)
    auto group_sync = [&] { asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(kEnergyThreads) : "memory"); };
    auto unit_sync = [&] {
        if (GROUPS == 1) group_sync();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + kWideGroups + unit), "n"(kUnitThreads) : "memory");
    };
    load_energy_tables(ws.tab, tid, kWideGroups * kEnergyThreads);
    for (int i = tid; i < kExpEntries * kExpCopies; i += kWideGroups * kEnergyThreads)
        ws.exp2x[i / kExpCopies][i % kExpCopies] = g_exp2_table[i / kExpCopies];
    EnergyGroupShared& g = ws.group[slot];
    if (gt < kFramePixels / 32) g.rare_bits[gt] = 0u;
    if (GROUPS == 2) {
        for (int k = tid; k < a.k_thr; k += kWideGroups * kEnergyThreads) ws.pos[k] = 0;
        if (blockIdx.x == 0 && tid == 0) atomicAdd(a.num, static_cast<unsigned long long>(a.n_frames));   // num += 1 per frame (:229)
    }
    __syncthreads();
    const uint32_t exp_table = smem_u32(&ws.exp2x[0][tid % kExpCopies]);

    // Unit u of CTA b takes frames u G + b, (kUnits + u) G + b, ... (G CTAs): a last, partial round then leaves every CTA
    // with about the same number of busy groups (taking kUnits consecutive frames per CTA instead would leave whole SMs
    // idle for the round: 4096 frames 13 % slower), and a batch smaller than 8 frames per SM still occupies every SM.
    for (long long frame = static_cast<long long>(unit) * gridDim.x + blockIdx.x; frame < a.n_frames;
         frame += static_cast<long long>(gridDim.x) * kUnits) {
        const float* img = a.img[group] + frame * kFrameValues;
        float lo = 0.f, hi = 1.f;
        if (a.normalize_first) group_minmax(img, kFrameValues / 4, gt, kEnergyThreads, g.red, group_sync, lo, hi);
        const FrameNormFast norm(lo, __fsub_rn(hi, lo));
        float* scaled = a.scaled[group] ? a.scaled[group] + frame * kFrameValues : nullptr;
        double* energy = a.energy[group] ? a.energy[group] + frame * kFramePixels : nullptr;
        frame_energy_pixels<kEnergyThreads, true, 8 * kExpCopies>(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy,
                                                                 g.map, g.rare_bits, ws.tab, g.stage, gt, exp_table);
        group_sync();
        if (frame_energy_fixup(img, 0, kFramePixels, a.normalize_first != 0, norm, scaled, energy, g.map, g.rare_bits, gt,
                               kEnergyThreads)) {
            group_sync();
            if (gt < kFramePixels / 32) g.rare_bits[gt] = 0u;
            group_sync();
        }
        if (GROUPS == 2 || a.mask[group] != nullptr || a.mean[group] != nullptr) {
            const double mean = frame_mean(g.map, g.sum.part, g.sum.leaf, &g.mean, gt, kEnergyThreads, group_sync);
            if (gt == 0 && a.mean[group] != nullptr) a.mean[group][frame] = mean;
            if (a.mask[group] != nullptr) {
                uint8_t* mask = a.mask[group] + frame * kFramePixels;
                if ((reinterpret_cast<uintptr_t>(mask) & 3u) == 0) {                  // four pixels per store
                    uint32_t* dst = reinterpret_cast<uint32_t*>(mask);
                    for (int q = gt; q < kFramePixels / 4; q += kEnergyThreads) {
                        const double* e = g.map + 4 * q;
                        dst[q] = (e[0] > mean ? 1u : 0u) | (e[1] > mean ? 0x100u : 0u) | (e[2] > mean ? 0x10000u : 0u) |
                                 (e[3] > mean ? 0x1000000u : 0u);
                    }
                } else {
                    for (int p = gt; p < kFramePixels; p += kEnergyThreads) mask[p] = g.map[p] > mean ? 1 : 0;
                }
            }
        }
        if (GROUPS == 2) {
            const EnergyGroupShared& ga = ws.group[unit * GROUPS], & gb = ws.group[unit * GROUPS + GROUPS - 1];
            if (ut < 2) ws.iu[unit][ut] = 0;
            unit_sync();                                                              // both maps and means are complete
            const double mean_a = ga.mean, mean_b = gb.mean;
            int inter = 0, uni = 0;
            for (int p = ut; p < kFramePixels; p += kUnitThreads) {                   // 1728 = 13.5 * 128: whole warps only
                const bool ma = ga.map[p] > mean_a, mb = gb.map[p] > mean_b;
                inter += __popc(__ballot_sync(0xffffffffu, ma && mb));
                uni += __popc(__ballot_sync(0xffffffffu, ma || mb));
            }
            if ((ut & 31) == 0) { atomicAdd(&ws.iu[unit][0], inter); atomicAdd(&ws.iu[unit][1], uni); }
            unit_sync();
            if (ut == 0) {
                if (a.inter != nullptr) a.inter[frame] = ws.iu[unit][0];
                if (a.uni != nullptr) a.uni[frame] = ws.iu[unit][1];
                ws.iou[unit] = __ddiv_rn(static_cast<double>(ws.iu[unit][0]), static_cast<double>(ws.iu[unit][1]));   // 0 / 0 = NaN: never counts
            }
            unit_sync();
            const double iou = ws.iou[unit];
            for (int k = ut; k < a.k_thr; k += kUnitThreads)
                if (iou > a.thr[k]) atomicAdd(&ws.pos[k], 1u);
        }
        unit_sync();     // the maps, the staging slots, iu and iou are reused by the unit's next frame
    }
    if (GROUPS == 2) {
        __syncthreads();
        for (int k = tid; k < a.k_thr; k += kWideGroups * kEnergyThreads)
            if (ws.pos[k] != 0) atomicAdd(a.pos + k, static_cast<unsigned long long>(ws.pos[k]));
    }
}

// ---- small batches: one frame (pair) per thread-block cluster -----------------------------------------------------
// Below ~150 frames one CTA per frame leaves most of the 148 SMs idle and a call costs a whole frame's 27 rounds.
// Here a cluster of 8 CTAs shares a frame: CTA r takes pixels [216 r, 216 r + 216), one pixel per thread in a single
// round; these are exactly leaves 2r (104 values) and 2r + 1 (112 values) of NumPy's pairwise-sum tree for
// n = 1728, so each CTA reduces its own two leaves and only eight float64 partial sums (plus, when normalising, eight
// min / max pairs, and for the ACIVW step eight (I, U) pairs) cross the cluster through distributed shared memory.
// Every CTA then walks the top of the tree itself, in NumPy's order: the mean - and with it the mask - is bit-identical
// to the one-CTA kernels.
constexpr int kClusterSize = 8;
constexpr int kSlicePixels = kFramePixels / kClusterSize;          // 216
constexpr int kClusterGroupThreads = 224;                          // seven warps: one pixel per thread, 8 lanes idle
static_assert(kSlicePixels == 104 + 112, "a slice is two leaves of the pairwise tree");
static_assert(kClusterGroupThreads >= kSlicePixels && kClusterGroupThreads % 32 == 0, "one round");

struct ClusterGroupShared {
    double map[kSlicePixels];
    double r8[2][8];
    double leaf[2];
    double mean;
    float lo, hi;
    float red[3 * (kClusterGroupThreads / 32)];
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
    auto group_sync = [&] { if (GROUPS == 1) __syncthreads(); else NamedBar<1, GROUPS, kClusterGroupThreads>::sync(1 + group); };
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
        frame_energy_pixels<kClusterGroupThreads, false>(img, p0, p0 + kSlicePixels, a.normalize_first != 0, norm, scaled, energy,
                                                         g.map, g.rare_bits, s_tab, nullptr, gt);
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

// The same with the frame resident in shared memory: one bulk asynchronous copy brings a frame's 82 944 bytes in, the CTA
// finds min / max and normalises it in place, one bulk copy takes it out - HBM sees each byte once in each direction
// (normalize_kernel's second pass over the frame re-reads it through L2, and its per-thread loads keep too few bytes in
// flight: 0.58 of the copy roof).  Two frame buffers per CTA, one CTA per SM: the next frame's load and the previous
// frame's store run while the current one is computed.  Needs 16-byte aligned buffers; `images` may alias `out`.
constexpr int kNormBulkThreads = 1024;
constexpr uint32_t kFrameBytes = kFrameValues * sizeof(float);
constexpr size_t kNormBulkSmem = 2 * static_cast<size_t>(kFrameBytes) + 128;
static_assert(kFrameBytes % 16 == 0, "bulk copies move 16-byte units");

__global__ void __launch_bounds__(kNormBulkThreads, 1)
normalize_bulk_kernel(const float* images, long long n_frames, float* out) {
    extern __shared__ __align__(128) unsigned char s_frames[];
    __shared__ unsigned long long s_bar[2];
    __shared__ float s_red[3 * (kNormBulkThreads / 32)];
    const int tid = threadIdx.x;
    float* buf[2] = {reinterpret_cast<float*>(s_frames), reinterpret_cast<float*>(s_frames + kFrameBytes)};
    const uint32_t bar[2] = {smem_u32(&s_bar[0]), smem_u32(&s_bar[1])};
    if (tid == 0) {
        mbar_init(bar[0], 1);
        mbar_init(bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    long long frame = blockIdx.x;
    if (tid == 0 && frame < n_frames) {
        mbar_arrive_expect_tx(bar[0], kFrameBytes);
        bulk_load_g2s(smem_u32(buf[0]), images + frame * kFrameValues, kFrameBytes, bar[0]);
    }
    for (unsigned int it = 0; frame < n_frames; frame += gridDim.x, ++it) {
        const unsigned int s = it & 1u;
        const long long next = frame + gridDim.x;
        if (tid == 0 && next < n_frames) {
            bulk_wait_read<0>();                               // the store of the previous frame has finished reading buf[s ^ 1]
            mbar_arrive_expect_tx(bar[s ^ 1u], kFrameBytes);
            bulk_load_g2s(smem_u32(buf[s ^ 1u]), images + next * kFrameValues, kFrameBytes, bar[s ^ 1u]);
        }
        mbar_wait(bar[s], (it >> 1) & 1u);
        float mn, mx;
        group_minmax(buf[s], kFrameValues / 4, tid, kNormBulkThreads, s_red, [] { __syncthreads(); }, mn, mx);
        const FrameNormFast norm(mn, __fsub_rn(mx, mn));
        float4* v4 = reinterpret_cast<float4*>(buf[s]);
        if (norm.mode == 2) {
            for (int i = tid; i < kFrameValues / 4; i += kNormBulkThreads) {
                float4 v = v4[i];
                v.x = norm.apply_mode2(v.x); v.y = norm.apply_mode2(v.y); v.z = norm.apply_mode2(v.z); v.w = norm.apply_mode2(v.w);
                v4[i] = v;
            }
        } else {
            for (int i = tid; i < kFrameValues / 4; i += kNormBulkThreads) {
                float4 v = v4[i];
                v.x = norm.apply(v.x); v.y = norm.apply(v.y); v.z = norm.apply(v.z); v.w = norm.apply(v.w);
                v4[i] = v;
            }
        }
        fence_proxy_async_smem();                              // the bulk copy reads what the threads just wrote
        __syncthreads();
        if (tid == 0) {
            bulk_store_s2g(out + frame * kFrameValues, smem_u32(buf[s]), kFrameBytes);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all<0>();                          // shared memory must outlive the copies that read it
}

}  // namespace aig
