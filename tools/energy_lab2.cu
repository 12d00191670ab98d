// Design lab, part 2: one thread per pixel (the winner of energy_lab.cu) with exp-table geometries, hand-written table
// addressing and input prefetch variants.  Each variant is checked against a plain IEEE reference kernel and timed.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/energy_lab2.bin tools/energy_lab2.cu
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kPixels = 1728, kCh = 12, kFrameValues = kPixels * kCh;

__constant__ double c_dct[24 * 12];
__constant__ double c_lifter[12], c_mfnorm;

__global__ void reference_kernel(const float* img, long long n_pixels, int normalize, const float* lohi, float* scaled, double* energy) {
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (p >= n_pixels) return;
    const long long frame = p / kPixels;
    const float lo = lohi[2 * frame], range = __fsub_rn(lohi[2 * frame + 1], lo);
    double z[12];
    for (int m = 0; m < 12; ++m) {
        float v = img[p * 12 + m];
        if (normalize) v = __fdiv_rn(__fsub_rn(v, lo), range);
        v = __double2float_rn(__ddiv_rn(static_cast<double>(v), c_lifter[m]));
        v = __double2float_rn(__dmul_rn(static_cast<double>(v), c_mfnorm));
        scaled[p * 12 + m] = v;
        z[m] = v;
    }
    double r[8];
    for (int j = 0; j < 24; ++j) {
        double mel = 0.0;
        for (int m = 0; m < 12; ++m) mel = fma(z[m], c_dct[j * 12 + m], mel);
        const double e = exp(mel);
        r[j & 7] = (j < 8) ? e : __dadd_rn(r[j & 7], e);
    }
    const double total = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    energy[p] = __ddiv_rn(1.0, total);
}

__global__ void minmax_kernel(const float* img, float* lohi) {
    __shared__ float s[2][8];
    const float* f = img + static_cast<long long>(blockIdx.x) * kFrameValues;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int i = threadIdx.x; i < kFrameValues; i += blockDim.x) { mn = fminf(mn, f[i]); mx = fmaxf(mx, f[i]); }
    for (int o = 16; o; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(~0u, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(~0u, mx, o)); }
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = mn; s[1][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < blockDim.x / 32; ++w) { mn = fminf(mn, s[0][w]); mx = fmaxf(mx, s[1][w]); }
        lohi[2 * blockIdx.x] = mn; lohi[2 * blockIdx.x + 1] = mx;
    }
}

// Tables of one exp geometry: 2^LOGN table entries, each replicated 1024 >> LOGN times (entry j of replica r at
// [j * REP + r]), projection coefficients scaled by 2^LOGN / ln2 so that the band sums come out in table-index units.
struct CoupleCoef {           // couple (j, 11 - j), j < 6
    double aa[3];             // channels m = 3, 7, 11 at j
    double ab[3];             // m = 1, 5, 9
    double b0[6];             // m = 0, 2, .., 10 at j
    double b1[6];             // ... at 11 - j
};
struct LabTables {
    CoupleCoef couple[6];     // j = 0 .. 5
    double lift[12][2];       // {1 / lifter, lifter}
    double poly[6];           // (ln2 / 2^LOGN)^i / i!, i = 1 .. 6
    double mfnorm, pad;
    double exp2[1024];
};
__device__ LabTables g_tables[3];         // LOGN = 6, 8, 10
__constant__ CoupleCoef c_couple10[6];    // LOGN = 10 coefficients in the constant bank (COEF variants)

static void fill_tables(int logn, LabTables& out) {
    const long double ln2 = 0.693147180559945309417232121458176568L;
    const int n = 1 << logn, rep = 1024 >> logn;
    const long double scale = n / ln2;
    double dct[24][12];
    for (int j = 0; j < 24; ++j)
        for (int m = 0; m < 12; ++m) dct[j][m] = std::cos((m + 1) * M_PI / 24 * (j + 0.5));
    for (int j = 0; j < 6; ++j) {
        for (int i = 0; i < 3; ++i) {
            out.couple[j].aa[i] = static_cast<double>(dct[j][3 + 4 * i] * scale);
            out.couple[j].ab[i] = static_cast<double>(dct[j][1 + 4 * i] * scale);
        }
        for (int i = 0; i < 6; ++i) {
            out.couple[j].b0[i] = static_cast<double>(dct[j][2 * i] * scale);
            out.couple[j].b1[i] = static_cast<double>(dct[11 - j][2 * i] * scale);
        }
    }
    for (int m = 0; m < 12; ++m) { const double L = 1 + 11.0 * std::sin(M_PI * (m + 1) / 22); out.lift[m][0] = 1.0 / L; out.lift[m][1] = L; }
    long double u = ln2 / n, f = 1;
    for (int i = 1; i <= 6; ++i) { f *= u / i; out.poly[i - 1] = static_cast<double>(f); }
    out.mfnorm = std::sqrt(2.0 / 24);
    out.pad = 0;
    for (int j = 0; j < n; ++j)
        for (int r = 0; r < rep; ++r) out.exp2[j * rep + r] = static_cast<double>(powl(2.0L, static_cast<long double>(j) / n));
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// exp(x * ln2 / 2^LOGN); tab_lane = shared-memory byte address of this lane's replica of entry 0.
template <int LOGN, bool HAND>
__device__ __forceinline__ double exp_scaled(double x, uint32_t tab_lane, const double (&poly)[6]) {
    constexpr int N = 1 << LOGN, REP = 1024 >> LOGN, DEG = LOGN == 6 ? 6 : (LOGN == 8 ? 5 : 4);
    const double magic = 6755399441055744.0;
    const double t = __dadd_rn(x, magic);
    const int k = __double2loint(t);
    const double kd = __dadd_rn(t, -magic);
    const double r = __dadd_rn(x, -kd);
    double p = poly[DEG - 1];
#pragma unroll
    for (int i = DEG - 2; i >= 0; --i) p = __fma_rn(p, r, poly[i]);
    const double q = __dmul_rn(p, r);
    const int idx = k & (N - 1);
    double tj, y;
    if (HAND) {
        tj = lds_f64(tab_lane + static_cast<uint32_t>(idx) * (REP * 8));
        y = __fma_rn(tj, q, tj);
        int hi;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(k - idx), "n"(1 << (20 - LOGN)), "r"(__double2hiint(y)));
        return __hiloint2double(hi, __double2loint(y));
    }
    tj = lds_f64(tab_lane + static_cast<uint32_t>(idx) * (REP * 8));
    y = __fma_rn(tj, q, tj);
    return __hiloint2double(__double2hiint(y) + ((k >> LOGN) << 20), __double2loint(y));
}

struct Couple { double lo_p, hi_p, lo_q, hi_q; };     // exp of bands j, 23 - j, 11 - j, 12 + j

template <int LOGN, bool HAND>
__device__ __forceinline__ Couple couple(const double (&z)[12], const CoupleCoef& c, uint32_t tab_lane, const double (&poly)[6]) {
    // (c may point into shared memory or into the constant bank)
    double aa = __dmul_rn(z[3], c.aa[0]);
    aa = __fma_rn(z[7], c.aa[1], aa);
    aa = __fma_rn(z[11], c.aa[2], aa);
    double ab = __dmul_rn(z[1], c.ab[0]);
    ab = __fma_rn(z[5], c.ab[1], ab);
    ab = __fma_rn(z[9], c.ab[2], ab);
    double b0 = __dmul_rn(z[0], c.b0[0]), b1 = __dmul_rn(z[0], c.b1[0]);
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        b0 = __fma_rn(z[2 * i], c.b0[i], b0);
        b1 = __fma_rn(z[2 * i], c.b1[i], b1);
    }
    const double a0 = __dadd_rn(aa, ab), a1 = __dadd_rn(aa, -ab);
    Couple r;
    r.lo_p = exp_scaled<LOGN, HAND>(__dadd_rn(a0, b0), tab_lane, poly);
    r.hi_p = exp_scaled<LOGN, HAND>(__dadd_rn(a0, -b0), tab_lane, poly);
    r.lo_q = exp_scaled<LOGN, HAND>(__dadd_rn(a1, b1), tab_lane, poly);
    r.hi_q = exp_scaled<LOGN, HAND>(__dadd_rn(a1, -b1), tab_lane, poly);
    return r;
}

__device__ __forceinline__ float max_nan_abs(float m, float a) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(m), "f"(fabsf(a)));
    return r;
}

struct LabArgs {
    const float* img;
    long long n_frames;
    int normalize;
    const float* lohi;
    float* scaled;
    double* energy;
    unsigned long long* rare_count;
};

// PREFETCH: 0 = prefetch.global.L1 hint, 1 = next pixel's values in registers, 2 = nothing,
//           3 = the next pixel's 48 bytes on their way into a per-thread shared-memory slot (cp.async) during this pixel's arithmetic
template <int LOGN, bool HAND, int PREFETCH, int THREADS, int MINB, int COEF = 0>
__global__ void __launch_bounds__(THREADS, MINB) lab_single_kernel(const LabArgs a) {
    __shared__ __align__(16) LabTables tab;
    __shared__ double s_map[kPixels];
    __shared__ float4 s_stage[PREFETCH == 3 ? 3 : 1][THREADS];
    constexpr int TI = LOGN == 6 ? 0 : (LOGN == 8 ? 1 : 2), REP = 1024 >> LOGN;
    {
        const double* src = reinterpret_cast<const double*>(&g_tables[TI]);
        double* dst = reinterpret_cast<double*>(&tab);
        for (int i = threadIdx.x; i < static_cast<int>(sizeof(LabTables) / 8); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t tab_lane = static_cast<uint32_t>(__cvta_generic_to_shared(tab.exp2)) + (lane & (REP - 1)) * 8;
    const double mfnorm = tab.mfnorm;
    double poly[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) poly[i] = tab.poly[i];
    unsigned long long rare_seen = 0;
    for (long long frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        const float* img = a.img + frame * kFrameValues;
        float lo = 0.f, range = 1.f, rcp = 1.f;
        if (a.normalize) { lo = a.lohi[2 * frame]; range = __fsub_rn(a.lohi[2 * frame + 1], lo); rcp = __frcp_rn(range); }
        float4 n0, n1, n2;
        if (PREFETCH == 3) {
            const float* src = img + threadIdx.x * kCh;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&s_stage[c][threadIdx.x]))), "l"(src + 4 * c) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (PREFETCH == 1) {
            const float4* s = reinterpret_cast<const float4*>(img + threadIdx.x * kCh);
            n0 = s[0]; n1 = s[1]; n2 = s[2];
        }
#pragma unroll 1
        for (int p = threadIdx.x; p < kPixels; p += THREADS) {
            float4 x0, x1, x2;
            if (PREFETCH == 1) {
                x0 = n0; x1 = n1; x2 = n2;
                if (p + THREADS < kPixels) {
                    const float4* s = reinterpret_cast<const float4*>(img + (p + THREADS) * kCh);
                    n0 = s[0]; n1 = s[1]; n2 = s[2];
                }
            } else if (PREFETCH == 3) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                x0 = s_stage[0][threadIdx.x]; x1 = s_stage[1][threadIdx.x]; x2 = s_stage[2][threadIdx.x];
                if (p + THREADS < kPixels) {
                    const float* src = img + (p + THREADS) * kCh;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&s_stage[c][threadIdx.x]))), "l"(src + 4 * c) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
            } else {
                const float4* s = reinterpret_cast<const float4*>(img + p * kCh);
                x0 = s[0]; x1 = s[1]; x2 = s[2];
                if (PREFETCH == 0 && p + THREADS < kPixels) asm volatile("prefetch.global.L1 [%0];" ::"l"(img + (p + THREADS) * kCh));
            }
            const float raw[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
            float v[12];
            double z[12];
            float big = 0.f;
#pragma unroll
            for (int m = 0; m < 12; ++m) {
                float x = raw[m];
                if (a.normalize) {
                    const float d = __fsub_rn(x, lo);
                    const float q0 = __fmul_rn(d, rcp);
                    x = __fmaf_rn(__fmaf_rn(-q0, range, d), rcp, q0);
                }
                const double dd = static_cast<double>(x), inv = tab.lift[m][0], L = tab.lift[m][1];
                const double q0 = __dmul_rn(dd, inv);
                const float f1 = __double2float_rn(__fma_rn(__fma_rn(-q0, L, dd), inv, q0));
                v[m] = __double2float_rn(__dmul_rn(static_cast<double>(f1), mfnorm));
                big = max_nan_abs(big, v[m]);
                z[m] = static_cast<double>(v[m]);
            }
            const bool rare = !(big <= 58.f);
            int zero = 0;
            if (COEF == 2) asm volatile("mov.u32 %0, 0;" : "=r"(zero));
            // r[k] = (e[k] + e[k+8]) + e[k+16];  lo(j) = e[j], hi(j) = e[23-j] for j < 12
            //   couple j holds lo(j) = e[j], hi(j) = e[23-j], lo(11-j) = e[11-j], hi(11-j) = e[12+j]
            Couple c0 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[0] : (COEF == 1 ? c_couple10[0] : c_couple10[0 + zero]), tab_lane, poly);   // e0 e23 e11 e12
            Couple c3 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[3] : (COEF == 1 ? c_couple10[3] : c_couple10[3 + zero]), tab_lane, poly);   // e3 e20 e8  e15
            Couple c4 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[4] : (COEF == 1 ? c_couple10[4] : c_couple10[4 + zero]), tab_lane, poly);   // e4 e19 e7  e16
            const double r0 = __dadd_rn(__dadd_rn(c0.lo_p, c3.lo_q), c4.hi_q);   // e0 + e8 + e16
            const double r7 = __dadd_rn(__dadd_rn(c4.lo_q, c3.hi_q), c0.hi_p);   // e7 + e15 + e23
            const double r3 = __dadd_rn(__dadd_rn(c3.lo_p, c0.lo_q), c4.hi_p);   // e3 + e11 + e19
            const double r4 = __dadd_rn(__dadd_rn(c4.lo_p, c0.hi_q), c3.hi_p);   // e4 + e12 + e20
            Couple c1 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[1] : (COEF == 1 ? c_couple10[1] : c_couple10[1 + zero]), tab_lane, poly);   // e1 e22 e10 e13
            Couple c2 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[2] : (COEF == 1 ? c_couple10[2] : c_couple10[2 + zero]), tab_lane, poly);   // e2 e21 e9  e14
            Couple c5 = couple<LOGN, HAND>(z, COEF == 0 ? tab.couple[5] : (COEF == 1 ? c_couple10[5] : c_couple10[5 + zero]), tab_lane, poly);   // e5 e18 e6  e17
            const double r1 = __dadd_rn(__dadd_rn(c1.lo_p, c2.lo_q), c5.hi_q);   // e1 + e9 + e17
            const double r6 = __dadd_rn(__dadd_rn(c5.lo_q, c2.hi_q), c1.hi_p);   // e6 + e14 + e22
            const double r2 = __dadd_rn(__dadd_rn(c2.lo_p, c1.lo_q), c5.hi_p);   // e2 + e10 + e18
            const double r5 = __dadd_rn(__dadd_rn(c5.lo_p, c1.hi_q), c2.hi_p);   // e5 + e13 + e21
            const double total = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
            if (rare) { ++rare_seen; continue; }
            if (a.scaled != nullptr) {
                float4* dst = reinterpret_cast<float4*>(a.scaled + frame * kFrameValues + p * kCh);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]); dst[1] = make_float4(v[4], v[5], v[6], v[7]); dst[2] = make_float4(v[8], v[9], v[10], v[11]);
            }
            const double en = __ddiv_rn(1.0, total);
            s_map[p] = en;
            a.energy[frame * kPixels + p] = en;
        }
        __syncthreads();
    }
    if (rare_seen) atomicAdd(a.rare_count, rare_seen);
}

static double lcg(unsigned long long& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (s >> 11) * (1.0 / 9007199254740992.0); }

template <typename F>
static float time_ms(F launch, int reps = 9) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

struct Variant { const char* name; void (*launch)(const LabArgs&, int sms); };
#define V(name, LOGN, HAND, PF, THREADS, MINB) {name, [](const LabArgs& a, int sms) { lab_single_kernel<LOGN, HAND, PF, THREADS, MINB><<<sms * MINB, THREADS>>>(a); }}
#define VC(name, LOGN, HAND, PF, THREADS, MINB, COEF) {name, [](const LabArgs& a, int sms) { lab_single_kernel<LOGN, HAND, PF, THREADS, MINB, COEF><<<sms * MINB, THREADS>>>(a); }}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 8192;
    const int only = argc > 2 ? atoi(argv[2]) : -1;
    const long long n_check = n < 64 ? n : 64;
    LabTables* lt = new LabTables[3];
    fill_tables(6, lt[0]); fill_tables(8, lt[1]); fill_tables(10, lt[2]);
    CK(cudaMemcpyToSymbol(g_tables, lt, 3 * sizeof(LabTables)));
    CK(cudaMemcpyToSymbol(c_couple10, lt[2].couple, sizeof(CoupleCoef) * 6));
    double dct[24][12], lifter[12], mfnorm = std::sqrt(2.0 / 24);
    for (int j = 0; j < 24; ++j) for (int m = 0; m < 12; ++m) dct[j][m] = std::cos((m + 1) * M_PI / 24 * (j + 0.5));
    for (int m = 0; m < 12; ++m) lifter[m] = 1 + 11.0 * std::sin(M_PI * (m + 1) / 22);
    CK(cudaMemcpyToSymbol(c_dct, dct, sizeof dct));
    CK(cudaMemcpyToSymbol(c_lifter, lifter, sizeof lifter));
    CK(cudaMemcpyToSymbol(c_mfnorm, &mfnorm, sizeof(double)));

    std::vector<float> h_img(static_cast<size_t>(n) * kFrameValues);
    unsigned long long seed = 12345;
    for (auto& v : h_img) v = static_cast<float>(lcg(seed) * 70.0 - 45.0);
    float *d_img, *d_lohi, *d_scaled, *d_scaled_ref;
    double *d_energy, *d_energy_ref;
    unsigned long long* d_rare;
    CK(cudaMalloc(&d_img, h_img.size() * 4)); CK(cudaMalloc(&d_lohi, n * 8));
    CK(cudaMalloc(&d_scaled, h_img.size() * 4)); CK(cudaMalloc(&d_scaled_ref, n_check * kFrameValues * 4));
    CK(cudaMalloc(&d_energy, n * kPixels * 8)); CK(cudaMalloc(&d_energy_ref, n_check * kPixels * 8));
    CK(cudaMalloc(&d_rare, 8));
    CK(cudaMemcpy(d_img, h_img.data(), h_img.size() * 4, cudaMemcpyHostToDevice));
    minmax_kernel<<<static_cast<unsigned>(n), 256>>>(d_img, d_lohi);
    CK(cudaDeviceSynchronize());
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    const Variant variants[] = {
        VC("1024x1 deg4, cp.async, coefficients: constant bank", 10, true, 3, 64, 8, 1),
        VC("1024x1 deg4, cp.async, constant bank, opaque index", 10, true, 3, 64, 8, 2),
        V("1024x1 deg4, hand, cp.async staging,      8x64", 10, true, 3, 64, 8),
        V("1024x1 deg4, hand, cp.async staging,      9x64", 10, true, 3, 64, 9),
        V("1024x1 deg4, hand, cp.async staging,      4x128", 10, true, 3, 128, 4),
        V("1024x1 deg4, hand, cp.async staging,      7x64", 10, true, 3, 64, 7),
        V("64x16 deg6, compiler addressing, L1 hint, 8x64", 6, false, 0, 64, 8),
        V("64x16 deg6, hand addressing,     L1 hint, 8x64", 6, true, 0, 64, 8),
        V("64x16 deg6, hand, register prefetch,      8x64", 6, true, 1, 64, 8),
        V("64x16 deg6, hand, no prefetch,            8x64", 6, true, 2, 64, 8),
        V("256x4 deg5, hand, L1 hint,                8x64", 8, true, 0, 64, 8),
        V("1024x1 deg4, hand, L1 hint,               8x64", 10, true, 0, 64, 8),
        V("1024x1 deg4, hand, register prefetch,     8x64", 10, true, 1, 64, 8),
        V("64x16 deg6, hand, L1 hint,                6x64", 6, true, 0, 64, 6),
        V("64x16 deg6, hand, L1 hint,                10x64", 6, true, 0, 64, 10),
        V("64x16 deg6, hand, L1 hint,                4x128", 6, true, 0, 128, 4),
        V("64x16 deg6, hand, L1 hint,                2x192", 6, true, 0, 192, 2),
    };
    for (int normalize = 1; normalize >= 0; --normalize) {
        reference_kernel<<<static_cast<unsigned>((n_check * kPixels + 255) / 256), 256>>>(d_img, n_check * kPixels, normalize, d_lohi, d_scaled_ref, d_energy_ref);
        CK(cudaDeviceSynchronize());
        std::vector<float> ref_scaled(n_check * kFrameValues), got_scaled(n_check * kFrameValues);
        std::vector<double> ref_energy(n_check * kPixels), got_energy(n_check * kPixels);
        CK(cudaMemcpy(ref_scaled.data(), d_scaled_ref, ref_scaled.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ref_energy.data(), d_energy_ref, ref_energy.size() * 8, cudaMemcpyDeviceToHost));
        for (int vi = 0; vi < static_cast<int>(sizeof variants / sizeof variants[0]); ++vi) {
            if (only >= 0 && vi != only) continue;
            LabArgs a{d_img, n, normalize, d_lohi, d_scaled, d_energy, d_rare};
            CK(cudaMemset(d_rare, 0, 8));
            CK(cudaMemset(d_energy, 0, n * kPixels * 8));
            variants[vi].launch(a, sms);
            CK(cudaDeviceSynchronize());
            CK(cudaGetLastError());
            unsigned long long rare = 0;
            CK(cudaMemcpy(&rare, d_rare, 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(got_scaled.data(), d_scaled, got_scaled.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(got_energy.data(), d_energy, got_energy.size() * 8, cudaMemcpyDeviceToHost));
            long long bad_scaled = 0, identical = 0;
            double max_rel = 0;
            for (size_t i = 0; i < ref_scaled.size(); ++i) bad_scaled += memcmp(&ref_scaled[i], &got_scaled[i], 4) != 0;
            for (size_t i = 0; i < ref_energy.size(); ++i) {
                const double rel = std::fabs(got_energy[i] - ref_energy[i]) / std::fabs(ref_energy[i]);
                if (!(rel <= max_rel)) max_rel = rel;
                identical += got_energy[i] == ref_energy[i];
            }
            const float ms = time_ms([&] { variants[vi].launch(a, sms); });
            printf("normalize=%d  [%2d] %-50s %7.3f ms  %6.2f M frames/s   scaled bad %lld  energy max rel %.2e  identical %.1f%%  rare %llu\n",
                   normalize, vi, variants[vi].name, ms, n / ms / 1e3, bad_scaled, max_rel, 100.0 * identical / ref_energy.size(), rare);
        }
    }
    return 0;
}
