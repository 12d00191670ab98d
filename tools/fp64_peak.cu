// What the FP64 pipe of one B200 SM really delivers: independent DFMA chains alone, and interleaved 1:1 with integer /
// float32 instructions (can the scheduler fill the pipe's idle issue cycle with other work?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak.bin tools/fp64_peak.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int CHAINS, int MIX>     // MIX: 0 = DFMA only, 1 = + one IMAD per DFMA, 2 = + one FFMA per DFMA, 3 = + one LDS per 2 DFMA
__global__ void __launch_bounds__(256) fp64_kernel(double* out, int iters, double a, double b) {
    __shared__ double s_tab[512];
    s_tab[threadIdx.x] = a + threadIdx.x; s_tab[threadIdx.x + 256] = b;
    __syncthreads();
    double x[CHAINS];
    int k[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; k[c] = threadIdx.x * 3 + c; f[c] = threadIdx.x + 0.5f * c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                x[c] = fma(x[c], a, b);
                if (MIX == 1) k[c] = k[c] * 5 + 7;
                if (MIX == 2) f[c] = fmaf(f[c], 1.0001f, 0.5f);
                if (MIX == 3 && (c & 1) == 0) { const double t = s_tab[(k[c] + u) & 511]; x[c] += 0.0 * t; k[c] += 1; }
            }
        }
    }
    double s = 0;
    int ks = 0;
    float fs = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { s += x[c]; ks += k[c]; fs += f[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ks + fs;
}

template <int CHAINS, int MIX>
static void run(const char* name, double* d_out, int sms, int ctas_per_sm) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fp64_kernel<CHAINS, MIX><<<sms * ctas_per_sm, 256>>>(d_out, 10, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    fp64_kernel<CHAINS, MIX><<<sms * ctas_per_sm, 256>>>(d_out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dfma = static_cast<double>(sms) * ctas_per_sm * 256 * iters * 8.0 * CHAINS;
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    printf("%-44s chains %d, %d warps/SM: %7.3f ms  %6.2f TFLOP/s  %5.1f DFMA/clk/SM at %d MHz\n", name, CHAINS, ctas_per_sm * 8, ms,
           2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / sms / (clock_khz * 1e3), clock_khz / 1000);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* d_out;
    CK(cudaMalloc(&d_out, static_cast<size_t>(sms) * 8 * 256 * 8));
    run<8, 0>("DFMA only", d_out, sms, 2);
    run<8, 0>("DFMA only", d_out, sms, 4);
    run<4, 0>("DFMA only", d_out, sms, 2);
    run<2, 0>("DFMA only", d_out, sms, 2);
    run<1, 0>("DFMA only (dependent chain)", d_out, sms, 1);
    run<1, 0>("DFMA only (dependent chain)", d_out, sms, 2);
    run<8, 1>("DFMA + IMAD 1:1", d_out, sms, 2);
    run<8, 2>("DFMA + FFMA 1:1", d_out, sms, 2);
    run<8, 3>("DFMA + (LDS + DADD + IADD) per 2", d_out, sms, 2);
    run<4, 1>("DFMA + IMAD 1:1", d_out, sms, 2);
    return 0;
}
