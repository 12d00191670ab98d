// Read-only HBM ceiling probe: how fast can a B200 stream 29 GB with (a) plain LDG.128 and (b) 1-D bulk TMA copies
// into shared memory?  Context for the MFCC kernel's roofline fraction (MEASURED_PEAKS.json is a read+write copy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/read_peak tools/read_peak.cu && /tmp/read_peak
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(512) ldg_sum(const float4* __restrict__ in, size_t n4, float* out) {
    float acc = 0.f;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4* p = in + i + k * stride;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    for (; i < n4; i += stride) { float4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one producer thread per CTA streams 32 KiB chunks through a 6-stage smem ring with cp.async.bulk; consumers just release
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(64) bulk_stream(const uint8_t* __restrict__ in, size_t n_chunks, float* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long full[STAGES], empty[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    auto wait = [](uint32_t bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    };
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            wait(s32(&empty[stage]), phase ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[stage])), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(smem + stage * CHUNK)), "l"(in + c * CHUNK), "r"(CHUNK), "r"(s32(&full[stage])) : "memory");
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0; float acc = 0.f;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            wait(s32(&full[stage]), phase);
            acc += reinterpret_cast<float*>(smem + stage * CHUNK)[7];
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[stage])) : "memory");
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (acc == 123.456f) *out = acc;
    }
}

// 2-D tensor-map TMA with the MFCC kernel's geometry: [ROWS x 32 floats] boxes, 128B swizzle, SLABS boxes per stage,
// tiles of ROWS spectra x 512 bins walked slab-major.  No compute: one thread releases each stage as soon as it lands.
template <int ROWS, int SLABS, int STAGES>
__global__ void __launch_bounds__(64) tma2d_stream(const __grid_constant__ CUtensorMap tmap, unsigned n_tiles, float* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long full[STAGES], empty[STAGES];
    constexpr int kStage = SLABS * ROWS * 128;
    const uint32_t ring = (s32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    auto wait = [](uint32_t bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    };
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int kb = 0; kb < 16 / SLABS; ++kb) {
                wait(s32(&empty[stage]), phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[stage])), "r"(kStage) : "memory");
                for (int s = 0; s < SLABS; ++s)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(ring + stage * kStage + s * ROWS * 128), "l"(reinterpret_cast<uint64_t>(&tmap)),
                                   "r"(s32(&full[stage])), "r"((kb * SLABS + s) * 32), "r"((int)(t * ROWS)) : "memory");
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int kb = 0; kb < 16 / SLABS; ++kb) {
                wait(s32(&full[stage]), phase);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[stage])) : "memory");
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        if (phase == 77) *out = 1.f;
    }
}

// Same ring, but with the MFCC kernel's consumer structure (ROWS threads, one spectrum each) and selectable work:
//   WORK 0: wait + release only;  1: + 8 LDS.128 per slab (swizzled, summed);  2: + 2 FMA per bin;  3: + 48 B store per row
template <int ROWS, int SLABS, int STAGES, int WORK, bool HINT>
__global__ void __launch_bounds__(ROWS + 32) tma2d_consume(const __grid_constant__ CUtensorMap tmap, unsigned n_tiles, float* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long full[STAGES], empty[STAGES];
    __shared__ __align__(128) float stage_out[ROWS * 12];
    constexpr int kStage = SLABS * ROWS * 128;
    const uint32_t ring = (s32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(ROWS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    auto wait = [](uint32_t bar, uint32_t parity) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }"
                                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    };
    int stage = 0; uint32_t phase = 0;
    if (threadIdx.x == ROWS) {
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int kb = 0; kb < 16 / SLABS; ++kb) {
                wait(s32(&empty[stage]), phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[stage])), "r"(kStage) : "memory");
                for (int s = 0; s < SLABS; ++s) {
                    if (HINT)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                                     ::"r"(ring + stage * kStage + s * ROWS * 128), "l"(reinterpret_cast<uint64_t>(&tmap)),
                                       "r"(s32(&full[stage])), "r"((kb * SLABS + s) * 32), "r"((int)(t * ROWS)), "l"(policy) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                     ::"r"(ring + stage * kStage + s * ROWS * 128), "l"(reinterpret_cast<uint64_t>(&tmap)),
                                       "r"(s32(&full[stage])), "r"((kb * SLABS + s) * 32), "r"((int)(t * ROWS)) : "memory");
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
    } else if (threadIdx.x < ROWS) {
        const uint32_t row_off = threadIdx.x * 128u, sw = (threadIdx.x & 7u) << 4;
        const int lane = threadIdx.x & 31;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                if (s % SLABS == 0) wait(s32(&full[stage]), phase);
                if (WORK >= 1) {
                    const uint32_t slab = ring + stage * kStage + (s % SLABS) * ROWS * 128 + row_off;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                     : "r"(slab + ((uint32_t(j) << 4) ^ sw)));
                        if (WORK >= 2) {
                            a0 = fmaf(v.x, 0.25f, a0); a1 = fmaf(v.x, 0.75f, a1); a2 = fmaf(v.y, 0.5f, a2); a3 = fmaf(v.y, 0.5f, a3);
                            a0 = fmaf(v.z, 0.125f, a0); a1 = fmaf(v.z, 0.875f, a1); a2 = fmaf(v.w, 0.3f, a2); a3 = fmaf(v.w, 0.7f, a3);
                        } else {
                            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
                        }
                    }
                }
                if (s % SLABS == SLABS - 1) {
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[stage])) : "memory");
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (WORK == 3) {            // 3 x STG.128 per thread, 48 B apart between lanes
                float4* o = reinterpret_cast<float4*>(out + (size_t(t) * ROWS + threadIdx.x) * 12);
                o[0] = make_float4(a0, a1, a2, a3); o[1] = make_float4(a1, a2, a3, a0); o[2] = make_float4(a2, a3, a0, a1);
            } else if (WORK == 4) {     // same with .cs (streaming) stores
                float4* o = reinterpret_cast<float4*>(out + (size_t(t) * ROWS + threadIdx.x) * 12);
                __stcs(o, make_float4(a0, a1, a2, a3)); __stcs(o + 1, make_float4(a1, a2, a3, a0)); __stcs(o + 2, make_float4(a2, a3, a0, a1));
            } else if (WORK == 5 || WORK == 6) {   // transpose through smem, lane-contiguous 16 B stores (full 128 B lines per quarter warp)
                float4* so = reinterpret_cast<float4*>(stage_out) + (threadIdx.x >> 5) * 96;
                so[lane * 3 + 0] = make_float4(a0, a1, a2, a3); so[lane * 3 + 1] = make_float4(a1, a2, a3, a0); so[lane * 3 + 2] = make_float4(a2, a3, a0, a1);
                __syncwarp();
                float4* o = reinterpret_cast<float4*>(out + (size_t(t) * ROWS + (threadIdx.x & ~31)) * 12);
#pragma unroll
                for (int k = 0; k < 3; ++k) { if (WORK == 5) o[k * 32 + lane] = so[k * 32 + lane]; else __stcs(o + k * 32 + lane, so[k * 32 + lane]); }
                __syncwarp();
            } else if (WORK == 10 || WORK == 11 || WORK == 12) {   // per-thread stores with an L2 policy: evict_last / evict_first / no_allocate-ish
                uint64_t pol;
                if (WORK == 10) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                else if (WORK == 11) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                else asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
                float* o = out + (size_t(t) * ROWS + threadIdx.x) * 12;
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "l"(pol) : "memory");
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o + 4), "f"(a1), "f"(a2), "f"(a3), "f"(a0), "l"(pol) : "memory");
                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o + 8), "f"(a2), "f"(a3), "f"(a0), "f"(a1), "l"(pol) : "memory");
            } else if (WORK == 8) {     // same stores as 3, but every tile overwrites the CTA's own 6 KiB: stays in L2, ~no DRAM writes
                float4* o = reinterpret_cast<float4*>(out + (size_t(blockIdx.x) * ROWS + threadIdx.x) * 12);
                o[0] = make_float4(a0, a1, a2, a3); o[1] = make_float4(a1, a2, a3, a0); o[2] = make_float4(a2, a3, a0, a1);
            } else if (WORK == 9) {     // a quarter of the bytes: 12 B per spectrum (what bf16-ish outputs would cost)
                float* o = out + (size_t(t) * ROWS + threadIdx.x) * 3;
                o[0] = a0; o[1] = a1; o[2] = a2;
            } else if (WORK == 7) {     // per-warp bulk store of 1536 B from smem (cp.async.bulk.global.shared::cta)
                float4* so = reinterpret_cast<float4*>(stage_out) + (threadIdx.x >> 5) * 96;
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                so[lane * 3 + 0] = make_float4(a0, a1, a2, a3); so[lane * 3 + 1] = make_float4(a1, a2, a3, a0); so[lane * 3 + 2] = make_float4(a2, a3, a0, a1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1536;"
                                 ::"l"(out + (size_t(t) * ROWS + (threadIdx.x & ~31)) * 12), "r"(s32(so)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (a0 + a1 + a2 + a3 == 123.456f) *out = a0;
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int SLABS, int STAGES, int WORK, bool HINT>
void run_consume(EncodeFn enc, uint8_t* buf, size_t bytes, float* out, int ctas_per_sm, cudaEvent_t e0, cudaEvent_t e1) {
    const size_t n_rows = bytes / 2048;
    CUtensorMap map;
    cuuint64_t dims[2] = {512, n_rows}; cuuint64_t strides[1] = {2048};
    cuuint32_t box[2] = {32, (cuuint32_t)ROWS}; cuuint32_t el[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    constexpr int smem = STAGES * SLABS * ROWS * 128 + 1024;
    auto k = tma2d_consume<ROWS, SLABS, STAGES, WORK, HINT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        k<<<148 * ctas_per_sm, ROWS + 32, smem>>>(map, (unsigned)(n_rows / ROWS), out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    fflush(stdout);
    printf("consume [%d x 128B] x %d slabs x %d stages, %d CTAs/SM, work %d, hint %d: %.3f ms  %.0f GB/s  (%s)\n", ROWS, SLABS,
           STAGES, ctas_per_sm, WORK, (int)HINT, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int ROWS, int SLABS, int STAGES>
void run_tma2d(EncodeFn enc, uint8_t* buf, size_t bytes, float* out, int ctas_per_sm, cudaEvent_t e0, cudaEvent_t e1) {
    const size_t n_rows = bytes / 2048;
    CUtensorMap map;
    cuuint64_t dims[2] = {512, n_rows}; cuuint64_t strides[1] = {2048};
    cuuint32_t box[2] = {32, (cuuint32_t)ROWS}; cuuint32_t el[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    constexpr int smem = STAGES * SLABS * ROWS * 128 + 1024;
    cudaFuncSetAttribute(tma2d_stream<ROWS, SLABS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        tma2d_stream<ROWS, SLABS, STAGES><<<148 * ctas_per_sm, 64, smem>>>(map, (unsigned)(n_rows / ROWS), out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    fflush(stdout);
    printf("tma 2d [%d x 128B] x %d slabs x %d stages, %d CTAs/SM (%d KiB in flight/SM): %.3f ms  %.0f GB/s  (%s)\n", ROWS, SLABS,
           STAGES, ctas_per_sm, ctas_per_sm * STAGES * SLABS * ROWS / 8, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t bytes = size_t(29) << 30;
    uint8_t* buf; float* out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, bytes / 2048 * 48 + 64);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(e0);
            ldg_sum<<<148 * blocks_per_sm, 512>>>(reinterpret_cast<const float4*>(buf), bytes / 16, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
        }
        printf("ldg.128 x8 unroll, %d CTAs/SM x 512 thr: %.3f ms  %.0f GB/s\n", blocks_per_sm, best, bytes / best / 1e6);
    }
    {
        constexpr int STAGES = 6, CHUNK = 32768;
        cudaFuncSetAttribute(bulk_stream<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(e0);
            bulk_stream<STAGES, CHUNK><<<148, 64, STAGES * CHUNK>>>(buf, bytes / CHUNK, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
        }
        printf("cp.async.bulk 32 KiB x 6 stages, 1 CTA/SM: %.3f ms  %.0f GB/s  (%s)\n", best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    {
        constexpr int STAGES = 3, CHUNK = 16384;
        cudaFuncSetAttribute(bulk_stream<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(e0);
            bulk_stream<STAGES, CHUNK><<<148 * 4, 64, STAGES * CHUNK>>>(buf, bytes / CHUNK, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
        }
        printf("cp.async.bulk 16 KiB x 3 stages, 4 CTAs/SM: %.3f ms  %.0f GB/s  (%s)\n", best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("driver entry point: %s fn=%p q=%d\n", cudaGetErrorString(ge), fn, (int)q); fflush(stdout);
    if (!fn) return 1;
    EncodeFn enc = (EncodeFn)fn;
    run_tma2d<128, 1, 4>(enc, buf, bytes, out, 3, e0, e1);
    run_tma2d<128, 4, 2>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<192, 4, 2>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<192, 1, 8>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<256, 1, 4>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<32, 16, 3>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<64, 16, 1>(enc, buf, bytes, out, 1, e0, e1);
    run_tma2d<32, 16, 1>(enc, buf, bytes, out, 3, e0, e1);
    run_tma2d<128, 2, 2>(enc, buf, bytes, out, 3, e0, e1);
    run_consume<128, 4, 2, 2, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 3, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 4, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 5, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 6, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 7, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 10, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 11, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 12, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<256, 1, 4, 10, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<256, 1, 4, 3, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 8, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 9, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 3, false>(enc, buf, bytes, out, 1, e0, e1);
    run_consume<128, 4, 2, 2, false>(enc, buf, bytes, out, 1, e0, e1);
    return 0;
}
