// Stage 3 kernels: mask IoU / consensus IoU and the all-thresholds success-count sweep.
//
// iou_sweep_kernel   F8  iouenergythreshold.py:224-229   one warp per frame pair
// ciou_sweep_kernel  F9  showimages_bb.py:288-321        one CTA per frame
//
// Intersection / union are integer counts (popcounts of byte-compare masks, integer box
// coverage), the ratio is one float64 division, and `iou > thr[k]` is evaluated for all K
// thresholds in the same pass (the reference re-runs the whole evaluation per threshold).
// Per-threshold successes are first accumulated in shared memory per CTA and flushed with one
// 64-bit atomicAdd per (CTA, k) - the count vector is what NCCL all-reduces across GPUs.
#pragma once

#include "aig_common.cuh"
#include "heatmap_kernel.cuh"   // linear_tap_exact / MaskTaps

namespace aig {

constexpr int kMaxThresholds = 1024;
constexpr int kIouThreads = 256;
constexpr int kMaskWords = kFramePixels / 4;     // 432 x 4 bytes

__global__ void __launch_bounds__(kIouThreads)
iou_sweep_kernel(const uint8_t* __restrict__ mask_a, const uint8_t* __restrict__ mask_b, long long n,
                 const double* __restrict__ thr, int k_thr, long long* __restrict__ inter_out,
                 long long* __restrict__ union_out, unsigned long long* __restrict__ pos,
                 unsigned long long* __restrict__ num) {
    __shared__ unsigned int s_pos[kMaxThresholds];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < k_thr; k += kIouThreads) s_pos[k] = 0;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(num, static_cast<unsigned long long>(n));   // num += 1 per frame (:229)
    __syncthreads();
    const long long warps_total = static_cast<long long>(gridDim.x) * (kIouThreads / 32);
    for (long long f = static_cast<long long>(blockIdx.x) * (kIouThreads / 32) + warp; f < n; f += warps_total) {
        const uint32_t* a = reinterpret_cast<const uint32_t*>(mask_a + f * kFramePixels);
        const uint32_t* b = reinterpret_cast<const uint32_t*>(mask_b + f * kFramePixels);
        int inter = 0, uni = 0;
        for (int w = lane; w < kMaskWords; w += 32) {
            const uint32_t ma = __vcmpne4(__ldg(a + w), 0u);      // 0xFF per non-zero byte
            const uint32_t mb = __vcmpne4(__ldg(b + w), 0u);
            inter += __popc(ma & mb);
            uni += __popc(ma | mb);
        }
        inter = warp_sum(inter) >> 3;
        uni = warp_sum(uni) >> 3;
        if (lane == 0) {
            if (inter_out != nullptr) inter_out[f] = inter;
            if (union_out != nullptr) union_out[f] = uni;
        }
        const double iou = __ddiv_rn(static_cast<double>(inter), static_cast<double>(uni));   // 0/0 = NaN
        for (int k = lane; k < k_thr; k += 32)
            if (iou > __ldg(thr + k)) atomicAdd(&s_pos[k], 1u);
    }
    __syncthreads();
    for (int k = tid; k < k_thr; k += kIouThreads)
        if (s_pos[k] != 0) atomicAdd(pos + k, static_cast<unsigned long long>(s_pos[k]));
}

// Per-clip variant for video streams (BASELINE configs[3]: 10 s clips of 120 or 300 frames): frame f belongs to clip
// f / frames_per_clip and the success counts are kept per clip, pos[clip][k], so that every clip's success curve and AUC
// come out of ONE launch instead of one sweep per clip.  One warp per frame pair; the (clip, k) counters are global
// 64-bit atomics (at most n * K increments spread over n_clips * K addresses).
__global__ void __launch_bounds__(kIouThreads)
iou_sweep_clips_kernel(const uint8_t* __restrict__ mask_a, const uint8_t* __restrict__ mask_b, long long n,
                       long long frames_per_clip, const double* __restrict__ thr, int k_thr,
                       long long* __restrict__ inter_out, long long* __restrict__ union_out,
                       unsigned long long* __restrict__ pos /* [n_clips, k_thr] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long warps_total = static_cast<long long>(gridDim.x) * (kIouThreads / 32);
    for (long long f = static_cast<long long>(blockIdx.x) * (kIouThreads / 32) + warp; f < n; f += warps_total) {
        const uint32_t* a = reinterpret_cast<const uint32_t*>(mask_a + f * kFramePixels);
        const uint32_t* b = reinterpret_cast<const uint32_t*>(mask_b + f * kFramePixels);
        int inter = 0, uni = 0;
        for (int w = lane; w < kMaskWords; w += 32) {
            const uint32_t ma = __vcmpne4(__ldg(a + w), 0u);
            const uint32_t mb = __vcmpne4(__ldg(b + w), 0u);
            inter += __popc(ma & mb);
            uni += __popc(ma | mb);
        }
        inter = warp_sum(inter) >> 3;
        uni = warp_sum(uni) >> 3;
        if (lane == 0) {
            if (inter_out != nullptr) inter_out[f] = inter;
            if (union_out != nullptr) union_out[f] = uni;
        }
        const double iou = __ddiv_rn(static_cast<double>(inter), static_cast<double>(uni));
        unsigned long long* clip_pos = pos + (f / frames_per_clip) * k_thr;
        for (int k = lane; k < k_thr; k += 32)
            if (iou > __ldg(thr + k)) atomicAdd(clip_pos + k, 1ull);
    }
}

// Dynamic shared memory: MaskTaps::bytes(out_h, out_w) + out_w + out_h bytes (box membership bits per column / row).
__global__ void __launch_bounds__(kIouThreads)
ciou_sweep_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ xmin, const int* __restrict__ xmax,
                  const int* __restrict__ ymin, const int* __restrict__ ymax, long long n, int out_h, int out_w,
                  const double* __restrict__ thr, int k_thr, long long* __restrict__ inter2_out,
                  long long* __restrict__ union2_out, unsigned long long* __restrict__ pos,
                  unsigned long long* __restrict__ num) {
    extern __shared__ int s_taps[];
    __shared__ unsigned int s_pos[kMaxThresholds];
    __shared__ uint8_t s_mask[kFramePixels];
    __shared__ int s_box[3][4];                 // xa, xb, ya, yb (xa > xb: absent)
    __shared__ int s_sum[2][kIouThreads / 32];
    __shared__ double s_iou;
    MaskTaps t;
    t.carve(s_taps, out_h, out_w);
    uint8_t* s_colbits = reinterpret_cast<uint8_t*>(t.rows + kFrameH * out_w);   // [out_w] bit c: column inside box c
    uint8_t* s_rowbits = s_colbits + out_w;                                       // [out_h] bit c: row inside box c
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < k_thr; k += kIouThreads) s_pos[k] = 0;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(num, static_cast<unsigned long long>(n));   // num += 1 per frame (:321)
    t.build_taps(out_h, out_w, tid, kIouThreads);
    const int yd = 2 * out_h, half = 2 * out_w * out_h;
    for (long long f = blockIdx.x; f < n; f += gridDim.x) {
        __syncthreads();
        for (int p = tid; p < kFramePixels; p += kIouThreads) s_mask[p] = mask[f * kFramePixels + p] != 0;
        if (tid < 3) {
            // cv2.rectangle(..., thickness=-1): both corners inclusive, any corner order, clipped
            const int x_lo = xmin[f * 3 + tid], x_hi = xmax[f * 3 + tid];
            const int y_lo = ymin[f * 3 + tid], y_hi = ymax[f * 3 + tid];
            int xa = max(min(x_lo, x_hi), 0), xb = min(max(x_lo, x_hi), out_w - 1);
            int ya = max(min(y_lo, y_hi), 0), yb = min(max(y_lo, y_hi), out_h - 1);
            if (x_hi == 0 || ya > yb) { xa = 1; xb = 0; }     // `if xmax[h, contour] != 0` (:290)
            s_box[tid][0] = xa; s_box[tid][1] = xb; s_box[tid][2] = ya; s_box[tid][3] = yb;
        }
        __syncthreads();
        t.blend_rows(s_mask, out_w, warp, lane, kIouThreads / 32);
        for (int x = tid; x < out_w; x += kIouThreads) {
            int bits = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) bits |= (x >= s_box[c][0] && x <= s_box[c][1]) << c;
            s_colbits[x] = static_cast<uint8_t>(bits);
        }
        for (int y = tid; y < out_h; y += kIouThreads) {
            int bits = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) bits |= (y >= s_box[c][2] && y <= s_box[c][3]) << c;
            s_rowbits[y] = static_cast<uint8_t>(bits);
        }
        __syncthreads();
        int inter2 = 0, union2 = 0;
        for (int y = warp; y < out_h; y += kIouThreads / 32) {
            const int yi = t.y0[y], wn = t.yn[y], rowbits = s_rowbits[y];
            const uint16_t* r0 = t.rows + (yi & 0xffff) * out_w;
            const uint16_t* r1 = t.rows + (yi >> 16) * out_w;
            for (int x = lane; x < out_w; x += 32) {
                const int pred = r0[x] * (yd - wn) + r1[x] * wn > half;
                const int g2 = min(__popc(s_colbits[x] & rowbits), 2);   // half units; mtot[mtot > 1] = 1 (:296)
                inter2 += pred ? g2 : 0;                                  // (mtot and m2) * mtot (:306-308)
                union2 += g2 > 0 ? g2 : 2 * pred;                         // (mtot or m2) + (mtot - [mtot>0]) (:310-316)
            }
        }
        inter2 = warp_sum(inter2);
        union2 = warp_sum(union2);
        if (lane == 0) { s_sum[0][warp] = inter2; s_sum[1][warp] = union2; }
        __syncthreads();
        if (tid == 0) {
            int i2 = 0, u2 = 0;
#pragma unroll
            for (int w = 0; w < kIouThreads / 32; ++w) { i2 += s_sum[0][w]; u2 += s_sum[1][w]; }
            if (inter2_out != nullptr) inter2_out[f] = i2;
            if (union2_out != nullptr) union2_out[f] = u2;
            s_iou = __ddiv_rn(static_cast<double>(i2), static_cast<double>(u2));
        }
        __syncthreads();
        const double iou = s_iou;
        for (int k = tid; k < k_thr; k += kIouThreads)
            if (iou > __ldg(thr + k)) s_pos[k] += 1u;          // thread k owns s_pos[k]: no atomics
    }
    __syncthreads();
    for (int k = tid; k < k_thr; k += kIouThreads)
        if (s_pos[k] != 0) atomicAdd(pos + k, static_cast<unsigned long long>(s_pos[k]));
}

}  // namespace aig
