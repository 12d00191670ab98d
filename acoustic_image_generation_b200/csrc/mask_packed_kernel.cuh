// The exact thresholded bilinear up-sampling of a 36 x 48 mask (showimages_bb.py:303-304: cv2.resize(m2 * 1.0, (298, 224)) > 0.5)
// for the reference's output sizes as compile-time constants, two pixels per 32-bit multiply-add.
//
// resize_mask_packed_kernel<W, H>   aig_resize_mask at 224 x 298 and 224 x 224 (resize_mask_kernel serves every other size)
// ciou_packed_kernel<W, H>          aig_ciou_sweep  at the same sizes          (ciou_sweep_kernel serves every other size)
//
// resize_mask_kernel / ciou_sweep_kernel spend 24-30 warp instructions per 32 output pixels (ncu, round 2: issue slots 74-77 %
// busy, DRAM 7 %): one pixel per lane, 32-bit integer arithmetic, run-time loop bounds, a byte store or three table look-ups
// per pixel.  Here:
//   * every tap numerator along x shares a factor FX with its denominator 2 W, along y a factor FY with 2 H (the numerators
//     are 96 d + 48 - W and 72 d + 36 - H): at 298 x 224 the blended rows are <= 298 and the vertical weights <= 112, so
//     a pixel's value r0 (YD - n) + r1 n <= 33 376 fits in 16 bits and TWO pixels share one 32-bit IMAD per source row;
//     adding K = 32767 - T first puts "value > T" (T = half of full scale, exact) into bit 15 of each half;
//   * a lane owns four adjacent columns: one 64-bit shared load per source row, reused by the output rows that
//     interpolate between the same two source rows (five of six);
//   * resize: the four flags become four bytes with one PRMT, one shift and one AND, rows are staged in shared memory in
//     chunks of 8 rows (a multiple of 16 bytes at both sizes) and leave the SM as bulk asynchronous copies;
//   * consensus IoU: nothing is stored.  With at most three boxes, min(count, 2) = count - [count == 3], so
//         2 I = P(B0) + P(B1) + P(B2) - P(B0 n B1 n B2)
//         2 U = area-weighted box coverage (closed form) + 2 [P(all) - P(B0 u B1 u B2)]
//     where P(R) counts the predicted pixels inside rectangle R: eight rectangle counts per frame.  The flags of 8 rows x
//     4 columns are collected in one register (one shift and one LOP3 per row) and counted against a rectangle with one
//     LOP3 and one POPC per 32 pixels.
// Both are bit-exact replacements (tests compare them with the generic kernels and the oracle).
#pragma once

#include "aig_common.cuh"
#include "heatmap_kernel.cuh"   // linear_tap_exact

namespace aig {

constexpr int packed_gcd(int a, int b) { return b == 0 ? (a < 0 ? -a : a) : packed_gcd(b, a % b); }

template <int W, int H>
struct PackedGeom {
    static constexpr int FX = packed_gcd(packed_gcd(2 * kFrameW, 2 * W), kFrameW - W);
    static constexpr int FY = packed_gcd(packed_gcd(2 * kFrameH, 2 * H), kFrameH - H);
    static constexpr int XD = 2 * W / FX, YD = 2 * H / FY;        // reduced denominators
    static constexpr int T = XD * YD / 2;                          // value > T  <=>  bilinear > 1/2
    static constexpr unsigned int K = 32767u - T;                  // value + K has bit 15 set iff value > T
    static constexpr unsigned int KK = K | (K << 16);
    static constexpr int WP = (W + 3) & ~3;                        // row stride of the blended rows (uint16)
    static constexpr int STEPS = (W / 4 + (W % 4 ? 1 : 0) + 31) / 32;   // column steps of 32 lanes x 4 columns
    static constexpr int CHUNKS = H / 8;
    static_assert((XD * YD) % 2 == 0 && XD * YD <= 65535, "two pixels per 32-bit multiply-add need 16-bit values");
    static_assert(W % 2 == 0 && (8 * W) % 16 == 0 && H % 8 == 0, "8-row chunks must be 16-byte multiples");
};

constexpr int kPackedWarps = 7;                    // H = 224: 28 chunks of 8 rows, four per warp
constexpr int kPackedThreads = kPackedWarps * 32;

template <int W, int H>
struct PackedTaps {
    using G = PackedGeom<W, H>;
    alignas(16) uint16_t rows[kFrameH][G::WP];     // horizontally blended source rows, values 0 .. XD
    alignas(16) int ytap[H];                       // i0 | i1 << 8 | n << 16 (n = reduced weight of row i1)
    int x01[W];                                    // i0 | i1 << 16
    int xn[W];                                     // reduced weight of column i1
    uint8_t mask[kFramePixels];

    __device__ __forceinline__ void build(int tid) {
        for (int d = tid; d < W; d += kPackedThreads) {
            int i0, i1, r; linear_tap_exact(d, kFrameW, W, &i0, &i1, &r);
            x01[d] = i0 | (i1 << 16); xn[d] = r / G::FX;
        }
        for (int d = tid; d < H; d += kPackedThreads) {
            int i0, i1, r; linear_tap_exact(d, kFrameH, H, &i0, &i1, &r);
            ytap[d] = i0 | (i1 << 8) | ((r / G::FY) << 16);
        }
        if constexpr (G::WP > W) {                   // padding columns: read by the last lane's 64-bit loads, never selected
            constexpr int kPad = G::WP - W;
            for (int i = tid; i < kFrameH * kPad; i += kPackedThreads) rows[i / kPad][W + i % kPad] = 0;
        }
    }
    __device__ __forceinline__ void load_mask(const uint8_t* src, int tid) {
        for (int p = tid; p < kFramePixels; p += kPackedThreads) mask[p] = src[p] != 0;
    }
    // a lane takes one output column and nine source rows at a time: the column's taps are loaded once
    __device__ __forceinline__ void blend(int warp, int lane) {
        constexpr int kColSteps = (W + 31) / 32, kRowGroups = 4, kRowsPer = kFrameH / kRowGroups;
        for (int u = warp; u < kColSteps * kRowGroups; u += kPackedWarps) {
            const int x = (u % kColSteps) * 32 + lane, r0 = (u / kColSteps) * kRowsPer;
            if (x >= W) continue;
            const int xi = x01[x], n = xn[x];
            const int c0 = xi & 0xffff, c1 = xi >> 16;
#pragma unroll
            for (int r = 0; r < kRowsPer; ++r) {
                const uint8_t* m = mask + (r0 + r) * kFrameW;
                rows[r0 + r][x] = static_cast<uint16_t>(m[c0] * (G::XD - n) + m[c1] * n);
            }
        }
    }
    // flags of the four pixels (x .. x + 3) of one output row: bit 7 of byte j set iff pixel x + j is above one half.
    // a, b: the lane's four blended values of the row's two source rows.
    static __device__ __forceinline__ unsigned int flags4(const uint2& a, const uint2& b, unsigned int n) {
        const unsigned int w0 = G::YD - n;
        unsigned int lo = a.x * w0 + G::KK; lo = b.x * n + lo;
        unsigned int hi = a.y * w0 + G::KK; hi = b.y * n + hi;
        return __byte_perm(lo, hi, 0x7531);
    }
};

// mask [n, 36, 48] u8 -> mask_up [n, H, W] u8 (1 iff bilinear(mask != 0) > 1/2).  mask_up must be 16-byte aligned.
template <int W, int H>
struct ResizePackedSmem {
    PackedTaps<W, H> t;
    alignas(16) uint8_t stage[kPackedWarps][2][8 * W];
};

template <int W, int H>
__global__ void __launch_bounds__(kPackedThreads)
resize_mask_packed_kernel(const uint8_t* __restrict__ mask, long long n_frames, uint8_t* __restrict__ mask_up) {
    using G = PackedGeom<W, H>;
    using T = PackedTaps<W, H>;
    extern __shared__ __align__(16) unsigned char s_packed_raw[];
    ResizePackedSmem<W, H>& s = *reinterpret_cast<ResizePackedSmem<W, H>*>(s_packed_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    s.t.build(tid);
    unsigned int chunk_it = 0;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        __syncthreads();                                   // the previous frame's rows have been read
        s.t.load_mask(mask + frame * kFramePixels, tid);
        __syncthreads();
        s.t.blend(warp, lane);
        __syncthreads();
        uint8_t* dst = mask_up + frame * static_cast<long long>(H) * W;
        for (int c = warp; c < G::CHUNKS; c += kPackedWarps, ++chunk_it) {
            const unsigned int buf = chunk_it & 1u;
            if (chunk_it >= 2u) {                          // the copy issued two chunks ago has finished reading this buffer
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
            }
            uint8_t* st = s.stage[warp][buf];
            int yt[8];
            {
                const int4 t0 = *reinterpret_cast<const int4*>(&s.t.ytap[8 * c]), t1 = *reinterpret_cast<const int4*>(&s.t.ytap[8 * c + 4]);
                yt[0] = t0.x; yt[1] = t0.y; yt[2] = t0.z; yt[3] = t0.w; yt[4] = t1.x; yt[5] = t1.y; yt[6] = t1.z; yt[7] = t1.w;
            }
#pragma unroll
            for (int k = 0; k < G::STEPS; ++k) {
                const int x = 4 * (lane + 32 * k);
                if (x >= W) break;
                uint2 a = make_uint2(0u, 0u), b = a;
                int cur = -1;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int pair = yt[r] & 0xffff;
                    if (pair != cur) {                     // warp-uniform: five rows of six keep their source rows
                        cur = pair;
                        a = *reinterpret_cast<const uint2*>(&s.t.rows[pair & 0xff][x]);
                        b = *reinterpret_cast<const uint2*>(&s.t.rows[pair >> 8][x]);
                    }
                    const unsigned int bytes = (T::flags4(a, b, static_cast<unsigned int>(yt[r]) >> 16) >> 7) & 0x01010101u;
                    if (W % 4 == 0) {
                        *reinterpret_cast<unsigned int*>(st + r * W + x) = bytes;
                    } else {                               // rows start on 2-byte boundaries only
                        *reinterpret_cast<unsigned short*>(st + r * W + x) = static_cast<unsigned short>(bytes);
                        if (x + 2 < W) *reinterpret_cast<unsigned short*>(st + r * W + x + 2) = static_cast<unsigned short>(bytes >> 16);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store_s2g(dst + static_cast<long long>(c) * 8 * W, smem_u32(st), 8 * W);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_all<0>();                     // shared memory must outlive the copies that read it
}

// Consensus IoU (showimages_bb.py:288-321) at a compile-time output size.  Rectangles: 0 = the image, 1-3 = the boxes,
// 4-6 = their pairwise intersections (01, 02, 12), 7 = the triple intersection.
constexpr int kPackedRects = 8;
constexpr int kPackedMaxThresholds = 1024;         // == kMaxThresholds of score_kernel.cuh

template <int W, int H>
struct CiouPackedSmem {
    PackedTaps<W, H> t;
    int rect[kPackedRects][4];                     // xa, xb, ya, yb (inclusive; xa > xb: empty)
    unsigned int rowmask[PackedGeom<W, H>::CHUNKS][kPackedRects];   // 0x01010101 * (bit r: row 8 c + r inside the rectangle)
    unsigned int active;                           // bit R: rectangle R is not empty
    int count[kPackedWarps][kPackedRects];
    unsigned int pos[kPackedMaxThresholds];
    double iou;
};

template <int W, int H>
__global__ void __launch_bounds__(kPackedThreads)
ciou_packed_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ xmin, const int* __restrict__ xmax,
                   const int* __restrict__ ymin, const int* __restrict__ ymax, long long n, const double* __restrict__ thr,
                   int k_thr, long long* __restrict__ inter2_out, long long* __restrict__ union2_out,
                   unsigned long long* __restrict__ pos, unsigned long long* __restrict__ num) {
    using G = PackedGeom<W, H>;
    using T = PackedTaps<W, H>;
    extern __shared__ __align__(16) unsigned char s_packed_raw[];
    CiouPackedSmem<W, H>& s = *reinterpret_cast<CiouPackedSmem<W, H>*>(s_packed_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < k_thr; k += kPackedThreads) s.pos[k] = 0;
    if (blockIdx.x == 0 && tid == 0) atomicAdd(num, static_cast<unsigned long long>(n));   // num += 1 per frame (:321)
    s.t.build(tid);
    for (long long f = blockIdx.x; f < n; f += gridDim.x) {
        __syncthreads();                                   // the previous frame's rows, rectangles and counts have been read
        s.t.load_mask(mask + f * kFramePixels, tid);
        if (tid == 0) {
            // cv2.rectangle(..., thickness=-1): both corners inclusive, any corner order, clipped
            int bx[3][4];
            for (int c = 0; c < 3; ++c) {
                const int x_lo = xmin[f * 3 + c], x_hi = xmax[f * 3 + c];
                const int y_lo = ymin[f * 3 + c], y_hi = ymax[f * 3 + c];
                int xa = max(min(x_lo, x_hi), 0), xb = min(max(x_lo, x_hi), W - 1);
                int ya = max(min(y_lo, y_hi), 0), yb = min(max(y_lo, y_hi), H - 1);
                if (x_hi == 0 || ya > yb || xa > xb) { xa = 1; xb = 0; ya = 1; yb = 0; }     // `if xmax[h, contour] != 0` (:290)
                bx[c][0] = xa; bx[c][1] = xb; bx[c][2] = ya; bx[c][3] = yb;
            }
            auto put = [&](int r, int xa, int xb, int ya, int yb) {
                if (xa > xb || ya > yb) { xa = 1; xb = 0; ya = 1; yb = 0; }
                s.rect[r][0] = xa; s.rect[r][1] = xb; s.rect[r][2] = ya; s.rect[r][3] = yb;
                return (xa <= xb) ? (1u << r) : 0u;
            };
            auto meet = [&](int r, const int* p, const int* q) {
                const bool both = p[0] <= p[1] && q[0] <= q[1];
                return both ? put(r, max(p[0], q[0]), min(p[1], q[1]), max(p[2], q[2]), min(p[3], q[3])) : put(r, 1, 0, 1, 0);
            };
            unsigned int act = put(0, 0, W - 1, 0, H - 1);
            for (int c = 0; c < 3; ++c) act |= put(1 + c, bx[c][0], bx[c][1], bx[c][2], bx[c][3]);
            act |= meet(4, bx[0], bx[1]);
            act |= meet(5, bx[0], bx[2]);
            act |= meet(6, bx[1], bx[2]);
            act |= meet(7, s.rect[4], bx[2]);
            s.active = act;
        }
        __syncthreads();
        for (int i = tid; i < G::CHUNKS * kPackedRects; i += kPackedThreads) {
            const int c = i / kPackedRects, r = i % kPackedRects;
            const int ya = s.rect[r][2] - 8 * c, yb = s.rect[r][3] - 8 * c;       // rows of this chunk inside: [max(ya, 0), min(yb, 7)]
            unsigned int bits = 0;
            if (s.rect[r][0] <= s.rect[r][1] && yb >= 0 && ya <= 7)
                bits = (0xffu >> (7 - min(yb, 7))) & (0xffu << max(ya, 0));
            s.rowmask[c][r] = bits * 0x01010101u;
        }
        s.t.blend(warp, lane);
        __syncthreads();
        const unsigned int active = s.active;
        int cnt[kPackedRects];
#pragma unroll
        for (int r = 0; r < kPackedRects; ++r) cnt[r] = 0;
#pragma unroll
        for (int k = 0; k < G::STEPS; ++k) {
            const int x = 4 * (lane + 32 * k);
            if (x >= W) break;
            unsigned int colsel[kPackedRects];             // 0xff in byte j iff column x + j lies inside the rectangle
#pragma unroll
            for (int r = 0; r < kPackedRects; ++r) {
                const int xa = s.rect[r][0], xb = s.rect[r][1];
                unsigned int sel = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) sel |= (x + j >= xa && x + j <= xb) ? (0xffu << (8 * j)) : 0u;
                colsel[r] = sel;
            }
            for (int c = warp; c < G::CHUNKS; c += kPackedWarps) {
                int yt[8];
                {
                    const int4 t0 = *reinterpret_cast<const int4*>(&s.t.ytap[8 * c]), t1 = *reinterpret_cast<const int4*>(&s.t.ytap[8 * c + 4]);
                    yt[0] = t0.x; yt[1] = t0.y; yt[2] = t0.z; yt[3] = t0.w; yt[4] = t1.x; yt[5] = t1.y; yt[6] = t1.z; yt[7] = t1.w;
                }
                uint2 a = make_uint2(0u, 0u), b = a;
                int cur = -1;
                unsigned int flags = 0;                     // bit r of byte j: pixel (8 c + r, x + j) is predicted
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int pair = yt[r] & 0xffff;
                    if (pair != cur) {
                        cur = pair;
                        a = *reinterpret_cast<const uint2*>(&s.t.rows[pair & 0xff][x]);
                        b = *reinterpret_cast<const uint2*>(&s.t.rows[pair >> 8][x]);
                    }
                    flags = (flags >> 1) | (T::flags4(a, b, static_cast<unsigned int>(yt[r]) >> 16) & 0x80808080u);
                }
#pragma unroll
                for (int r = 0; r < kPackedRects; ++r)
                    if (active & (1u << r)) cnt[r] += __popc(flags & colsel[r] & s.rowmask[c][r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kPackedRects; ++r) {
            if (active & (1u << r)) {                       // warp-uniform
                const int v = warp_sum(cnt[r]);
                if (lane == 0) s.count[warp][r] = v;
            } else if (lane == 0) {
                s.count[warp][r] = 0;
            }
        }
        __syncthreads();
        if (tid == 0) {
            long long p[kPackedRects], area[kPackedRects];
            for (int r = 0; r < kPackedRects; ++r) {
                p[r] = 0;
                for (int w = 0; w < kPackedWarps; ++w) p[r] += s.count[w][r];
                const int* q = s.rect[r];
                area[r] = q[0] <= q[1] ? static_cast<long long>(q[1] - q[0] + 1) * (q[3] - q[2] + 1) : 0;
            }
            // half units: a pixel covered by c boxes weighs g2 = min(c, 2) = c - [c == 3]
            const long long i2 = p[1] + p[2] + p[3] - p[7];                                  // sum over predicted pixels of g2 (:306-308)
            const long long covered = p[1] + p[2] + p[3] - p[4] - p[5] - p[6] + p[7];        // predicted pixels inside any box
            const long long u2 = (area[1] + area[2] + area[3] - area[7]) + 2 * (p[0] - covered);   // (:310-316)
            if (inter2_out != nullptr) inter2_out[f] = i2;
            if (union2_out != nullptr) union2_out[f] = u2;
            s.iou = __ddiv_rn(static_cast<double>(i2), static_cast<double>(u2));
        }
        __syncthreads();
        const double iou = s.iou;
        for (int k = tid; k < k_thr; k += kPackedThreads)
            if (iou > __ldg(thr + k)) s.pos[k] += 1u;          // thread k % 224 owns s.pos[k]: no atomics
    }
    __syncthreads();
    for (int k = tid; k < k_thr; k += kPackedThreads)
        if (s.pos[k] != 0) atomicAdd(pos + k, static_cast<unsigned long long>(s.pos[k]));
}

}  // namespace aig
