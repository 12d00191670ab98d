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
//   * the vertical taps repeat every 56 output rows = 9 source rows (224 / 36 = 56 / 9), i.e. every 7 chunks of 8 rows:
//     warp w of 7 takes chunks w, w + 7, w + 14, w + 21, whose source rows and weights are compile-time constants of w
//     (a 7-way switch on the warp index): weights are immediates, shared-memory offsets are immediates, and which rows
//     reuse the previous row's loads is decided by the compiler.  The clamped taps of the first and last rows fit the
//     same pattern because the blended rows carry a copy of the first row above them and of the last row below;
//   * resize: the four flags become four bytes with one PRMT, one shift and one AND, rows are staged in shared memory in
//     chunks of 8 rows (a multiple of 16 bytes at both sizes) and leave the SM as bulk asynchronous copies;
//   * consensus IoU: nothing is stored.  With at most three boxes, min(count, 2) = count - [count == 3], so
//         2 I = P(B0) + P(B1) + P(B2) - P(B0 n B1 n B2)
//         2 U = area-weighted box coverage (closed form) + 2 [P(all) - P(B0 u B1 u B2)]
//     where P(R) counts the predicted pixels inside rectangle R: eight rectangle counts per frame.  The flags of 8 rows x
//     4 columns are collected in one register (one shift and one LOP3 per row) and counted against a rectangle with one
//     LOP3 and one POPC per 32 pixels.
// Both are bit-exact replacements (tests compare them with the generic kernels and the oracle; tests/test_mask_packed_cpu.py
// restates this arithmetic in NumPy against the oracle).  Measured on 8192 frames (tools/mask_probe.py): resize 54.6 M frames/s
// at 224 x 298 = 3.6 TB/s written and 82.9 M at 224 x 224 (generic: 15.9 M / 22.8 M); consensus IoU with 101 thresholds
// 43.2 M / 59.6 M (generic: 13.8 M / 20.1 M); 13.9 k / 17.9 k warp instructions per frame, issue slots 73 % busy.
#pragma once

#include <type_traits>

#include "aig_common.cuh"
#include "heatmap_kernel.cuh"   // linear_tap_exact

namespace aig {

__host__ __device__ constexpr int packed_gcd(int a, int b) { return b == 0 ? (a < 0 ? -a : a) : packed_gcd(b, a % b); }
__host__ __device__ constexpr int packed_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

constexpr int kPackedWarps = 7;                    // H = 224: 28 chunks of 8 rows, four per warp
constexpr int kPackedThreads = kPackedWarps * 32;

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
    static constexpr int PERIOD_OUT = H / packed_gcd(H, kFrameH);  // the vertical taps repeat every PERIOD_OUT output rows ...
    static constexpr int PERIOD_SRC = kFrameH / packed_gcd(H, kFrameH);   // ... = PERIOD_SRC source rows
    static constexpr int PERIODS = H / PERIOD_OUT;
    // output row y (unclamped): between source rows lo and lo + 1 with weight n / YD on the latter.  Rows above the first
    // source row (lo = -1) and below the last fit in because the blended rows are stored with one copy of the edge row
    // on either side (index lo + 1 into PackedTaps::rows).
    static __host__ __device__ constexpr int tap_lo(int y) { return packed_floor_div((2 * y + 1) * kFrameH - H, 2 * H); }
    static __host__ __device__ constexpr int tap_n(int y) { return ((2 * y + 1) * kFrameH - H - tap_lo(y) * 2 * H) / FY; }
    static_assert((XD * YD) % 2 == 0 && XD * YD <= 65535, "two pixels per 32-bit multiply-add need 16-bit values");
    static_assert(W % 2 == 0 && (8 * W) % 16 == 0 && H % 8 == 0, "8-row chunks must be 16-byte multiples");
    static_assert(PERIOD_OUT == 8 * kPackedWarps, "one chunk of every period per warp");
    static_assert(tap_lo(0) == -1 && tap_lo(H - 1) == kFrameH - 1, "one padding row on either side is enough");
};

template <int W, int H>
struct PackedTaps {
    using G = PackedGeom<W, H>;
    alignas(16) uint16_t rows[kFrameH + 2][G::WP]; // horizontally blended source rows (values 0 .. XD); [0] = [1], [37] = [36]
    int x01[W];                                    // i0 | i1 << 16
    int xn[W];                                     // reduced weight of column i1
    uint8_t mask[kFramePixels];

    __device__ __forceinline__ void build(int tid) {
        for (int d = tid; d < W; d += kPackedThreads) {
            int i0, i1, r; linear_tap_exact(d, kFrameW, W, &i0, &i1, &r);
            x01[d] = i0 | (i1 << 16); xn[d] = r / G::FX;
        }
        if constexpr (G::WP > W) {                   // padding columns: read by the last lane's 64-bit loads, never used
            constexpr int kPad = G::WP - W;
            for (int i = tid; i < (kFrameH + 2) * kPad; i += kPackedThreads) rows[i / kPad][W + i % kPad] = 0;
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
            uint16_t first = 0, last = 0;
#pragma unroll
            for (int r = 0; r < kRowsPer; ++r) {
                const uint8_t* m = mask + (r0 + r) * kFrameW;
                const uint16_t v = static_cast<uint16_t>(m[c0] * (G::XD - n) + m[c1] * n);
                rows[r0 + r + 1][x] = v;
                if (r == 0) first = v;
                if (r == kRowsPer - 1) last = v;
            }
            if (r0 == 0) rows[0][x] = first;                                   // the copies above the first and below the last row
            if (r0 + kRowsPer == kFrameH) rows[kFrameH + 1][x] = last;
        }
    }
    // flags of the four pixels (x .. x + 3) of one output row: bit 7 of byte j set iff pixel x + j is above one half.
    // a, b: the lane's four blended values of the row's two source rows, n the weight of b.
    static __device__ __forceinline__ unsigned int flags4(const uint2& a, const uint2& b, unsigned int n) {
        const unsigned int w0 = G::YD - n;
        unsigned int lo = a.x * w0 + G::KK; lo = b.x * n + lo;
        unsigned int hi = a.y * w0 + G::KK; hi = b.y * n + hi;
        return __byte_perm(lo, hi, 0x7531);
    }
    // The 8 rows of chunk PHASE + 7 * period for columns x .. x + 3: emit(r, flags) for r = 0 .. 7.  Everything that depends
    // on the row - its two source rows, its weight, whether the previous row's loads still serve - is a constant.
    template <int PHASE, typename Emit>
    __device__ __forceinline__ void chunk(int period, int x, Emit emit) const {
        const uint16_t* base = &rows[period * G::PERIOD_SRC][x];
        uint2 a = make_uint2(0u, 0u), b = a;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            constexpr int y0 = 8 * PHASE;
            const int p = G::tap_lo(y0 + r) + 1, prev = r ? G::tap_lo(y0 + r - 1) + 1 : -5;
            if (p == prev + 1) {
                a = b;
                b = *reinterpret_cast<const uint2*>(base + (p + 1) * G::WP);
            } else if (p != prev) {
                a = *reinterpret_cast<const uint2*>(base + p * G::WP);
                b = *reinterpret_cast<const uint2*>(base + (p + 1) * G::WP);
            }
            emit(r, flags4(a, b, static_cast<unsigned int>(G::tap_n(y0 + r))));
        }
    }
};

// f(std::integral_constant<int, warp>) for the calling warp: every warp runs code specialised for its chunks
template <typename F>
__device__ __forceinline__ void packed_dispatch(int warp, F f) {
    switch (warp) {
        case 0: f(std::integral_constant<int, 0>{}); break;
        case 1: f(std::integral_constant<int, 1>{}); break;
        case 2: f(std::integral_constant<int, 2>{}); break;
        case 3: f(std::integral_constant<int, 3>{}); break;
        case 4: f(std::integral_constant<int, 4>{}); break;
        case 5: f(std::integral_constant<int, 5>{}); break;
        default: f(std::integral_constant<int, 6>{}); break;
    }
}
static_assert(kPackedWarps == 7, "packed_dispatch lists the warps");

// mask [n, 36, 48] u8 -> mask_up [n, H, W] u8 (1 iff bilinear(mask != 0) > 1/2).  mask_up must be 16-byte aligned.
template <int W, int H>
struct ResizePackedSmem {
    PackedTaps<W, H> t;
    alignas(16) uint8_t stage[kPackedWarps][8 * W];   // one slot per warp: 44 KB in all at 224 x 298, five CTAs per SM
};

template <int W, int H>
__global__ void __launch_bounds__(kPackedThreads)
resize_mask_packed_kernel(const uint8_t* __restrict__ mask, long long n_frames, uint8_t* __restrict__ mask_up) {
    using G = PackedGeom<W, H>;
    extern __shared__ __align__(16) unsigned char s_packed_raw[];
    ResizePackedSmem<W, H>& s = *reinterpret_cast<ResizePackedSmem<W, H>*>(s_packed_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    s.t.build(tid);
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {
        __syncthreads();                                   // the previous frame's rows have been read
        s.t.load_mask(mask + frame * kFramePixels, tid);
        __syncthreads();
        s.t.blend(warp, lane);
        __syncthreads();
        uint8_t* dst = mask_up + frame * static_cast<long long>(H) * W;
        packed_dispatch(warp, [&](auto phase) {
            constexpr int PHASE = decltype(phase)::value;
            for (int period = 0; period < G::PERIODS; ++period) {
                if (lane == 0) bulk_wait_read<0>();        // the previous chunk's copy has finished reading the slot
                __syncwarp();
                uint8_t* st = s.stage[PHASE];
#pragma unroll 1                                           // (seven warps run seven different bodies: unrolled further, instruction fetch stalls them)
                for (int k = 0; k < G::STEPS; ++k) {
                    const int x = 4 * (lane + 32 * k);
                    if (x >= W) break;
                    s.t.template chunk<PHASE>(period, x, [&](int r, unsigned int flags) {
                        const unsigned int bytes = (flags >> 7) & 0x01010101u;
                        if (W % 4 == 0) {
                            *reinterpret_cast<unsigned int*>(st + r * W + x) = bytes;
                        } else {                           // rows start on 2-byte boundaries only
                            *reinterpret_cast<unsigned short*>(st + r * W + x) = static_cast<unsigned short>(bytes);
                            if (x + 2 < W) *reinterpret_cast<unsigned short*>(st + r * W + x + 2) = static_cast<unsigned short>(bytes >> 16);
                        }
                    });
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    bulk_store_s2g(dst + static_cast<long long>(PHASE + kPackedWarps * period) * 8 * W, smem_u32(st), 8 * W);
                    bulk_commit();
                }
            }
        });
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
    unsigned int rowmask[kPackedRects][PackedGeom<W, H>::CHUNKS];   // 0x01010101 * (bit r: row 8 c + r inside the rectangle)
    unsigned int active;                           // bit R: rectangle R is not empty
    int count[kPackedWarps][kPackedRects];
    unsigned int pos[kPackedMaxThresholds];
    double iou;
};

template <int W, int H>
__global__ void __launch_bounds__(kPackedThreads, 5)
ciou_packed_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ xmin, const int* __restrict__ xmax,
                   const int* __restrict__ ymin, const int* __restrict__ ymax, long long n, const double* __restrict__ thr,
                   int k_thr, long long* __restrict__ inter2_out, long long* __restrict__ union2_out,
                   unsigned long long* __restrict__ pos, unsigned long long* __restrict__ num) {
    using G = PackedGeom<W, H>;
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
            const int r = i / G::CHUNKS, c = i % G::CHUNKS;
            const int ya = s.rect[r][2] - 8 * c, yb = s.rect[r][3] - 8 * c;       // rows of this chunk inside: [max(ya, 0), min(yb, 7)]
            unsigned int bits = 0;
            if (s.rect[r][0] <= s.rect[r][1] && yb >= 0 && ya <= 7)
                bits = (0xffu >> (7 - min(yb, 7))) & (0xffu << max(ya, 0));
            s.rowmask[r][c] = bits * 0x01010101u;
        }
        s.t.blend(warp, lane);
        __syncthreads();
        // Per column step k and chunk: bit r of byte j of `flags` says pixel (8 chunk + r, 4 (lane + 32 k) + j) is predicted; a
        // rectangle's count grows by the bits inside its columns (colsel) and rows (rowmask).  Both loops stay rolled: seven
        // warps run seven different bodies, and unrolled further (4096 instructions) instruction fetch stalled them
        // (ncu: 7.5 warps waiting for instructions per instruction issued).
        const unsigned int active = s.active;
        int cnt[kPackedRects];
#pragma unroll
        for (int r = 0; r < kPackedRects; ++r) cnt[r] = 0;
#pragma unroll 1
        for (int k = 0; k < G::STEPS; ++k) {
            const unsigned int x = 4 * (lane + 32 * k);
            if (x >= W) break;
            unsigned int colsel[kPackedRects];              // 0xff in byte j iff column x + j lies inside the rectangle
#pragma unroll
            for (int r = 0; r < kPackedRects; ++r) {
                colsel[r] = 0;
                if (active & (1u << r)) {                   // warp-uniform
                    const unsigned int xa = s.rect[r][0], span = s.rect[r][1] - s.rect[r][0];
#pragma unroll
                    for (int j = 0; j < 4; ++j) colsel[r] |= (x + j - xa <= span) ? (0xffu << (8 * j)) : 0u;
                }
            }
            packed_dispatch(warp, [&](auto phase) {
                constexpr int PHASE = decltype(phase)::value;
#pragma unroll 1
                for (int period = 0; period < G::PERIODS; ++period) {
                    unsigned int flags = 0;
                    s.t.template chunk<PHASE>(period, x, [&](int, unsigned int fl) { flags = (flags >> 1) | (fl & 0x80808080u); });
                    const unsigned int* rowmask = &s.rowmask[0][PHASE + kPackedWarps * period];
#pragma unroll
                    for (int r = 0; r < kPackedRects; ++r)
                        if (active & (1u << r)) cnt[r] += __popc(flags & colsel[r] & rowmask[r * G::CHUNKS]);
                }
            });
        }
#pragma unroll
        for (int r = 0; r < kPackedRects; ++r) {
            int v = 0;
            if (active & (1u << r)) v = warp_sum(cnt[r]);
            if (lane == 0) s.count[warp][r] = v;
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
