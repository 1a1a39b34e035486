// color.cuh -- the still-image colour path: BGR -> YUV (or YCrCb), equalize / CLAHE on channel 0, -> BGR.
//
// Replaces cvtColor(BGR2YUV) -> split -> equalizeHist / clahe->apply -> merge -> cvtColor(YUV2BGR) of
// singlecolor.cpp:39-66 and clahe1frame.cpp:83-102 (8-bit Q14 fixed point, descale = (x+8192)>>14).
// The intermediate 3-channel YUV image, the split planes and the merged image are never materialised: pass 1 writes
// only the Y plane (W*H bytes, consumed by the equalize / CLAHE kernel), pass 2 re-derives U and V from the BGR
// input and combines them with the equalized Y.
#pragma once
#include "common.cuh"
#include "equalize.cuh"

namespace nv12eq {

constexpr int kColorThreads = 256;
enum : int { COLOR_YUV = 0, COLOR_YCRCB = 1 };

__device__ __forceinline__ int descale14(int v) { return (v + 8192) >> 14; }
__device__ __forceinline__ int sat8i(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ int bgr_luma(int B, int G, int R) { return descale14(1868 * B + 9617 * G + 4899 * R); }

// Forward chroma (saturated to 8 bits exactly as the intermediate 8UC3 image would hold it), then inverse with the
// new luma y2.  Returns packed B | G<<8 | R<<16.
__device__ __forceinline__ uint32_t bgr_recombine(int B, int G, int R, int y2, int mode) {
    const int Y = bgr_luma(B, G, R);
    int b2, g2, r2;
    if (mode == COLOR_YUV) {
        const int U = sat8i(descale14((B - Y) * 8061 + (128 << 14))) - 128;
        const int V = sat8i(descale14((R - Y) * 14369 + (128 << 14))) - 128;
        b2 = y2 + descale14(U * 33292);
        g2 = y2 + descale14(U * -6472 + V * -9519);
        r2 = y2 + descale14(V * 18678);
    } else {
        const int Cr = sat8i(descale14((R - Y) * 11682 + (128 << 14))) - 128;
        const int Cb = sat8i(descale14((B - Y) * 9241 + (128 << 14))) - 128;
        b2 = y2 + descale14(Cb * 29049);
        g2 = y2 + descale14(Cb * -5636 + Cr * -11698);
        r2 = y2 + descale14(Cr * 22987);
    }
    return (uint32_t)sat8i(b2) | ((uint32_t)sat8i(g2) << 8) | ((uint32_t)sat8i(r2) << 16);
}

struct ColorParams {
    const uint8_t* bgr_in;
    uint8_t* bgr_out;
    unsigned long long bgr_pitch;  // bytes between BGR frames
    int n_frames;
    int w, h, stride;              // stride in bytes of a BGR row
    uint8_t* y_plane;              // [n_frames][h][w] tightly packed
    const uint8_t* y2_plane;       // equalized luma, same layout
    int mode;
};

// 4 pixels per thread: 12 BGR bytes (three 32-bit words) -> 4 luma bytes (one word).
__global__ void __launch_bounds__(kColorThreads) bgr_to_luma_kernel(const ColorParams p) {
    const int f = blockIdx.y;
    const uint8_t* src = p.bgr_in + (unsigned long long)f * p.bgr_pitch;
    uint8_t* dst = p.y_plane + (size_t)f * p.w * p.h;
    const bool vec = (p.stride == 3 * p.w) && ((((uintptr_t)src | (uintptr_t)dst) & 3) == 0);
    const long long npx = (long long)p.w * p.h;
    if (vec) {
        const long long nquad = npx >> 2;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
        for (long long q = (long long)blockIdx.x * kColorThreads + threadIdx.x; q < nquad; q += (long long)gridDim.x * kColorThreads) {
            const uint32_t a = __ldg(s32 + 3 * q), b = __ldg(s32 + 3 * q + 1), c = __ldg(s32 + 3 * q + 2);
            const int y0 = bgr_luma(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            const int y1 = bgr_luma(a >> 24, b & 255, (b >> 8) & 255);
            const int y2 = bgr_luma((b >> 16) & 255, b >> 24, c & 255);
            const int y3 = bgr_luma((c >> 8) & 255, (c >> 16) & 255, c >> 24);
            d32[q] = (uint32_t)y0 | ((uint32_t)y1 << 8) | ((uint32_t)y2 << 16) | ((uint32_t)y3 << 24);
        }
        for (long long i = (nquad << 2) + (long long)blockIdx.x * kColorThreads + threadIdx.x; i < npx;
             i += (long long)gridDim.x * kColorThreads)
            dst[i] = (uint8_t)bgr_luma(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
    } else {
        for (long long i = (long long)blockIdx.x * kColorThreads + threadIdx.x; i < npx; i += (long long)gridDim.x * kColorThreads) {
            const int r = (int)(i / p.w), c = (int)(i - (long long)r * p.w);
            const uint8_t* px = src + (size_t)r * p.stride + 3 * (size_t)c;
            dst[i] = (uint8_t)bgr_luma(px[0], px[1], px[2]);
        }
    }
}

__global__ void __launch_bounds__(kColorThreads) bgr_recombine_kernel(const ColorParams p) {
    const int f = blockIdx.y;
    const uint8_t* src = p.bgr_in + (unsigned long long)f * p.bgr_pitch;
    uint8_t* dst = p.bgr_out + (unsigned long long)f * p.bgr_pitch;
    const uint8_t* y2p = p.y2_plane + (size_t)f * p.w * p.h;
    const bool vec = (p.stride == 3 * p.w) && ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)y2p) & 3) == 0);
    const long long npx = (long long)p.w * p.h;
    if (vec) {
        const long long nquad = npx >> 2;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        const uint32_t* y32 = reinterpret_cast<const uint32_t*>(y2p);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
        for (long long q = (long long)blockIdx.x * kColorThreads + threadIdx.x; q < nquad; q += (long long)gridDim.x * kColorThreads) {
            const uint32_t a = __ldcs(s32 + 3 * q), b = __ldcs(s32 + 3 * q + 1), c = __ldcs(s32 + 3 * q + 2);
            const uint32_t yy = __ldcs(y32 + q);
            const uint32_t p0 = bgr_recombine(a & 255, (a >> 8) & 255, (a >> 16) & 255, yy & 255, p.mode);
            const uint32_t p1 = bgr_recombine(a >> 24, b & 255, (b >> 8) & 255, (yy >> 8) & 255, p.mode);
            const uint32_t p2 = bgr_recombine((b >> 16) & 255, b >> 24, c & 255, (yy >> 16) & 255, p.mode);
            const uint32_t p3 = bgr_recombine((c >> 8) & 255, (c >> 16) & 255, c >> 24, yy >> 24, p.mode);
            __stcs(d32 + 3 * q, p0 | (p1 << 24));
            __stcs(d32 + 3 * q + 1, (p1 >> 8) | (p2 << 16));
            __stcs(d32 + 3 * q + 2, (p2 >> 16) | (p3 << 8));
        }
        for (long long i = (nquad << 2) + (long long)blockIdx.x * kColorThreads + threadIdx.x; i < npx;
             i += (long long)gridDim.x * kColorThreads) {
            const uint32_t o = bgr_recombine(src[3 * i], src[3 * i + 1], src[3 * i + 2], y2p[i], p.mode);
            dst[3 * i] = (uint8_t)o; dst[3 * i + 1] = (uint8_t)(o >> 8); dst[3 * i + 2] = (uint8_t)(o >> 16);
        }
    } else {
        for (long long i = (long long)blockIdx.x * kColorThreads + threadIdx.x; i < npx; i += (long long)gridDim.x * kColorThreads) {
            const int r = (int)(i / p.w), c = (int)(i - (long long)r * p.w);
            const uint8_t* px = src + (size_t)r * p.stride + 3 * (size_t)c;
            uint8_t* o8 = dst + (size_t)r * p.stride + 3 * (size_t)c;
            const uint32_t o = bgr_recombine(px[0], px[1], px[2], y2p[i], p.mode);
            o8[0] = (uint8_t)o; o8[1] = (uint8_t)(o >> 8); o8[2] = (uint8_t)(o >> 16);
        }
    }
}


// ---- fused colour equalization (flat, 16-byte aligned BGR frames) ---------------------------------------------------
// Two passes over the BGR frame and nothing else: the luma plane is never written.
//   hist item : BGR chunk -> Y (Q14, two IDP.2A per pixel) -> shared histogram -> global histogram of the frame
//   apply item: BGR chunk -> Y -> LUT[Y]; chroma re-derived from B, R and Y (saturated to 8 bits exactly as the
//               intermediate 8UC3 image would hold it) -> inverse conversion with the new luma -> BGR chunk
// Same ticket-lag schedule and the same histogram-as-completion-flag as equalize_kernel.  Algorithmic bytes per frame:
// 6*W*H (read BGR, write BGR); this kernel moves 9*W*H (BGR is read twice) against 14*W*H of the three-pass form.
// A warp owns 512 consecutive pixels (1536 bytes) per round: every lane cp.asyncs three 16-byte pieces (coalesced),
// then reads back its own 48 bytes = 16 pixels (conflict-free 16-byte LDS at a 48-byte stride).
struct ColorEqParams {
    const uint8_t* in;
    uint8_t* out;
    unsigned long long pitch;   // bytes between BGR frames
    int n_frames;
    unsigned long long rounds;  // warp rounds per frame = ceil(W*H / 512)
    unsigned long long npx;     // W*H (multiple of 16)
    int chunks;                 // C items per frame and phase
    unsigned long long rounds_chunk;
    int lag;
    int kB, kR;                 // forward chroma gains: c1 = descale((B - Y) * kB), c2 = descale((R - Y) * kR)
    int iB, iG1, iG2, iR;       // inverse: B = Y + descale(c1*iB), G = Y + descale(c1*iG1 + c2*iG2), R = Y + descale(c2*iR)
    uint32_t* hist;     // [n_frames][256], zero on entry, self-cleaned
    uint32_t* applied;  // [n_frames]
    uint32_t* ticket;
    uint32_t* status;
};
#ifndef NV12EQ_COLOR_BULK
#define NV12EQ_COLOR_BULK 0   // 1: a warp round arrives as ONE bulk copy of the TMA unit (cp.async.bulk + mbarrier) that carries an L2 policy;
#endif                        // 0: three 16-byte cp.async per lane (no policy: ptxas 12.9 miscompiles cp.async with a cache hint here)
constexpr bool kColorBulk = NV12EQ_COLOR_BULK != 0;
constexpr int kColorRingDepth = 3;                       // warp rounds in flight
constexpr int kColorRoundBytes = 1536;                   // 512 pixels
constexpr int kColorRingBytes = kWarps * kColorRingDepth * kColorRoundBytes;
constexpr int kColorEqSmemBytes = kLaneTableBytes + kColorRingBytes;

// pixel i (0..3) of a 12-byte group {q0, q1, q2} as an aligned word {B, G, R, x}
template <int I>
__device__ __forceinline__ uint32_t bgr_px(uint32_t q0, uint32_t q1, uint32_t q2) {
    if (I == 0) return q0;
    if (I == 1) return __byte_perm(q0, q1, 0x6543);
    if (I == 2) return __byte_perm(q1, q2, 0x5432);
    return q2 >> 8;
}
// IDP.2A: a = two 16-bit factors, b = four bytes; lo uses bytes 0,1 of b, hi bytes 2,3.  The pixel bytes are unsigned;
// the chroma gains come as {+k, -k} signed halves.
__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_lo(a, b, c); }
__device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_hi(a, b, c); }
__device__ __forceinline__ int dp2a_lo_su(int a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c) {  // unsigned 16-bit factors x signed bytes
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(int a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// I2IP: d = (c << 16) | (sat(a) << 8) | sat(b) -- two saturating conversions and a pack in one instruction
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack_sat_s8(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// keeps a value opaque so that (x >> 14) << 7 is not re-associated into shift + mask + add (SHF + LEA is one less)
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
// Q14 luma before the shift: 1868*B + 9617*G + 4899*R + 8192 (two IDP.2A); byte 3 of px (the next pixel's B, or 0) is
// multiplied by the zero high half of the second factor word
__device__ __forceinline__ uint32_t bgr_luma14(uint32_t px) {
    return dp2a_hi_uu(4899u, px, dp2a_lo_uu((9617u << 16) | 1868u, px, 8192u));
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) color_equalize_kernel(const ColorEqParams p) {
    extern __shared__ __align__(16) uint32_t smem[];  // 32 KB lane table (hist or LUT), then the per-warp cp.async rings
    __shared__ __align__(16) uint8_t s_lut[256];
    __shared__ uint32_t s_ticket[2];
    __shared__ int s_flag;
    __shared__ __align__(8) unsigned long long s_bar[kWarps * kColorRingDepth];   // one mbarrier per warp and ring stage (bulk copies)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lane_base = smem_u32(smem) + lane * 4;
    const uint32_t bars = smem_u32(s_bar) + (uint32_t)warp * (kColorRingDepth * 8);
    if (kColorBulk) {
        if (lane < kColorRingDepth) mbar_init(bars + lane * 8, 1);
        fence_mbar_init();
        __syncwarp();
    }
    uint32_t seq = 0;   // warp rounds this warp has consumed since the kernel started: stage = seq % depth, phase parity = (seq / depth) & 1
    // The first pass keeps the frame in L2 for the second one (evict_last), the second pass and the stores let go of it.
    const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
    const uint32_t ring = smem_u32(smem) + kLaneTableBytes + (uint32_t)warp * (kColorRingDepth * kColorRoundBytes);
    const int C = p.chunks;
    const int lag = p.lag;
    const uint32_t total_items = (uint32_t)(p.n_frames + lag) * (uint32_t)(2 * C);
    const unsigned long long total_vec = p.npx * 3 / 16;  // 16-byte pieces in a frame

    TicketQueue q{p.ticket, s_ticket, 0u, 0u, false};
    q.start();
    for (;;) {
        const uint32_t item = q.current();
        if (item >= total_items) break;
        const int g = (int)(item / (uint32_t)(2 * C));
        const int r2c = (int)(item % (uint32_t)(2 * C));
        const bool hist_item = r2c < C;
        const int c = hist_item ? r2c : r2c - C;
        const int f = g - lag;
        const unsigned long long rd0 = min((unsigned long long)c * p.rounds_chunk, p.rounds);
        const unsigned long long rd1 = min(rd0 + p.rounds_chunk, p.rounds);
        // rounds of this warp: rd0 + warp, rd0 + warp + kWarps, ...
        const long long nr = rd1 > rd0 + warp ? (long long)((rd1 - rd0 - warp + kWarps - 1) / kWarps) : 0;

        // issue the three 16-byte pieces of this lane for warp round `k` (0-based within the item) into ring stage k % depth
        const uint32_t seq0 = seq;
        auto issue = [&](const uint8_t* frame, long long k) {
            if (kColorBulk) {
                if (k < nr && lane == 0) {
                    const unsigned long long round = rd0 + warp + (unsigned long long)k * kWarps;
                    const unsigned long long b0 = round * kColorRoundBytes, bend = p.npx * 3;
                    const uint32_t bytes = (uint32_t)min((unsigned long long)kColorRoundBytes, bend - b0);
                    const uint32_t sq = seq0 + (uint32_t)k, stg = sq % kColorRingDepth;
                    mbar_arrive_expect_tx(bars + stg * 8, bytes);
                    bulk_load(ring + stg * kColorRoundBytes, frame + b0, bytes, bars + stg * 8, hist_item ? pol_keep : pol_drop);
                }
                return;
            }
            if (k < nr) {
                const unsigned long long round = rd0 + warp + (unsigned long long)k * kWarps;
                const unsigned long long v0 = round * 96;  // first 16-byte piece of the round
                const uint32_t st = ring + (uint32_t)(k % kColorRingDepth) * kColorRoundBytes;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const unsigned long long v = v0 + lane + 32 * j;
                    if (v < total_vec) cp_async16(st + (uint32_t)(lane + 32 * j) * 16u, frame + v * 16);
                }
            }
            cp_async_commit();
        };

        if (hist_item && g < p.n_frames) {
            // ---------------- histogram of chunk c of frame g ----------------
            const uint8_t* frame = p.in + (unsigned long long)g * p.pitch;
            lane_table_zero(smem);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kColorRingDepth - 1; ++k) issue(frame, k);
            for (long long k = 0; k < nr; ++k) {
                if (k + 2 >= nr) q.prefetch();  // late: an early draw would queue the next item behind this one
                issue(frame, k + kColorRingDepth - 1);
                const uint32_t sq = seq0 + (uint32_t)k, stg = kColorBulk ? sq % kColorRingDepth : (uint32_t)(k % kColorRingDepth);
                if (kColorBulk) {
                    if (!mbar_wait(bars + stg * 8, (sq / kColorRingDepth) & 1u)) atomicExch(p.status, 3u);
                } else {
                    cp_async_wait<kColorRingDepth - 1>();
                }
                __syncwarp();
                const unsigned long long px0 = (rd0 + warp + (unsigned long long)k * kWarps) * 512 + (unsigned long long)lane * 16;
                if (px0 < p.npx) {
                    const uint32_t st = ring + stg * kColorRoundBytes + (uint32_t)lane * 48u;
                    const int4 a = lds_s4(st), b = lds_s4(st + 16), cc = lds_s4(st + 32);
                    const uint32_t qw[12] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w, (uint32_t)b.x, (uint32_t)b.y,
                                             (uint32_t)b.z, (uint32_t)b.w, (uint32_t)cc.x, (uint32_t)cc.y, (uint32_t)cc.z, (uint32_t)cc.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t q0 = qw[3 * t], q1 = qw[3 * t + 1], q2 = qw[3 * t + 2];
                        red_shared_inc(lane_base + (opaque(bgr_luma14(bgr_px<0>(q0, q1, q2)) >> 14) << 7));
                        red_shared_inc(lane_base + (opaque(bgr_luma14(bgr_px<1>(q0, q1, q2)) >> 14) << 7));
                        red_shared_inc(lane_base + (opaque(bgr_luma14(bgr_px<2>(q0, q1, q2)) >> 14) << 7));
                        red_shared_inc(lane_base + (opaque(bgr_luma14(bgr_px<3>(q0, q1, q2)) >> 14) << 7));
                    }
                }
                __syncwarp();  // the stage is re-filled two rounds from now by this warp's own copies
            }
            seq += (uint32_t)nr;
            __syncthreads();
            if (tid < 256) {
                const uint32_t cnt = lane_table_row_sum(smem, tid);
                if (cnt) atomicAdd(p.hist + (size_t)g * 256 + tid, cnt);
            }
        } else if (!hist_item && f >= 0) {
            // ---------------- LUT + recombine for chunk c of frame f ----------------
            const uint8_t* frame = p.in + (unsigned long long)f * p.pitch;
            uint8_t* dst = p.out + (unsigned long long)f * p.pitch;
#pragma unroll
            for (int k = 0; k < kColorRingDepth - 1; ++k) issue(frame, k);  // pixels stream in while the LUT is built
            if (warp == 0) {
                uint32_t* gh = p.hist + (size_t)f * 256;
                bool ok = equalize_lut_warp(gh, (long long)p.npx, (long long)p.npx, s_lut, lane);
                if (!ok) {
                    const long long t0 = clock64();
                    unsigned ns = 64;
                    while (!(ok = equalize_lut_warp(gh, (long long)p.npx, (long long)p.npx, s_lut, lane))) {
                        __nanosleep(ns);
                        if (ns < 2048) ns <<= 1;
                        if (clock64() - t0 > kSpinCycles) break;
                    }
                }
                if (lane == 0) {
                    s_flag = ok;
                    if (!ok) atomicExch(p.status, 1u);
                }
                if (ok) {
                    uint32_t last = 0;
                    if (lane == 0) last = (atomicAdd(p.applied + f, 1u) == (uint32_t)(C - 1));
                    if (__shfl_sync(0xffffffffu, last, 0)) {
                        uint4* g4 = reinterpret_cast<uint4*>(gh) + lane * 2;
                        g4[0] = make_uint4(0, 0, 0, 0);
                        g4[1] = make_uint4(0, 0, 0, 0);
                        if (lane == 0) p.applied[f] = 0;
                    }
                }
            }
            __syncthreads();
            if (!s_flag) break;
            lane_table_fill_from_lut(smem, s_lut);
            __syncthreads();
            const int kBn = -p.kB, kRn = -p.kR;
            const int cB = (int)(((uint32_t)(kBn & 0xffff) << 16) | (uint32_t)(p.kB & 0xffff));  // {kB, -kB} as s16x2
            const int cR = (int)(((uint32_t)(kRn & 0xffff) << 16) | (uint32_t)(p.kR & 0xffff));
            const uint32_t fB = (uint32_t)p.iB & 0xffffu;                                   // {iB, 0}: multiplies c1 only
            const uint32_t fR = ((uint32_t)p.iR & 0xffffu) << 16;                           // {0, iR}: multiplies c2 only
            const int fG = (int)((((uint32_t)p.iG2 & 0xffffu) << 16) | ((uint32_t)p.iG1 & 0xffffu));
            for (long long k = 0; k < nr; ++k) {
                if (k + 2 >= nr) q.prefetch();
                issue(frame, k + kColorRingDepth - 1);
                const uint32_t sq = seq0 + (uint32_t)k, stg = kColorBulk ? sq % kColorRingDepth : (uint32_t)(k % kColorRingDepth);
                if (kColorBulk) {
                    if (!mbar_wait(bars + stg * 8, (sq / kColorRingDepth) & 1u)) atomicExch(p.status, 3u);
                } else {
                    cp_async_wait<kColorRingDepth - 1>();
                }
                __syncwarp();
                const unsigned long long round = rd0 + warp + (unsigned long long)k * kWarps;
                const unsigned long long px0 = round * 512 + (unsigned long long)lane * 16;
                if (px0 < p.npx) {
                    const uint32_t st = ring + stg * kColorRoundBytes + (uint32_t)lane * 48u;
                    const int4 a = lds_s4(st), b = lds_s4(st + 16), cc = lds_s4(st + 32);
                    const uint32_t qw[12] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w, (uint32_t)b.x, (uint32_t)b.y,
                                             (uint32_t)b.z, (uint32_t)b.w, (uint32_t)cc.x, (uint32_t)cc.y, (uint32_t)cc.z, (uint32_t)cc.w};
                    uint32_t ow[12];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t q0 = qw[3 * t], q1 = qw[3 * t + 1], q2 = qw[3 * t + 2];
                        uint32_t o[4];
                        const uint32_t px[4] = {bgr_px<0>(q0, q1, q2), bgr_px<1>(q0, q1, q2), bgr_px<2>(q0, q1, q2), bgr_px<3>(q0, q1, q2)};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t Y = opaque(bgr_luma14(px[i]) >> 14);
                            const int y2 = (int)lds_u8(lane_base + (Y << 7));
                            const uint32_t byry = __byte_perm(px[i], Y, 0x4240);            // {B, Y, R, Y}
                            const int c1 = dp2a_lo_su(cB, byry, 8192) >> 14;                 // descale((B - Y) * kB): never saturates
                            const int c2 = dp2a_hi_su(cR, byry, 8192) >> 14;                 // descale((R - Y) * kR)
                            // {c1, sat8s(c2)} as signed bytes: sat8(128 + c2) - 128 is the chroma the 8UC3 image would hold
                            const uint32_t cc8 = pack_sat_s8(c2, c1, 0u);
                            const int b2 = y2 + (dp2a_lo_us(fB, cc8, 8192) >> 14);           // Y + descale(c1*iB)
                            const int g2 = y2 + (dp2a_lo_ss(fG, cc8, 8192) >> 14);           // Y + descale(c1*iG1 + c2*iG2)
                            const int r2v = y2 + (dp2a_lo_us(fR, cc8, 8192) >> 14);          // Y + descale(c2*iR)
                            o[i] = pack_sat_u8(g2, b2, pack_sat_u8(0, r2v, 0u));             // {B, G, R, 0}, each saturated
                        }
                        ow[3 * t] = __byte_perm(o[0], o[1], 0x4210);       // B0 G0 R0 B1
                        ow[3 * t + 1] = __byte_perm(o[1], o[2], 0x5421);   // G1 R1 B2 G2
                        ow[3 * t + 2] = __byte_perm(o[2], o[3], 0x6542);   // R2 B3 G3 R3
                    }
                    uint4* d4 = reinterpret_cast<uint4*>(dst + px0 * 3);
                    __stcs(d4, make_uint4(ow[0], ow[1], ow[2], ow[3]));
                    __stcs(d4 + 1, make_uint4(ow[4], ow[5], ow[6], ow[7]));
                    __stcs(d4 + 2, make_uint4(ow[8], ow[9], ow[10], ow[11]));
                }
                __syncwarp();
            }
            seq += (uint32_t)nr;
        }
        q.advance();
    }
    q.finish();
}

// ---- cvtColor(COLOR_BGR2YUV_I420), 1frameMeasure.cpp:32 ------------------------------------------
// Q20 limited-range BT.601; chroma taken from the top-left pixel of every 2x2 block; planar output.  One thread
// converts a 4x2 pixel block: two 12-byte row pieces in, two luma words and two chroma byte pairs out.
struct I420Params {
    const uint8_t* bgr; uint8_t* out;
    unsigned long long bgr_pitch, out_pitch;   // bytes between frames
    int w, h, stride;                          // even w, h; stride in bytes of a BGR row
};
__device__ __forceinline__ uint32_t i420_luma(int B, int G, int R) { return (uint32_t)((269484 * R + 528482 * G + 102760 * B + (16 << 20) + (1 << 19)) >> 20); }
__device__ __forceinline__ uint32_t i420_u(int B, int G, int R) { return (uint32_t)((-155188 * R - 305135 * G + 460324 * B + (128 << 20) + (1 << 19)) >> 20); }
__device__ __forceinline__ uint32_t i420_v(int B, int G, int R) { return (uint32_t)((460324 * R - 385875 * G - 74448 * B + (128 << 20) + (1 << 19)) >> 20); }

__global__ void __launch_bounds__(kColorThreads) bgr_to_i420_kernel(const I420Params p) {
    const int f = blockIdx.y;
    const uint8_t* src = p.bgr + (unsigned long long)f * p.bgr_pitch;
    uint8_t* Y = p.out + (unsigned long long)f * p.out_pitch;
    uint8_t* U = Y + (size_t)p.w * p.h;
    uint8_t* V = U + (size_t)(p.w / 2) * (p.h / 2);
    const int bw = (p.w + 3) / 4, bh = p.h / 2;          // 4x2 blocks (the last block of a row may be 2 wide)
    const long long nblk = (long long)bw * bh;
    const bool vec = ((p.w & 3) == 0) && ((p.stride & 3) == 0) && ((((uintptr_t)src | (uintptr_t)Y) & 3) == 0) &&
                     ((((uintptr_t)U | (uintptr_t)V) & 1) == 0);
    for (long long b = (long long)blockIdx.x * kColorThreads + threadIdx.x; b < nblk; b += (long long)gridDim.x * kColorThreads) {
        const int by = (int)(b / bw), bx = (int)(b - (long long)by * bw);
        const int x = bx * 4, y = by * 2;
        if (vec) {
            uint32_t luma[2];
            uint32_t u01 = 0, v01 = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + (size_t)(y + r) * p.stride + 3 * (size_t)x);
                const uint32_t a = __ldg(s32), bb = __ldg(s32 + 1), c = __ldg(s32 + 2);
                const int B0 = a & 255, G0 = (a >> 8) & 255, R0 = (a >> 16) & 255;
                const int B1 = a >> 24, G1 = bb & 255, R1 = (bb >> 8) & 255;
                const int B2 = (bb >> 16) & 255, G2 = bb >> 24, R2 = c & 255;
                const int B3 = (c >> 8) & 255, G3 = (c >> 16) & 255, R3 = c >> 24;
                luma[r] = i420_luma(B0, G0, R0) | (i420_luma(B1, G1, R1) << 8) | (i420_luma(B2, G2, R2) << 16) | (i420_luma(B3, G3, R3) << 24);
                if (r == 0) {
                    u01 = i420_u(B0, G0, R0) | (i420_u(B2, G2, R2) << 8);
                    v01 = i420_v(B0, G0, R0) | (i420_v(B2, G2, R2) << 8);
                }
            }
            *reinterpret_cast<uint32_t*>(Y + (size_t)y * p.w + x) = luma[0];
            *reinterpret_cast<uint32_t*>(Y + (size_t)(y + 1) * p.w + x) = luma[1];
            const size_t k = (size_t)by * (p.w / 2) + x / 2;
            *reinterpret_cast<uint16_t*>(U + k) = (uint16_t)u01;
            *reinterpret_cast<uint16_t*>(V + k) = (uint16_t)v01;
        } else {
            const int xe = min(x + 4, p.w);
            for (int r = 0; r < 2; ++r)
                for (int c = x; c < xe; ++c) {
                    const uint8_t* px = src + (size_t)(y + r) * p.stride + 3 * (size_t)c;
                    Y[(size_t)(y + r) * p.w + c] = (uint8_t)i420_luma(px[0], px[1], px[2]);
                    if (r == 0 && !(c & 1)) {
                        const size_t k = (size_t)by * (p.w / 2) + c / 2;
                        U[k] = (uint8_t)i420_u(px[0], px[1], px[2]);
                        V[k] = (uint8_t)i420_v(px[0], px[1], px[2]);
                    }
                }
        }
    }
}

// ---- NV12 <-> BGR adapters ----------------------------------------------------------------
// nv12_to_bgr_kernel: cvtColor(nv12, COLOR_YUV2BGR_NV12), OpenCV's YUV420sp2RGB (Q20 limited-range BT.601, one chroma pair per
// 2x2 block, saturating) -- the display-side inverse of the NV12 path; the reference's still-image tools convert back with
// cvtColor at singlecolor.cpp:66 / clahe1frame.cpp:102.  bgr_to_nv12_kernel: the arithmetic of bgr_to_i420_kernel with the chroma
// interleaved (U first), i.e. the frame the NV12 operators take.  One thread converts a 4x2 pixel block.
struct Nv12BgrParams {
    const uint8_t* in; uint8_t* out;
    unsigned long long in_pitch, out_pitch;   // bytes between frames
    int w, h;                                 // even
    int nv12_stride, bgr_stride;              // bytes per row
};
__device__ __forceinline__ uint32_t sat_u8(int v) { return (uint32_t)min(max(v, 0), 255); }
// one pixel: luma byte + chroma terms -> {B, G, R} in the low three bytes
__device__ __forceinline__ uint32_t nv12_px(uint32_t y, int ruv, int guv, int buv) {
    const int yy = (int)(y > 16u ? y - 16u : 0u) * 1220542;
    return sat_u8((yy + buv) >> 20) | (sat_u8((yy + guv) >> 20) << 8) | (sat_u8((yy + ruv) >> 20) << 16);
}
__global__ void __launch_bounds__(kColorThreads) nv12_to_bgr_kernel(const Nv12BgrParams p) {
    const int f = blockIdx.y;
    const uint8_t* Y = p.in + (unsigned long long)f * p.in_pitch;
    const uint8_t* UV = Y + (size_t)p.nv12_stride * p.h;
    uint8_t* dst = p.out + (unsigned long long)f * p.out_pitch;
    const int bw = (p.w + 3) / 4, bh = p.h / 2;          // 4x2 blocks (the last block of a row may be 2 wide)
    const long long nblk = (long long)bw * bh;
    const bool vec = ((p.w & 3) == 0) && ((p.nv12_stride & 3) == 0) && ((p.bgr_stride & 3) == 0) && ((((uintptr_t)Y | (uintptr_t)dst) & 3) == 0);
    for (long long b = (long long)blockIdx.x * kColorThreads + threadIdx.x; b < nblk; b += (long long)gridDim.x * kColorThreads) {
        const int by = (int)(b / bw), bx = (int)(b - (long long)by * bw);
        const int x = bx * 4, y = by * 2;
        const int xe = min(x + 4, p.w);
        uint32_t uvw;
        if (vec) uvw = __ldg(reinterpret_cast<const uint32_t*>(UV + (size_t)by * p.nv12_stride + x));
        else {
            const uint8_t* q = UV + (size_t)by * p.nv12_stride + x;
            uvw = q[0] | (q[1] << 8);
            if (xe - x > 2) uvw |= (q[2] << 16) | (q[3] << 24);
        }
        int ruv[2], guv[2], buv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int u = (int)((uvw >> (16 * k)) & 255u) - 128, v = (int)((uvw >> (16 * k + 8)) & 255u) - 128;
            ruv[k] = (1 << 19) + 1673527 * v;
            guv[k] = (1 << 19) - 852492 * v - 409993 * u;
            buv[k] = (1 << 19) + 2116026 * u;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint8_t* yrow = Y + (size_t)(y + r) * p.nv12_stride + x;
            uint8_t* d = dst + (size_t)(y + r) * p.bgr_stride + 3 * (size_t)x;
            if (vec) {
                const uint32_t yw = __ldg(reinterpret_cast<const uint32_t*>(yrow));
                const uint32_t p0 = nv12_px(yw & 255u, ruv[0], guv[0], buv[0]), p1 = nv12_px((yw >> 8) & 255u, ruv[0], guv[0], buv[0]);
                const uint32_t p2 = nv12_px((yw >> 16) & 255u, ruv[1], guv[1], buv[1]), p3 = nv12_px(yw >> 24, ruv[1], guv[1], buv[1]);
                uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
                d32[0] = p0 | (p1 << 24);
                d32[1] = (p1 >> 8) | (p2 << 16);
                d32[2] = (p2 >> 16) | (p3 << 8);
            } else {
                for (int c = x; c < xe; ++c) {
                    const int k = (c - x) >> 1;
                    const uint32_t px = nv12_px(yrow[c - x], ruv[k], guv[k], buv[k]);
                    uint8_t* o = d + 3 * (c - x);
                    o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
                }
            }
        }
    }
}
__global__ void __launch_bounds__(kColorThreads) bgr_to_nv12_kernel(const Nv12BgrParams p) {
    const int f = blockIdx.y;
    const uint8_t* src = p.in + (unsigned long long)f * p.in_pitch;
    uint8_t* Y = p.out + (unsigned long long)f * p.out_pitch;
    uint8_t* UV = Y + (size_t)p.nv12_stride * p.h;
    const int bw = (p.w + 3) / 4, bh = p.h / 2;
    const long long nblk = (long long)bw * bh;
    const bool vec = ((p.w & 3) == 0) && ((p.nv12_stride & 3) == 0) && ((p.bgr_stride & 3) == 0) && ((((uintptr_t)Y | (uintptr_t)src) & 3) == 0);
    for (long long b = (long long)blockIdx.x * kColorThreads + threadIdx.x; b < nblk; b += (long long)gridDim.x * kColorThreads) {
        const int by = (int)(b / bw), bx = (int)(b - (long long)by * bw);
        const int x = bx * 4, y = by * 2;
        if (vec) {
            uint32_t uvw = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + (size_t)(y + r) * p.bgr_stride + 3 * (size_t)x);
                const uint32_t a = __ldg(s32), bb = __ldg(s32 + 1), c = __ldg(s32 + 2);
                const int B0 = a & 255, G0 = (a >> 8) & 255, R0 = (a >> 16) & 255;
                const int B1 = a >> 24, G1 = bb & 255, R1 = (bb >> 8) & 255;
                const int B2 = (bb >> 16) & 255, G2 = bb >> 24, R2 = c & 255;
                const int B3 = (c >> 8) & 255, G3 = (c >> 16) & 255, R3 = c >> 24;
                *reinterpret_cast<uint32_t*>(Y + (size_t)(y + r) * p.nv12_stride + x) =
                    i420_luma(B0, G0, R0) | (i420_luma(B1, G1, R1) << 8) | (i420_luma(B2, G2, R2) << 16) | (i420_luma(B3, G3, R3) << 24);
                if (r == 0) uvw = i420_u(B0, G0, R0) | (i420_v(B0, G0, R0) << 8) | (i420_u(B2, G2, R2) << 16) | (i420_v(B2, G2, R2) << 24);
            }
            *reinterpret_cast<uint32_t*>(UV + (size_t)by * p.nv12_stride + x) = uvw;
        } else {
            const int xe = min(x + 4, p.w);
            for (int r = 0; r < 2; ++r)
                for (int c = x; c < xe; ++c) {
                    const uint8_t* px = src + (size_t)(y + r) * p.bgr_stride + 3 * (size_t)c;
                    Y[(size_t)(y + r) * p.nv12_stride + c] = (uint8_t)i420_luma(px[0], px[1], px[2]);
                    if (r == 0 && !(c & 1)) {
                        uint8_t* d = UV + (size_t)by * p.nv12_stride + c;
                        d[0] = (uint8_t)i420_u(px[0], px[1], px[2]);
                        d[1] = (uint8_t)i420_v(px[0], px[1], px[2]);
                    }
                }
        }
    }
}

}  // namespace nv12eq
