// color.cuh -- the still-image colour path: BGR -> YUV (or YCrCb), equalize / CLAHE on channel 0, -> BGR.
//
// Replaces cvtColor(BGR2YUV) -> split -> equalizeHist / clahe->apply -> merge -> cvtColor(YUV2BGR) of
// singlecolor.cpp:39-66 and clahe1frame.cpp:83-102 (8-bit Q14 fixed point, descale = (x+8192)>>14).
// The intermediate 3-channel YUV image, the split planes and the merged image are never materialised: pass 1 writes
// only the Y plane (W*H bytes, consumed by the equalize / CLAHE kernel), pass 2 re-derives U and V from the BGR
// input and combines them with the equalized Y.
#pragma once
#include "common.cuh"

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

}  // namespace nv12eq
