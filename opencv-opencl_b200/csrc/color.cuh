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

}  // namespace nv12eq
