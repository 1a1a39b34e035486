// clahe16.cuh -- CLAHE on CV_16UC1 planes (P010 luma and other 16-bit content): SURVEY.md section 8f rank 3.
//
// OpenCV's CLAHE accepts 16-bit input with histSize = 65536; the reference never feeds it 16-bit data, so this is a
// widening row, built correctness-first: 65536-bin tile histograms and 128 KB tile LUTs do not fit shared memory, they
// live in global memory and are served by L2.
//   clahe16_hist_kernel   one CTA per (tile strip, frame): red.global.add into hist[frame][tile][65536]
//                         (reflect-101 padding by index reflection, as in the 8-bit kernel)
//   clahe16_lut_kernel    one CTA per (tile, frame): clip at clipLimit, redistribute the excess exactly as OpenCV
//                         (redistBatch to every bin, +1 to every residualStep-th bin while the residual lasts), block-wide
//                         scan, lut = saturate_cast<ushort>(cvRound(sum * lutScale)); returns the histogram to zero
//   clahe16_interp_kernel per pixel: four 16-bit gathers from the neighbouring tile LUTs and OpenCV's blend op for op in
//                         unfused fp32 (same weights and rounding as the 8-bit path)
// Bound: L2 atomics (histogram) and L2 gathers (interpolation), not HBM; algorithmic bytes are 4*W*H per plane.
#pragma once
#include "clahe.cuh"

namespace nv12eq {

constexpr int kBins16 = 65536;
constexpr int kC16Threads = 256;
constexpr int kC16LutThreads = 1024;

struct Clahe16Params {
    const uint16_t* in;
    uint16_t* out;
    unsigned long long pitch;   // elements between planes
    int n_planes;
    int w, h, stride;           // stride in elements
    int tx, ty, tw, th;
    int clip_limit;
    float lut_scale, inv_tw, inv_th;
    uint32_t* hist;             // [n_planes][tx*ty][65536], zero on entry, zero again after clahe16_lut_kernel
    uint16_t* luts;             // [n_planes][tx*ty][65536]
    int strips;                 // row strips per tile in the histogram kernel
};

__global__ void __launch_bounds__(kC16Threads) clahe16_hist_kernel(const Clahe16Params p) {
    const int T = p.tx * p.ty;
    const int f = blockIdx.z, t = blockIdx.y, strip = blockIdx.x;
    const int tyi = t / p.tx, txi = t - tyi * p.tx;
    const int x0 = txi * p.tw, y0 = tyi * p.th;
    const int rows_strip = (p.th + p.strips - 1) / p.strips;
    const int r0 = strip * rows_strip, r1 = min(r0 + rows_strip, p.th);
    const uint16_t* src = p.in + (unsigned long long)f * p.pitch;
    uint32_t* hist = p.hist + ((size_t)f * T + t) * kBins16;
    const int n = (r1 - r0) * p.tw;
    for (int i = threadIdx.x; i < n; i += kC16Threads) {
        const int r = r0 + i / p.tw, c = i - (i / p.tw) * p.tw;
        const uint16_t v = src[(size_t)reflect101(y0 + r, p.h) * p.stride + reflect101(x0 + c, p.w)];
        atomicAdd(hist + v, 1u);
    }
}

// Block-wide inclusive scan of one int per thread (1024 threads); *total = sum over the block.
__device__ __forceinline__ int block_incl_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int tt = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += tt;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += tt;
        }
        s_warp[lane] = wi - w;            // exclusive prefix of the warps
        if (lane == 31) s_warp[32] = wi;  // block total
    }
    __syncthreads();
    const int r = s_warp[warp] + incl;
    *total = s_warp[32];
    __syncthreads();                      // s_warp is reused by the next call
    return r;
}

// One CTA per (tile, plane).  All accesses are coalesced: thread t owns bins t, t + 1024, ... (64 chunks of 1024 bins).
__global__ void __launch_bounds__(kC16LutThreads) clahe16_lut_kernel(const Clahe16Params p) {
    __shared__ int s_warp[33];
    const int T = p.tx * p.ty;
    const int f = blockIdx.y, t = blockIdx.x;
    uint32_t* hist = p.hist + ((size_t)f * T + t) * kBins16;
    uint16_t* lut = p.luts + ((size_t)f * T + t) * kBins16;
    constexpr int kChunks = kBins16 / kC16LutThreads;
    int batch = 0, residual = 0, step = 1;
    uint32_t magic = 0;   // floor(2^32 / step) + 1: (i * magic) >> 32 == i / step for i < 65536, 2 <= step <= 65536
    if (p.clip_limit > 0) {
        int part = 0;     // a tile has fewer than 2^31 pixels: int sums are safe
#pragma unroll 4
        for (int c = 0; c < kChunks; ++c) part += max((int)hist[c * kC16LutThreads + threadIdx.x] - p.clip_limit, 0);
        int clipped;
        block_incl_scan(part, s_warp, &clipped);
        batch = clipped / kBins16;
        residual = clipped - batch * kBins16;
        if (residual != 0) step = max(kBins16 / residual, 1);
        if (step > 1) magic = (uint32_t)(0x100000000ull / (uint32_t)step) + 1u;
    }
    int carry = 0;
    for (int c = 0; c < kChunks; ++c) {
        const int i = c * kC16LutThreads + threadIdx.x;
        int hv = (int)hist[i];
        hist[i] = 0;      // ready for the next launch
        if (p.clip_limit > 0) {
            hv = min(hv, p.clip_limit) + batch;
            // for (i = 0; i < histSize && residual > 0; i += step, residual--) h[i]++
            if (residual != 0) {
                const int q = step > 1 ? (int)(((unsigned long long)(uint32_t)i * magic) >> 32) : i;
                if (q * step == i && q < residual) hv += 1;
            }
        }
        int chunk_total;
        const int run = carry + block_incl_scan(hv, s_warp, &chunk_total);
        carry += chunk_total;
        const int r = __float2int_rn(__fmul_rn(__int2float_rn(run), p.lut_scale));
        lut[i] = (uint16_t)min(max(r, 0), 65535);
    }
}

__global__ void __launch_bounds__(kC16Threads) clahe16_interp_kernel(const Clahe16Params p) {
    const int T = p.tx * p.ty;
    const int f = blockIdx.z, y = blockIdx.y;
    const uint16_t* src = p.in + (unsigned long long)f * p.pitch + (size_t)y * p.stride;
    uint16_t* dst = p.out + (unsigned long long)f * p.pitch + (size_t)y * p.stride;
    const uint16_t* luts = p.luts + (size_t)f * T * kBins16;
    float ya, ya1;
    axis_weight(y, p.inv_th, ya, ya1);
    const int tyf = (int)floorf(__fsub_rn(__fmul_rn((float)y, p.inv_th), 0.5f));
    const int ty1 = max(tyf, 0), ty2 = min(tyf + 1, p.ty - 1);
    for (int x = blockIdx.x * kC16Threads + threadIdx.x; x < p.w; x += gridDim.x * kC16Threads) {
        float xa, xa1;
        axis_weight(x, p.inv_tw, xa, xa1);
        const int txf = (int)floorf(__fsub_rn(__fmul_rn((float)x, p.inv_tw), 0.5f));
        const int tx1 = max(txf, 0), tx2 = min(txf + 1, p.tx - 1);
        const uint32_t v = src[x];
        const float l11 = (float)__ldg(luts + (size_t)(ty1 * p.tx + tx1) * kBins16 + v);
        const float l12 = (float)__ldg(luts + (size_t)(ty1 * p.tx + tx2) * kBins16 + v);
        const float l21 = (float)__ldg(luts + (size_t)(ty2 * p.tx + tx1) * kBins16 + v);
        const float l22 = (float)__ldg(luts + (size_t)(ty2 * p.tx + tx2) * kBins16 + v);
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
        const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
        dst[x] = (uint16_t)min(max(__float2int_rn(res), 0), 65535);
    }
}

}  // namespace nv12eq
