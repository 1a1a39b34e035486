// equalize.cuh -- global histogram equalization of the Y plane of a batch of NV12 frames, one launch per batch.
//
// Replaces cv::equalizeHist as called at nextimprovement.cpp:168 (hist -> LUT -> apply) together
// with the NV12 rebuild around it (UV memcpy nextimprovement.cpp:160 / memset 128 OpenCVequalHist.cpp:162).
//
// Schedule ("ticket-lag"): work is cut into items drawn by CTAs from a global ticket counter, so items START
// strictly in ticket order.  Slot g of the ticket space holds 2C items:
//     r <  C : histogram chunk r of frame g           (if g < n_frames) -> smem hist[256][32] -> global hist[g][256]
//     r >= C : LUT + apply + UV of chunk r-C of frame g - lag (if >= 0);  waits until the global histogram of
//              that frame adds up to W*H pixels (the histogram is its own completion flag)
// A waiting CTA can only wait on items with smaller tickets, which are resident or finished, so the wait always
// ends (any lag >= 0, any grid size).  With lag >= 1 a frame's histogram is complete before its apply items are
// drawn, the Y plane it just streamed through is still in the 126 MB L2 when it is read the second time, and
// atomics-bound histogram work and HBM-bound apply/copy work overlap on every SM.  The same kernel runs the plain two-launch schedule
// (phases = HIST, then phases = APPLY) and the stage-level forms of the spatially split mode.
//
// Roofline: HBM.  Algorithmic bytes per frame = 3*W*H (read NV12 once, write NV12 once; the second read of Y is
// expected to hit L2).  The secondary limiter is the LSU: one shared atomic and one shared gather per pixel.
#pragma once
#include "common.cuh"

namespace nv12eq {

enum : int { PH_HIST = 1, PH_APPLY = 2, PH_EXTERNAL_HIST = 4 };

struct EqParams {
    const uint8_t* in;
    uint8_t* out;
    unsigned long long pitch;  // bytes between frames
    int n_frames;
    int w, h, stride;
    int flat;     // stride == w: planes are contiguous byte spans
    int uv_mode;  // UV_COPY / UV_GRAY128 / UV_SKIP
    int chunks;   // C, chunks per frame
    int lag;      // frames between a frame's histogram items and its apply items
    int phases;   // PH_* mask
    unsigned long long y_bytes, uv_bytes;      // flat: W*H and W*(H/2)
    unsigned long long y_chunk, uv_chunk;      // flat: bytes per chunk (multiples of 4096)
    int y_rows_chunk, uv_rows_chunk;           // strided: rows per chunk
    long long total_px;                        // pixel count the histogram describes (W*H unless spatially split)
    uint32_t* hist;     // [n_frames][256], zero on entry; self-cleaned unless PH_EXTERNAL_HIST
    uint32_t* applied;  // [n_frames] apply items that have read the histogram; self-cleaned
    uint32_t* ticket;   // [1] work counter; self-cleaned
    uint32_t* status;   // [1] sticky error word (spin timeout)
};

// One warp: 256-bin histogram (global) -> equalization LUT (shared), SURVEY.md A.1 / oracle_equalize_lut.
// Returns false when the histogram does not (yet) add up to `expect` pixels.  The histogram itself is the completion
// signal: it starts at zero, every pixel is added exactly once with an L2 atomic, so sum == expect means every
// chunk of the frame has been added and the values are final -- no fence, no flag, no extra round trip.
__device__ __forceinline__ bool equalize_lut_warp(const uint32_t* __restrict__ ghist, long long total, long long expect,
                                                  uint8_t* __restrict__ slut, int lane) {
    uint32_t h[8];
    const uint4* g4 = reinterpret_cast<const uint4*>(ghist) + lane * 2;
    uint4 a = __ldcg(g4), b = __ldcg(g4 + 1);
    h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
    uint32_t lsum = 0;
    int first = 8;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        lsum += h[j];
        if (h[j] != 0) first = j;
    }
    const uint32_t nz = __ballot_sync(0xffffffffu, first < 8);
    const uint32_t incl = warp_incl_scan(lsum, lane);
    if (expect >= 0 && (long long)__shfl_sync(0xffffffffu, incl, 31) != expect) return false;
    if (nz == 0) {  // empty plane: nothing will be read from the LUT
#pragma unroll
        for (int j = 0; j < 8; ++j) slut[lane * 8 + j] = 0;
        return true;
    }
    const int l0 = __ffs(nz) - 1;
    const int j0 = __shfl_sync(0xffffffffu, first, l0);
    const int i0 = l0 * 8 + j0;
    // cumulative count up to and including bin i0, and hist[i0]
    uint32_t run = incl - lsum;  // exclusive prefix of this lane
    uint32_t cum_i0_local = 0, h_i0_local = 0;
    {
        uint32_t r = run;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r += h[j];
            if (j == j0) { cum_i0_local = r; h_i0_local = h[j]; }
        }
    }
    const uint32_t cum_i0 = __shfl_sync(0xffffffffu, cum_i0_local, l0);
    const uint32_t h_i0 = __shfl_sync(0xffffffffu, h_i0_local, l0);
    if ((long long)h_i0 == total) {  // constant image: dst = i0 everywhere
#pragma unroll
        for (int j = 0; j < 8; ++j) slut[lane * 8 + j] = (uint8_t)i0;
        return true;
    }
    const float scale = __fdiv_rn(255.0f, (float)(total - (long long)h_i0));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        run += h[j];
        const int i = lane * 8 + j;
        uint32_t v = 0;
        if (i > i0) v = round_sat_u8(__fmul_rn(__uint2float_rn(run - cum_i0), scale));
        slut[i] = (uint8_t)v;
    }
    return true;
}

// Chunk c of a plane of `rows` rows.  Flat planes are cut by bytes, strided planes by rows.
struct PlaneChunk {
    unsigned long long b0, b1;  // flat: byte range
    int r0, r1;                 // strided: row range
};
__device__ __forceinline__ PlaneChunk plane_chunk(int flat, int c, unsigned long long bytes, unsigned long long chunk,
                                                  int rows, int rows_chunk) {
    PlaneChunk pc;
    if (flat) {
        pc.b0 = min((unsigned long long)c * chunk, bytes);
        pc.b1 = min(pc.b0 + chunk, bytes);
        pc.r0 = pc.r1 = 0;
    } else {
        pc.b0 = pc.b1 = 0;
        pc.r0 = min(c * rows_chunk, rows);
        pc.r1 = min(pc.r0 + rows_chunk, rows);
    }
    return pc;
}

// Chroma rows of chunk c: passthrough copy or neutral grey, by threads [tid0, tid0 + nthr) of the CTA.
__device__ __forceinline__ void equalize_uv_chunk(const EqParams& p, const uint8_t* src, uint8_t* dst, int c, int tid,
                                                  int nthr) {
    const bool copy_uv = p.uv_mode == UV_COPY && src != dst;
    if (!(copy_uv || p.uv_mode == UV_GRAY128) || tid < 0) return;
    const size_t uv_off = (size_t)p.stride * p.h;
    const PlaneChunk uc = plane_chunk(p.flat, c, p.uv_bytes, p.uv_chunk, p.h / 2, p.uv_rows_chunk);
    if (p.flat) {
        const size_t n = (size_t)(uc.b1 - uc.b0);
        if (copy_uv) copy_span(src + uv_off + uc.b0, dst + uv_off + uc.b0, n, tid, nthr);
        else fill_span(dst + uv_off + uc.b0, n, tid, nthr, 128);
    } else {
        const int lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
        for (int r = uc.r0 + w; r < uc.r1; r += nw) {
            const size_t off = uv_off + (size_t)r * p.stride;
            if (copy_uv) copy_span(src + off, dst + off, (size_t)p.w, lane, 32);
            else fill_span(dst + off, (size_t)p.w, lane, 32, 128);
        }
    }
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) equalize_kernel(const EqParams p) {
    extern __shared__ __align__(16) uint32_t smem[];  // 32 KB: hist[256][32] (histogram items) or table[256][32] (apply items)
    __shared__ __align__(16) uint8_t s_lut[256];
    __shared__ uint32_t s_ticket[2];
    __shared__ int s_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lane_base = smem_u32(smem) + lane * 4;
    const int C = p.chunks;
    const bool do_hist = (p.phases & PH_HIST) != 0, do_apply = (p.phases & PH_APPLY) != 0;
    const bool external = (p.phases & PH_EXTERNAL_HIST) != 0;
    const int lag = (do_hist && do_apply) ? p.lag : 0;
    const uint32_t total_items = (uint32_t)(p.n_frames + lag) * (uint32_t)(2 * C);

    TicketQueue q{p.ticket, s_ticket, 0u, 0u, false};
    q.start();
    for (;;) {
        const uint32_t item = q.current();
        if (item >= total_items) break;
#ifdef NV12EQ_EQ_EARLY_TICKET
        q.prefetch();
#endif
        const int g = (int)(item / (uint32_t)(2 * C));
        const int r2 = (int)(item % (uint32_t)(2 * C));
        const bool hist_item = r2 < C;
        const int c = hist_item ? r2 : r2 - C;
        const int f = g - lag;

        if (hist_item && do_hist && g < p.n_frames) {
            // ---------------- histogram of chunk c of frame g ----------------
            const uint8_t* y = p.in + (unsigned long long)g * p.pitch;
            lane_table_zero(smem);
            __syncthreads();
            const PlaneChunk pc = plane_chunk(p.flat, c, p.y_bytes, p.y_chunk, p.h, p.y_rows_chunk);
            if (p.flat) {
                // the next ticket is drawn when ~3/4 of the item is done (see TicketQueue): the last quarter hides the atomic
                const size_t n = (size_t)(pc.b1 - pc.b0), n1 = (n - n / 4) & ~(size_t)4095;
                hist_span(y + pc.b0, n1, tid, kThreads, lane_base);
                q.prefetch();
                hist_span(y + pc.b0 + n1, n - n1, tid, kThreads, lane_base);
            } else {
                for (int r = pc.r0 + warp; r < pc.r1; r += kWarps)
                    hist_span(y + (size_t)r * p.stride, (size_t)p.w, lane, 32, lane_base);
            }
            __syncthreads();
            if (tid < 256) {
                const uint32_t cnt = lane_table_row_sum(smem, tid);
                if (cnt) atomicAdd(p.hist + (size_t)g * 256 + tid, cnt);
            }
        } else if (!hist_item && do_apply && f >= 0) {
            // ---------------- LUT + apply + UV for chunk c of frame f ----------------
            const uint8_t* src = p.in + (unsigned long long)f * p.pitch;
            uint8_t* dst = p.out + (unsigned long long)f * p.pitch;
            if (warp == 0) {
                // wait for the complete histogram of frame f and build its LUT; the other warps move chroma meanwhile
                uint32_t* gh = p.hist + (size_t)f * 256;
                const long long expect = external ? -1ll : p.total_px;
                bool ok = equalize_lut_warp(gh, p.total_px, expect, s_lut, lane);
                if (!ok) {
                    const long long t0 = clock64();
                    unsigned ns = 64;
                    while (!(ok = equalize_lut_warp(gh, p.total_px, expect, s_lut, lane))) {
                        __nanosleep(ns);
                        if (ns < 2048) ns <<= 1;
                        if (clock64() - t0 > kSpinCycles) break;
                    }
                }
                if (lane == 0) {
                    s_flag = ok;
                    if (!ok) atomicExch(p.status, 1u);
                }
                if (ok && !external) {
                    // every apply item reads the histogram once; the last reader returns it to its zero state
                    uint32_t last = 0;
                    if (lane == 0) last = (atomicAdd(p.applied + f, 1u) == (uint32_t)(C - 1));
                    if (__shfl_sync(0xffffffffu, last, 0)) {
                        uint4* g4 = reinterpret_cast<uint4*>(gh) + lane * 2;
                        g4[0] = make_uint4(0, 0, 0, 0);
                        g4[1] = make_uint4(0, 0, 0, 0);
                        if (lane == 0) p.applied[f] = 0;
                    }
                }
            } else {
                equalize_uv_chunk(p, src, dst, c, tid - 32, kThreads - 32);
            }
            __syncthreads();
            if (!s_flag) break;  // dependency wait timed out (cannot happen; see kSpinCycles)
            lane_table_fill_from_lut(smem, s_lut);
            __syncthreads();
            const PlaneChunk pc = plane_chunk(p.flat, c, p.y_bytes, p.y_chunk, p.h, p.y_rows_chunk);
            if (p.flat) {
                const size_t n = (size_t)(pc.b1 - pc.b0), n1 = (n - n / 4) & ~(size_t)4095;
                lut_span(src + pc.b0, dst + pc.b0, n1, tid, kThreads, lane_base);
                q.prefetch();
                lut_span(src + pc.b0 + n1, dst + pc.b0 + n1, n - n1, tid, kThreads, lane_base);
            } else {
                for (int r = pc.r0 + warp; r < pc.r1; r += kWarps)
                    lut_span(src + (size_t)r * p.stride, dst + (size_t)r * p.stride, (size_t)p.w, lane, 32, lane_base);
            }
        }
        q.advance();
    }
    q.finish();
}

// ---- Appendix B synthetic frames, generated on the device (bench / test utility) -------------------------
constexpr int kSynthThreads = 256;
__global__ void __launch_bounds__(kSynthThreads) synth_nv12_kernel(uint8_t* out, unsigned long long pitch, int w, int h,
                                                              int stride, uint32_t seed, uint32_t first_frame) {
    const int f = blockIdx.y;
    uint8_t* base = out + (unsigned long long)f * pitch;
    const uint32_t frame = first_frame + (uint32_t)f;
    const int rows = h + h / 2;
    const int bw = max(w / 16, 1), bh = max(h / 9, 1);
    const long long total = (long long)rows * w;
    for (long long i = (long long)blockIdx.x * kSynthThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kSynthThreads) {
        const int r = (int)(i / w), c = (int)(i - (long long)r * w);
        uint8_t v;
        if (r < h) {
            const uint32_t idx = (uint32_t)r * (uint32_t)w + (uint32_t)c;
            const uint32_t k = fmix32(idx * 0x9E3779B1u + seed * 0x85EBCA77u + frame * 0xC2B2AE3Du);
            int base_v = 48 + (c * 128) / w + (r * 48) / h;
            int noise = (int)(k & 63u) - 32;
            if (((c / bw) + (r / bh)) % 5 == 0) { base_v = 200; noise = (int)(k & 3u); }
            v = (uint8_t)min(max(base_v + noise, 0), 255);
        } else {
            const uint32_t j = (uint32_t)(r - h) * (uint32_t)w + (uint32_t)c;
            const uint32_t k = fmix32(j * 0x9E3779B1u + seed + 0x01234567u + frame * 0xC2B2AE3Du);
            v = (uint8_t)(128 + (int)(k & 31u) - 16);
        }
        base[(size_t)r * stride + c] = v;
    }
}

__global__ void __launch_bounds__(kSynthThreads) synth_bgr_kernel(uint8_t* out, unsigned long long pitch, int w, int h,
                                                             int stride, uint32_t first_frame) {
    const int f = blockIdx.y;
    uint8_t* base = out + (unsigned long long)f * pitch;
    const uint32_t frame = first_frame + (uint32_t)f;
    const int bw = max(w / 16, 1), bh = max(h / 9, 1);
    const long long total = (long long)h * w * 3;
    for (long long i = (long long)blockIdx.x * kSynthThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kSynthThreads) {
        const long long px = i / 3;
        const int ch = (int)(i - px * 3);
        const int r = (int)(px / w), c = (int)(px - (long long)r * w);
        const uint32_t seed = 3026u + 1000u * (uint32_t)ch;
        const uint32_t idx = (uint32_t)r * (uint32_t)w + (uint32_t)c;
        const uint32_t k = fmix32(idx * 0x9E3779B1u + seed * 0x85EBCA77u + frame * 0xC2B2AE3Du);
        int base_v = 48 + (c * 128) / w + (r * 48) / h;
        int noise = (int)(k & 63u) - 32;
        if (((c / bw) + (r / bh)) % 5 == 0) { base_v = 200; noise = (int)(k & 3u); }
        base[(size_t)r * stride + (size_t)c * 3 + ch] = (uint8_t)min(max(base_v + noise, 0), 255);
    }
}

}  // namespace nv12eq
