// clahe.cuh -- CLAHE on the Y plane of a batch of NV12 frames, one launch per batch.
//
// Replaces cv::createCLAHE(clip, Size(tx,ty))->apply as called at clahevideo.cpp:184-195 together
// with the NV12 rebuild around it (clahevideo.cpp:200-201).
//
// Stages (all inside one kernel, scheduled with the same ticket-lag scheme as equalize.cuh):
//   tile item  (frame g, tile t): 256-bin histogram of the tile in smem hist[256][32] (conflict-free lane columns),
//              then one warp clips at clipLimit, redistributes the excess exactly as OpenCV does (redistBatch to
//              every bin, then +1 to every residualStep-th bin while residual lasts), scans, and writes the tile's
//              256-byte LUT (cvRound(sum * lutScale), fp32, round-half-even).  The padding path
//              (copyMakeBorder BORDER_REFLECT_101 when the grid does not divide the image) is an index reflection
//              in the tile reader; the padded image is never materialised.
//   cell item  (frame f = g - lag, interpolation cell (i, j)): a cell is a rectangle of pixels that blend the same
//              four tile LUTs.  The CTA packs those four LUTs into table[v][16] = {bf16 L11,L12,L21,L22} (exact:
//              0..255 fit bf16 and widening bf16->fp32 is a shift; 16 replicas make a half-warp 8-byte gather
//              conflict-free), then every pixel does ONE shared gather and OpenCV's blend op for op in unfused fp32
//              (products two pixels at a time with FMUL2, sums as scalar FADD so that nothing is contracted):
//                  res = (L11*xa1 + L12*xa)*ya1 + (L21*xa1 + L22*xa)*ya ;  dst = saturate(cvRound(res))
//   uv item    (frame f, chunk): chroma passthrough / 128 fill.
//
// Roofline: HBM, 3*W*H algorithmic bytes per frame (tile LUTs are 16 KB per frame).  Secondary limiters: shared
// atomics (tile items) and instruction issue (cell items: ~18 instructions per pixel).
#pragma once
#include "common.cuh"

namespace nv12eq {

constexpr int kMaxCells = 4096;  // per axis (tiles + 1); plenty
#ifndef NV12EQ_PREFETCH_ROUNDS
#define NV12EQ_PREFETCH_ROUNDS 10
#endif
constexpr int kPrefetchRounds = NV12EQ_PREFETCH_ROUNDS;  // L2 prefetch distance of the row-strided walkers

struct ClaheParams {
    const uint8_t* in;
    uint8_t* out;
    unsigned long long pitch;
    int n_frames;
    int w, h, stride;
    int flat;
    int uv_mode;
    int tx, ty;        // tile grid
    int tw, th;        // tile size in the (virtually) padded image
    int padded;        // grid does not divide the image: tiles read through reflect101
    int clip_limit;    // integer clip limit, 0 = no clipping
    float lut_scale;   // 255.f / (tw*th)
    float inv_tw, inv_th;
    int nxc, nyc;      // interpolation cells per axis
    const int4* xcells;  // [nxc] {x0, x1, tx1, tx2}
    const int4* ycells;  // [nyc] {y0, y1, ty1, ty2}
    int uv_chunks;     // chroma items per frame (0 when nothing to do)
    unsigned long long uv_bytes, uv_chunk;  // flat
    int uv_rows_chunk;                      // strided
    int lag;
    uint8_t* luts;         // [n_frames][tx*ty][256]
    uint32_t* tiles_done;  // [n_frames] self-cleaned
    uint32_t* ticket;      // [1] self-cleaned
    uint32_t* status;      // [1]
    unsigned long long* trace;  // optional [items][4] (developer tool)
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// One warp: tile histogram (shared, 256 ints) -> clip -> redistribute -> scan -> LUT bytes (global).
__device__ __forceinline__ void clahe_tile_lut_warp(const uint32_t* __restrict__ s_bins, int clip_limit, float lut_scale,
                                                    uint8_t* __restrict__ glut, int lane) {
    int h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = (int)s_bins[lane * 8 + j];
    if (clip_limit > 0) {
        int clipped = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (h[j] > clip_limit) { clipped += h[j] - clip_limit; h[j] = clip_limit; }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) clipped += __shfl_xor_sync(0xffffffffu, clipped, d);
        const int batch = clipped / 256;
        const int residual = clipped - batch * 256;
        int step = 1;
        if (residual != 0) step = max(256 / residual, 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = lane * 8 + j;
            h[j] += batch;
            // for (i = 0; i < 256 && residual > 0; i += step, residual--) h[i]++
            if (residual != 0 && (i % step) == 0 && (i / step) < residual) h[j] += 1;
        }
    }
    int lsum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) lsum += h[j];
    int run = (int)warp_incl_scan((uint32_t)lsum, lane) - lsum;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        run += h[j];
        const uint32_t v = round_sat_u8(__fmul_rn(__int2float_rn(run), lut_scale));
        if (j < 4) lo |= v << (8 * j); else hi |= v << (8 * (j - 4));
    }
    reinterpret_cast<uint2*>(glut)[lane] = make_uint2(lo, hi);
}

// bf16 halves of a word -> exact floats
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// One pixel of the blend (general path); e = {L11|L12<<16, L21|L22<<16} as bf16 pairs.  Returns the float whose LOW
// BYTE is the result: res is a convex-ish combination of values in [0,255], so 0 <= res < 255.5 (the weights sum to 1
// within a few ulp); adding 1.5*2^23 performs cvRound's round-half-to-even in the FADD and leaves the integer 0..255
// in the low mantissa byte -- saturate_cast is the identity here, so no clamp instructions are needed.
__device__ __forceinline__ uint32_t clahe_blend_bits(uint2 e, float xa, float xa1, float ya, float ya1) {
    const float top = __fadd_rn(__fmul_rn(bf16_lo(e.x), xa1), __fmul_rn(bf16_hi(e.x), xa));
    const float bot = __fadd_rn(__fmul_rn(bf16_lo(e.y), xa1), __fmul_rn(bf16_hi(e.y), xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    return __float_as_uint(__fadd_rn(res, 12582912.0f));
}
// Two horizontally adjacent pixels at once: the six products per pixel are issued as FMUL2 (two pixels per
// instruction), the three sums per pixel as scalar FADD, the rounding add as FADD2.  Same operations, same order,
// same rounding as the scalar form above.
__device__ __forceinline__ void clahe_blend_pair(uint2 e0, uint2 e1, uint64_t xa_p, uint64_t xa1_p, uint64_t ya_p, uint64_t ya1_p,
                                                 uint32_t& o0, uint32_t& o1) {
    const uint64_t a = pack_f2(bf16_lo(e0.x), bf16_lo(e1.x));  // L11 of both pixels
    const uint64_t b = pack_f2(bf16_hi(e0.x), bf16_hi(e1.x));  // L12
    const uint64_t c = pack_f2(bf16_lo(e0.y), bf16_lo(e1.y));  // L21
    const uint64_t d = pack_f2(bf16_hi(e0.y), bf16_hi(e1.y));  // L22
    float p0, p1, q0, q1;
    unpack_f2(mul_f2(a, xa1_p), p0, p1);
    unpack_f2(mul_f2(b, xa_p), q0, q1);
    const uint64_t top = pack_f2(__fadd_rn(p0, q0), __fadd_rn(p1, q1));
    unpack_f2(mul_f2(c, xa1_p), p0, p1);
    unpack_f2(mul_f2(d, xa_p), q0, q1);
    const uint64_t bot = pack_f2(__fadd_rn(p0, q0), __fadd_rn(p1, q1));
    unpack_f2(mul_f2(top, ya1_p), p0, p1);
    unpack_f2(mul_f2(bot, ya_p), q0, q1);
    const uint64_t res = add_f2(pack_f2(__fadd_rn(p0, q0), __fadd_rn(p1, q1)), pack_f2(12582912.0f, 12582912.0f));
    unpack_u2(res, o0, o1);
}
// low bytes of four words -> one packed word (3 PRMT)
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__device__ __forceinline__ void axis_weight(int pos, float inv, float& a, float& a1) {
    const float f = __fsub_rn(__fmul_rn((float)pos, inv), 0.5f);
    const float t1 = floorf(f);
    a = __fsub_rn(f, t1);
    a1 = __fsub_rn(1.0f, a);
}
// keeps a value in its register: stops the compiler from re-deriving xa1 = 1 - xa inside the pixel loop
__device__ __forceinline__ void pin_register(uint64_t& v) { asm volatile("" : "+l"(v)); }
__device__ __forceinline__ void pin_register(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin_register(uint32_t& v) { asm volatile("" : "+r"(v)); }

template <int MIN_CTAS>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) clahe_kernel(const ClaheParams p) {
    extern __shared__ __align__(16) uint32_t smem[];  // 32 KB: hist[256][32] for tile items, table[256][16] uint2 for cells
    __shared__ uint32_t s_bins[256];
    __shared__ uint32_t s_ticket[2];
    __shared__ int s_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.tx * p.ty;
    const int I = p.nxc * p.nyc;
    const int U = p.uv_chunks;
    const int per_slot = T + I + U;
    const uint32_t total_items = (uint32_t)(p.n_frames + p.lag) * (uint32_t)per_slot;
    const uint32_t smem_base = smem_u32(smem);

    TicketQueue q{p.ticket, s_ticket, 0u, 0u};
    q.start();
    for (;;) {
        const uint32_t item = q.current();
        if (item >= total_items) break;
        q.prefetch();
        const int g = (int)(item / (uint32_t)per_slot);
        const int r = (int)(item % (uint32_t)per_slot);
        const int f = g - p.lag;
        const ItemTrace tr_{p.trace};
        tr_.mark(item, 0);
        tr_.mark(item, 1);
        tr_.kind(item, r < T ? 1u : (r < T + I ? 2u : 3u));

        if (r < T) {
            if (g < p.n_frames) {
                // ------------------------- tile item: histogram -> clip -> LUT -------------------------
                const uint8_t* y = p.in + (unsigned long long)g * p.pitch;
                const int tyi = r / p.tx, txi = r - tyi * p.tx;
                const int x0 = txi * p.tw, y0 = tyi * p.th;
                const uint32_t lane_base = smem_base + lane * 4;
                lane_table_zero(smem);
                __syncthreads();
                const bool vec_ok = !p.padded && (p.tw & 15) == 0 && (p.stride & 15) == 0 &&
                                    (((uintptr_t)y + (uintptr_t)x0) & 15) == 0 && (p.tw >> 4) <= kThreads;
                if (vec_ok) {
                    // threads form a (rows_per_pass x vectors_per_row) grid over the tile; loads are software pipelined
                    const int vpr = p.tw >> 4;
                    const int rpp = kThreads / vpr;
                    const int tr = tid / vpr, tc = tid - tr * vpr;
                    if (tr < rpp) {
                        const uint64_t keep = l2_policy_evict_last();
                        const uint8_t* col = y + (size_t)y0 * p.stride + x0 + tc * 16;
                        const size_t rstep = (size_t)rpp * p.stride;
                        int row = tr;
                        const uint8_t* ptr = col + (size_t)row * p.stride;
                        bool have = row + rpp < p.th;  // a full round of 2 rows
                        int4 c0, c1;
                        if (have) {
                            c0 = ldg128_hint(ptr, keep);
                            c1 = ldg128_hint(ptr + rstep, keep);
                        }
                        while (have) {
                            const int rn = row + 2 * rpp;
                            const uint8_t* pn = ptr + 2 * rstep;
                            const bool more = rn + rpp < p.th;
                            int4 n0, n1;
                            if (more) {
                                n0 = ldg128_hint(pn, keep);
                                n1 = ldg128_hint(pn + rstep, keep);
                            }
                            hist_vec(c0, lane_base);
                            hist_vec(c1, lane_base);
                            if (more) { c0 = n0; c1 = n1; }
                            row = rn; ptr = pn; have = more;
                        }
                        for (; row < p.th; row += rpp, ptr += rstep) hist_vec(ldg128_hint(ptr, keep), lane_base);
                    }
                } else {
                    // general path: one warp per tile row, byte spans inside the image, reflected reads outside
                    for (int row = warp; row < p.th; row += kWarps) {
                        const uint8_t* src_row = y + (size_t)reflect101(y0 + row, p.h) * p.stride;
                        const int xin = min(x0 + p.tw, p.w);  // end of the in-image part
                        if (x0 < xin) hist_span(src_row + x0, (size_t)(xin - x0), lane, 32, lane_base);
                        for (int x = max(x0, p.w) + lane; x < x0 + p.tw; x += 32)
                            hist_byte(src_row[reflect101(x, p.w)], lane_base);
                    }
                }
                __syncthreads();
                if (tid < 256) s_bins[tid] = lane_table_row_sum(smem, tid);
                __syncthreads();
                if (warp == 0) {
                    clahe_tile_lut_warp(s_bins, p.clip_limit, p.lut_scale, p.luts + ((size_t)g * T + r) * 256, lane);
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicAdd(p.tiles_done + g, 1u);
                }
            }
        } else if (f >= 0) {
            const uint8_t* src = p.in + (unsigned long long)f * p.pitch;
            uint8_t* dst = p.out + (unsigned long long)f * p.pitch;
            if (r < T + I) {
                // ------------------------- cell item: blend four tile LUTs -------------------------
                if (tid == 0) {
                    bool ok = true;
                    if (ld_acquire_u32(p.tiles_done + f) < (uint32_t)T) {
                        const long long t0 = clock64();
                        unsigned ns = 64;
                        while (ld_acquire_u32(p.tiles_done + f) < (uint32_t)T) {
                            __nanosleep(ns);
                            if (ns < 2048) ns <<= 1;
                            if (clock64() - t0 > kSpinCycles) { ok = false; break; }
                        }
                    }
                    if (!ok) atomicExch(p.status, 1u);
                    s_flag = ok;
                }
                __syncthreads();
                if (!s_flag) break;
                tr_.mark(item, 1);
                const int ci = r - T;
                const int cy = ci / p.nxc, cx = ci - cy * p.nxc;
                const int4 xc = p.xcells[cx], yc = p.ycells[cy];
                // pack the four LUTs: table[v][rep] (uint2), 16 replicas so a half-warp 8-byte gather is conflict-free
                {
                    constexpr int kShare = kThreads / 256, kPer = 16 / kShare;
                    const uint8_t* L = p.luts + (size_t)f * T * 256;
                    const int v = tid & 255, part = tid >> 8;
                    const uint32_t l11 = __ldcg(L + (size_t)(yc.z * p.tx + xc.z) * 256 + v);
                    const uint32_t l12 = __ldcg(L + (size_t)(yc.z * p.tx + xc.w) * 256 + v);
                    const uint32_t l21 = __ldcg(L + (size_t)(yc.w * p.tx + xc.z) * 256 + v);
                    const uint32_t l22 = __ldcg(L + (size_t)(yc.w * p.tx + xc.w) * 256 + v);
                    uint2 e;
                    e.x = (__float_as_uint((float)l11) >> 16) | (__float_as_uint((float)l12) & 0xffff0000u);
                    e.y = (__float_as_uint((float)l21) >> 16) | (__float_as_uint((float)l22) & 0xffff0000u);
                    uint2* row = reinterpret_cast<uint2*>(smem) + v * 16;
#pragma unroll
                    for (int j = 0; j < kPer; ++j) row[(part * kPer + j + v) & 15] = e;
                }
                __syncthreads();
                uint32_t rep_base = smem_base + (uint32_t)(lane & 15) * 8u;
                pin_register(rep_base);
                const int cw = xc.y - xc.x;                 // cell width in pixels
                const int gpr = (cw + 7) >> 3;              // 8-pixel groups per row
                const bool fast = ((xc.x & 7) == 0) && ((p.stride & 7) == 0) && ((((uintptr_t)src | (uintptr_t)dst) & 7) == 0) &&
                                  gpr <= kThreads;
                if (fast) {
                    const int rpp = kThreads / gpr;
                    const int tr = tid / gpr, tc = tid - tr * gpr;
                    const int xg = xc.x + tc * 8;
                    if (tr < rpp) {
                        const int npx = min(8, xc.y - xg);  // < 8 only in the last group of a ragged cell
                        uint64_t xa_p[4], xa1_p[4];  // x weights of the pixel pairs (0,1) (2,3) (4,5) (6,7)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float a0, b0, a1, b1;
                            axis_weight(xg + 2 * k, p.inv_tw, a0, b0);
                            axis_weight(xg + 2 * k + 1, p.inv_tw, a1, b1);
                            pin_register(b0);  // xa1 = 1 - xa must stay in a register, not be re-derived per row
                            pin_register(b1);
                            xa_p[k] = pack_f2(a0, a1);
                            xa1_p[k] = pack_f2(b0, b1);
                            pin_register(xa1_p[k]);
                        }
                        const uint64_t once = l2_policy_evict_first();
                        int yrow = yc.x + tr;
                        const size_t rstep = (size_t)rpp * p.stride;
                        const uint8_t* sp = src + (size_t)yrow * p.stride + xg;
                        uint8_t* dp = dst + (size_t)yrow * p.stride + xg;
                        // three rows in flight per thread
                        uint2 cur = make_uint2(0, 0), n1 = make_uint2(0, 0), n2 = make_uint2(0, 0);
                        if (yrow < yc.y) cur = ldg64_hint(sp, once);
                        if (yrow + rpp < yc.y) n1 = ldg64_hint(sp + rstep, once);
                        if (yrow + 2 * rpp < yc.y) n2 = ldg64_hint(sp + 2 * rstep, once);
                        const uint8_t* sp3 = sp + 3 * rstep;
                        for (; yrow < yc.y; yrow += rpp, sp3 += rstep, dp += rstep) {
                            uint2 n3 = make_uint2(0, 0);
                            if (yrow + 3 * rpp < yc.y) n3 = ldg64_hint(sp3, once);
                            float ya, ya1;
                            axis_weight(yrow, p.inv_th, ya, ya1);
                            const uint64_t ya_p = pack_f2(ya, ya), ya1_p = pack_f2(ya1, ya1);
                            uint32_t o[8];
                            clahe_blend_pair(lds_u64(rep_base + (byte_of<0>(cur.x) << 7)), lds_u64(rep_base + (byte_of<1>(cur.x) << 7)),
                                             xa_p[0], xa1_p[0], ya_p, ya1_p, o[0], o[1]);
                            clahe_blend_pair(lds_u64(rep_base + (byte_of<2>(cur.x) << 7)), lds_u64(rep_base + (byte_of<3>(cur.x) << 7)),
                                             xa_p[1], xa1_p[1], ya_p, ya1_p, o[2], o[3]);
                            clahe_blend_pair(lds_u64(rep_base + (byte_of<0>(cur.y) << 7)), lds_u64(rep_base + (byte_of<1>(cur.y) << 7)),
                                             xa_p[2], xa1_p[2], ya_p, ya1_p, o[4], o[5]);
                            clahe_blend_pair(lds_u64(rep_base + (byte_of<2>(cur.y) << 7)), lds_u64(rep_base + (byte_of<3>(cur.y) << 7)),
                                             xa_p[3], xa1_p[3], ya_p, ya1_p, o[6], o[7]);
                            if (npx == 8) {
                                stg64_hint(dp, make_uint2(pack_low_bytes(o[0], o[1], o[2], o[3]), pack_low_bytes(o[4], o[5], o[6], o[7])),
                                           once);
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if (k < npx) dp[k] = (uint8_t)o[k];
                            }
                            cur = n1; n1 = n2; n2 = n3;
                        }
                    }
                } else {
                    // general path: one pixel per thread per step
                    const int ch = yc.y - yc.x;
                    const long long npix = (long long)cw * ch;
                    for (long long i = tid; i < npix; i += kThreads) {
                        const int ry = (int)(i / cw), rx = (int)(i - (long long)ry * cw);
                        const int x = xc.x + rx, yy = yc.x + ry;
                        float xa, xa1, ya, ya1;
                        axis_weight(x, p.inv_tw, xa, xa1);
                        axis_weight(yy, p.inv_th, ya, ya1);
                        const uint32_t v = src[(size_t)yy * p.stride + x];
                        dst[(size_t)yy * p.stride + x] = (uint8_t)clahe_blend_bits(lds_u64(rep_base + (v << 7)), xa, xa1, ya, ya1);
                    }
                }
            } else {
                // ------------------------- uv item -------------------------
                const int c = r - T - I;
                const bool copy_uv = p.uv_mode == UV_COPY && src != dst;
                const size_t uv_off = (size_t)p.stride * p.h;
                if (p.flat) {
                    const unsigned long long b0 = min((unsigned long long)c * p.uv_chunk, p.uv_bytes);
                    const unsigned long long b1 = min(b0 + p.uv_chunk, p.uv_bytes);
                    if (copy_uv) copy_span(src + uv_off + b0, dst + uv_off + b0, (size_t)(b1 - b0), tid, kThreads);
                    else if (p.uv_mode == UV_GRAY128) fill_span(dst + uv_off + b0, (size_t)(b1 - b0), tid, kThreads, 128);
                } else {
                    const int rows = p.h / 2;
                    const int r0 = min(c * p.uv_rows_chunk, rows), r1 = min(r0 + p.uv_rows_chunk, rows);
                    for (int row = r0 + warp; row < r1; row += kWarps) {
                        const size_t off = uv_off + (size_t)row * p.stride;
                        if (copy_uv) copy_span(src + off, dst + off, (size_t)p.w, lane, 32);
                        else if (p.uv_mode == UV_GRAY128) fill_span(dst + off, (size_t)p.w, lane, 32, 128);
                    }
                }
            }
        }
        tr_.mark(item, 2);
        q.advance();
    }
    // the last CTA out returns the per-frame counters to zero for the next launch
    if (q.finish())
        for (int i = tid; i < p.n_frames; i += kThreads) p.tiles_done[i] = 0;
}

}  // namespace nv12eq
