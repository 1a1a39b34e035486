// clahe16.cuh -- CLAHE on CV_16UC1 planes (P010 luma and other 16-bit content): SURVEY.md section 8f rank 3.
//
// OpenCV's CLAHE accepts 16-bit input with histSize = 65536; the reference never feeds it 16-bit data, so this is a
// widening row.  65536-bin tile histograms (256 KB) and tile LUTs (128 KB) do not fit shared memory as they are; the
// working set of one plane (histograms, LUTs, cell tables: ~66 MB for an 8x8 grid) is kept L2-resident by processing
// few planes per pass.
//   clahe16_hist_kernel        one CTA per (tile strip, tile, plane), 1024 threads, one CTA per SM: the strip's histogram is
//                              built in 128 KB of shared memory as 65536 packed 16-bit counters (a strip has < 65536
//                              pixels, so none overflows; the word index is XOR-swizzled so that P010's multiples of 64
//                              do not all land in bank 0) and only the non-zero counters are added to the global
//                              histogram (reflect-101 padding by index reflection, as in the 8-bit kernel)
//   clahe16_lut_kernel         one cluster of 8 CTAs per tile, 8192 bins per CTA, one read of the histogram: clip, exchange
//                              the partial sums through distributed shared memory, redistribute the excess exactly as
//                              OpenCV does (redistBatch to every bin, +1 to every residualStep-th bin while the residual
//                              lasts; the number of those bins in front of a part has a closed form), one block scan,
//                              lut = saturate_cast<ushort>(cvRound(sum * lutScale)); returns the histogram to zero
//   clahe16_cell_table_kernel  per interpolation cell (the region between four tile centres): table[v] =
//                              {L11, L12, L21, L22}[v], 8 bytes, so that the blend needs ONE gather per pixel
//   clahe16_interp_kernel      per pixel: one 8-byte gather and OpenCV's blend op for op in unfused fp32 (same weights and
//                              rounding as the 8-bit path)
// Bound: the gather (one distinct cache line per pixel through L1TEX) and the shared-memory atomics, not HBM;
// algorithmic bytes are 4*W*H per plane.
#pragma once
#include <cooperative_groups.h>
#include "clahe.cuh"

namespace nv12eq {

constexpr int kBins16 = 65536;
constexpr int kC16Threads = 256;
constexpr int kC16LutThreads = 512;
constexpr int kC16BinsPerThread = 16;
constexpr int kC16HistThreads = 1024;
constexpr int kC16HistSmemBytes = kBins16 * 2;           // packed 16-bit counters
constexpr int kC16StripPixels = 65535;                   // a 16-bit counter cannot overflow within one strip
constexpr int kC16Parts = kBins16 / (kC16LutThreads * kC16BinsPerThread);  // 8 parts of 8192 bins: one cluster of 8 CTAs per tile
constexpr int kC16PartBins = kBins16 / kC16Parts;

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
    uint16_t* luts;             // [n_planes][tx*ty][65536], compact: entry k is the LUT value of k << z
    uint2* cells;               // [n_planes][(ty+1)*(tx+1)][65536]: {L11 | L12 << 16, L21 | L22 << 16} at index v >> z
    uint32_t* ormask;           // [n_planes]: OR of all pixel values of the plane (zero on entry to the histogram kernel)
    int strips, rows_strip;     // histogram kernel: row strips per tile, rows per strip
};

// word of the packed shared histogram that holds bins 2*w and 2*w + 1 (an involution: it is its own inverse)
__device__ __forceinline__ uint32_t c16_swizzle(uint32_t w) { return w ^ ((w >> 5) & 31u) ^ ((w >> 10) & 31u); }
__device__ __forceinline__ void c16_count(uint32_t* cnt, uint32_t v) {
    atomicAdd(cnt + c16_swizzle(v >> 1), (v & 1u) * 0xffffu + 1u);   // +1 in the low or in the high half
}
__device__ __forceinline__ void c16_count2(uint32_t* cnt, uint32_t w) {
    c16_count(cnt, w & 0xffffu);
    c16_count(cnt, w >> 16);
}
// Number of low bits that are zero in every pixel of the plane (P010: 6, 12-bit content in 16-bit words: 4, full range: 0).
// Only the LUT entries v = k << z can ever be looked up, so the cell tables are built and indexed at v >> z: for 10-bit
// video that is 8 KB per cell instead of 512 KB, and neighbouring grey levels share cache lines.
__device__ __forceinline__ int c16_zero_bits(uint32_t ormask) {
    const uint32_t m = ormask & 0xffffu;
    return m ? __ffs((int)m) - 1 : 16;
}

__global__ void __launch_bounds__(kC16HistThreads, 1) clahe16_hist_kernel(const Clahe16Params p) {
    uint32_t* cnt = nv12eq_smem_rows;   // 32768 words
    const int T = p.tx * p.ty;
    const int f = blockIdx.z, t = blockIdx.y, strip = blockIdx.x;
    const int tyi = t / p.tx, txi = t - tyi * p.tx;
    const int x0 = txi * p.tw, y0 = tyi * p.th;
    const int r0 = strip * p.rows_strip, r1 = min(r0 + p.rows_strip, p.th);
    if (r0 >= r1) return;
    const uint16_t* src = p.in + (unsigned long long)f * p.pitch;
    uint32_t* hist = p.hist + ((size_t)f * T + t) * kBins16;
    __shared__ uint32_t s_seen;
    if (threadIdx.x == 0) s_seen = 0;
    uint32_t seen = 0;
    for (int i = threadIdx.x; i < kBins16 / 8; i += kC16HistThreads) reinterpret_cast<uint4*>(cnt)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const bool vec = x0 + p.tw <= p.w && (p.tw & 7) == 0 && (p.stride & 7) == 0 && (((uintptr_t)src + 2 * (uintptr_t)x0) & 15) == 0;
    if (vec) {
        const int vpr = p.tw >> 3, n = (r1 - r0) * vpr;
        // vector i of the strip is vector c of row r; the pair advances by one CTA stride without a division
        const int dr = kC16HistThreads / vpr, dc = kC16HistThreads - dr * vpr;
        int r = (int)threadIdx.x / vpr, c = (int)threadIdx.x - r * vpr;
        const uint16_t* tile = src + x0;
        auto load = [&]() {
            const int y = y0 + r0 + r;
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(tile + (size_t)(y < p.h ? y : reflect101(y, p.h)) * p.stride) + c);
            c += dc; r += dr;
            if (c >= vpr) { c -= vpr; ++r; }
            return q;
        };
        // two vectors in flight per thread: the loads of the next round overlap the shared atomics of this one
        int i = threadIdx.x;
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
        if (i < n) q0 = load();
        if (i + kC16HistThreads < n) q1 = load();
        while (i < n) {
            const uint4 q = q0;
            q0 = q1;
            if (i + 2 * kC16HistThreads < n) q1 = load();
            seen |= q.x | q.y | q.z | q.w;
            c16_count2(cnt, q.x); c16_count2(cnt, q.y); c16_count2(cnt, q.z); c16_count2(cnt, q.w);
            i += kC16HistThreads;
        }
    } else {
        const int n = (r1 - r0) * p.tw;
        for (int i = threadIdx.x; i < n; i += kC16HistThreads) {
            const int r = i / p.tw, c = i - r * p.tw;
            const uint32_t v = src[(size_t)reflect101(y0 + r0 + r, p.h) * p.stride + reflect101(x0 + c, p.w)];
            seen |= v;
            c16_count(cnt, v);
        }
    }
    seen = __reduce_or_sync(0xffffffffu, (seen | (seen >> 16)) & 0xffffu);
    if ((threadIdx.x & 31) == 0 && (seen & ~s_seen) != 0) atomicOr(&s_seen, seen);
    __syncthreads();
    if (threadIdx.x == 0 && s_seen != 0) atomicOr(p.ormask + f, s_seen);
    for (int i = threadIdx.x; i < kBins16 / 2; i += kC16HistThreads) {
        const uint32_t c = cnt[i];
        if (c) {
            const uint32_t b = c16_swizzle((uint32_t)i) * 2u;
            if (c & 0xffffu) atomicAdd(hist + b, c & 0xffffu);
            if (c >> 16) atomicAdd(hist + b + 1, c >> 16);
        }
    }
}

// Block-wide inclusive scan of one int per thread (up to 1024 threads); *total = sum over the block.
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
        const int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
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

// Sum of an int pair over the block (up to 1024 threads); valid in every thread.
__device__ __forceinline__ int2 block_sum2(int a, int b, int* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    if (lane == 0) { s_warp[warp] = a; s_warp[32 + warp] = b; }
    __syncthreads();
    const bool live = lane < (int)(blockDim.x >> 5);
    a = live ? s_warp[lane] : 0; b = live ? s_warp[32 + lane] : 0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    return make_int2(a, b);
}

// grid (kC16Parts, tiles, planes) in clusters of kC16Parts CTAs: one cluster per tile histogram.  Only the bins v = k << z
// (z = c16_zero_bits of the plane, k < 65536 >> z) can be non-zero and only their LUT entries can be looked up, so the
// kernel works on those "compact" bins: CTA `part` owns compact bins [8192 part, 8192 (part + 1)), thread t 16 of them.
// The parts exchange their sums through distributed shared memory, so the histogram is read once.  OpenCV's
// redistribution (redistBatch to every one of the 65536 bins, +1 to bins 0, step, 2 step, ... while the residual lasts)
// has a closed form for the cumulative sum at bin v:
//     cum(v) = sum_{u <= v} min(h[u], clip)  +  redistBatch * (v + 1)  +  min(residual, v / step + 1)   [last term if residual > 0]
// The LUT is written compact: entry k of a tile's table holds the value for v = k << z.
__global__ void __cluster_dims__(kC16Parts, 1, 1) __launch_bounds__(kC16LutThreads, 3) clahe16_lut_kernel(const Clahe16Params p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int s_warp[64];
    __shared__ int2 s_part;     // (sum of min(h, clip), clipped excess) of this part; read by the whole cluster
    __shared__ int s_hdr[4];    // batch, residual, step, clipped histogram mass in front of this part
    const int T = p.tx * p.ty;
    const int part = blockIdx.x, t = blockIdx.y, f = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int z = c16_zero_bits(__ldg(p.ormask + f));
    const int nb = kBins16 >> z;
    constexpr int B = kC16BinsPerThread;
    // At most 8192 compact bins (13-bit content or less, e.g. P010): part 0 owns them all and runs on its own; the
    // decision is the same in every CTA of the cluster, so nobody waits for the CTAs that leave here.
    const bool solo = nb <= kC16PartBins;
    if (solo && part > 0) return;
    const int k0 = part * kC16PartBins + threadIdx.x * B;   // first compact bin of this thread
    uint32_t* hist = p.hist + ((size_t)f * T + t) * kBins16;
    uint16_t* lut = p.luts + ((size_t)f * T + t) * kBins16;
    int h[B];
    if (z == 0) {
#pragma unroll
        for (int j = 0; j < B / 4; ++j) {
            const uint4 r = reinterpret_cast<const uint4*>(hist + k0)[j];
            h[4 * j] = (int)r.x; h[4 * j + 1] = (int)r.y; h[4 * j + 2] = (int)r.z; h[4 * j + 3] = (int)r.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < B; ++j) h[j] = k0 + j < nb ? (int)hist[(size_t)(k0 + j) << z] : 0;
    }
    int kept = 0, excess = 0;   // a tile has fewer than 2^31 pixels: int sums are safe
#pragma unroll
    for (int j = 0; j < B; ++j) {
        if (p.clip_limit > 0) { excess += max(h[j] - p.clip_limit, 0); h[j] = min(h[j], p.clip_limit); }
        kept += h[j];
    }
    const int2 mine = block_sum2(kept, excess, s_warp);
    if (threadIdx.x == 0) s_part = mine;
    if (!solo) cluster.sync();   // release/acquire on s_part; no global store is pending yet, so the fence is cheap
    else __syncthreads();
    if (warp == 0) {
        int2 v = make_int2(0, 0);
        if (solo) { if (lane == 0) v = s_part; }
        else if (lane < kC16Parts) v = *cluster.map_shared_rank(&s_part, lane);
        int clipped = v.y;
        int before = lane < part ? v.x : 0;
#pragma unroll
        for (int d = kC16Parts / 2; d >= 1; d >>= 1) {
            clipped += __shfl_xor_sync(0xffffffffu, clipped, d);
            before += __shfl_xor_sync(0xffffffffu, before, d);
        }
        const int batch = clipped / kBins16;
        const int residual = clipped - batch * kBins16;
        if (lane == 0) { s_hdr[0] = batch; s_hdr[1] = residual; s_hdr[2] = residual != 0 ? max(kBins16 / residual, 1) : 1; s_hdr[3] = before; }
    }
    __syncthreads();
    // the remote reads of this CTA are done: arrive now, wait at the very end (s_part must outlive the other CTAs' reads);
    // relaxed, because nothing written after this point is read inside the cluster
    if (!solo) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    // histogram back to zero, ready for the next launch
    if (z == 0) {
#pragma unroll
        for (int j = 0; j < B / 4; ++j) reinterpret_cast<uint4*>(hist + k0)[j] = make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int j = 0; j < B; ++j)
            if (k0 + j < nb) hist[(size_t)(k0 + j) << z] = 0;
    }
    const int batch = s_hdr[0], residual = s_hdr[1], step = s_hdr[2];
    // floor(2^32 / step) + 1 in 32-bit arithmetic: (v * magic) >> 32 == v / step for v < 65536, 2 <= step <= 65536
    const uint32_t magic = step > 1 ? 0xffffffffu / (uint32_t)step + 1u + ((step & (step - 1)) == 0 ? 1u : 0u) : 0u;
    int part_total;
    int run = s_hdr[3] + block_incl_scan(kept, s_warp, &part_total) - kept;
    const bool whole = k0 + B <= nb;
#pragma unroll
    for (int j8 = 0; j8 < B; j8 += 8) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            run += h[j8 + j];
            const uint32_t v = (uint32_t)(k0 + j8 + j) << z;   // < 65536 for every bin that is stored
            const int vq = step > 1 ? (int)(((unsigned long long)v * magic) >> 32) : (int)v;
            const int cum = run + batch * (int)(v + 1u) + (residual != 0 ? min(residual, vq + 1) : 0);
            o[j] = (uint32_t)min(max(__float2int_rn(__fmul_rn(__int2float_rn(cum), p.lut_scale)), 0), 65535);
        }
        if (whole) {
            *reinterpret_cast<uint4*>(lut + k0 + j8) = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k0 + j8 + j < nb) lut[k0 + j8 + j] = (uint16_t)o[j];
        }
    }
    if (!solo) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

// grid (8, cells, planes), 256 threads x 4 table entries per round: the four tile LUTs a cell blends, interleaved per
// value.  Cell (cy, cx) lies between tile rows cy-1, cy and tile columns cx-1, cx (clamped to the grid).  Like the tile
// LUTs the table is compact: entry k holds value k << z (see c16_zero_bits), entries past 65536 >> z do not exist.
__global__ void __launch_bounds__(kC16Threads) clahe16_cell_table_kernel(const Clahe16Params p) {
    const int T = p.tx * p.ty, ncx = p.tx + 1;
    const int cell = blockIdx.y, f = blockIdx.z;
    const int z = c16_zero_bits(__ldg(p.ormask + f));
    const int nb = kBins16 >> z;
    const int cy = cell / ncx, cx = cell - cy * ncx;
    const int ty1 = max(cy - 1, 0), ty2 = min(cy, p.ty - 1), tx1 = max(cx - 1, 0), tx2 = min(cx, p.tx - 1);
    const uint16_t* luts = p.luts + (size_t)f * T * kBins16;
    const uint16_t* la = luts + (size_t)(ty1 * p.tx + tx1) * kBins16;
    const uint16_t* lb = luts + (size_t)(ty1 * p.tx + tx2) * kBins16;
    const uint16_t* lc = luts + (size_t)(ty2 * p.tx + tx1) * kBins16;
    const uint16_t* ld = luts + (size_t)(ty2 * p.tx + tx2) * kBins16;
    uint2* table = p.cells + ((size_t)f * (p.ty + 1) * ncx + cell) * kBins16;
    for (int v4 = blockIdx.x * kC16Threads + threadIdx.x; 4 * v4 < nb; v4 += gridDim.x * kC16Threads) {   // entries 4*v4 .. 4*v4+3
        uint2* dst = table + (size_t)v4 * 4;
        if (4 * v4 + 4 <= nb) {
            const uint2 a = reinterpret_cast<const uint2*>(la)[v4], b = reinterpret_cast<const uint2*>(lb)[v4];
            const uint2 c = reinterpret_cast<const uint2*>(lc)[v4], d = reinterpret_cast<const uint2*>(ld)[v4];
            // 0x5410: low halves of the two words packed, 0x7632: high halves
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(__byte_perm(a.x, b.x, 0x5410), __byte_perm(c.x, d.x, 0x5410),
                                                          __byte_perm(a.x, b.x, 0x7632), __byte_perm(c.x, d.x, 0x7632));
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(__byte_perm(a.y, b.y, 0x5410), __byte_perm(c.y, d.y, 0x5410),
                                                          __byte_perm(a.y, b.y, 0x7632), __byte_perm(c.y, d.y, 0x7632));
        } else {
            for (int k = 4 * v4; k < nb; ++k)   // fewer than four entries in all (z > 14)
                dst[k - 4 * v4] = make_uint2((uint32_t)la[k] | ((uint32_t)lb[k] << 16), (uint32_t)lc[k] | ((uint32_t)ld[k] << 16));
        }
    }
}

constexpr int kC16RowsPerCta = 16;   // rows per CTA of the blend kernel, in batches of kC16RowBatch per thread
constexpr int kC16RowBatch = 8;

// One column of up to 8 rows: all pixel loads first, then all table gathers, then the blends.  FULL: all 8 rows exist
// (no predicates in the unrolled code).
template <bool FULL>
__device__ __forceinline__ void c16_blend_rows(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, const uint2* __restrict__ cells,
                                               uint32_t stride, int nrows, int z, float xa, float xa1, const float2* s_yw,
                                               const uint32_t* s_row) {
    uint32_t v[kC16RowBatch];
#pragma unroll
    for (int k = 0; k < kC16RowBatch; ++k) v[k] = (FULL || k < nrows) ? src[(uint32_t)k * stride] : 0u;
    uint2 e[kC16RowBatch];
#pragma unroll
    for (int k = 0; k < kC16RowBatch; ++k) e[k] = __ldg(cells + (s_row[k] + (v[k] >> z)));
#pragma unroll
    for (int k = 0; k < kC16RowBatch; ++k) {
        if (FULL || k < nrows) {
            const float2 yw = s_yw[k];
            // u16 -> float without the conversion pipe: 0x4B00xxxx is 2^23 + x exactly
            const float l11 = __fsub_rn(__uint_as_float(__byte_perm(e[k].x, 0x4B000000u, 0x7610)), 8388608.0f);
            const float l12 = __fsub_rn(__uint_as_float(__byte_perm(e[k].x, 0x4B000000u, 0x7632)), 8388608.0f);
            const float l21 = __fsub_rn(__uint_as_float(__byte_perm(e[k].y, 0x4B000000u, 0x7610)), 8388608.0f);
            const float l22 = __fsub_rn(__uint_as_float(__byte_perm(e[k].y, 0x4B000000u, 0x7632)), 8388608.0f);
            const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
            const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
            const float res = __fadd_rn(__fmul_rn(top, yw.x), __fmul_rn(bot, yw.y));
            // 0 <= res < 65535.5 (a blend of values in [0, 65535] with weights in [0, 1] that sum to 1 up to rounding), so
            // adding 1.5 * 2^23 performs cvRound's round-half-to-even and leaves the integer in the low 16 mantissa bits;
            // saturate_cast is the identity
            dst[(uint32_t)k * stride] = (uint16_t)__float_as_uint(__fadd_rn(res, 12582912.0f));
        }
    }
}

// grid (ceil(w / 256), ceil(h / 16), planes): a thread owns one column of 16 rows, so the x weights are computed once; the
// weights and cell rows of the 16 rows are shared by the CTA.
__global__ void __launch_bounds__(kC16Threads) clahe16_interp_kernel(const Clahe16Params p) {
    __shared__ float2 s_yw[kC16RowsPerCta];   // (ya1, ya)
    __shared__ uint32_t s_row[kC16RowsPerCta];  // first entry of the row's cell row
    const int ncx = p.tx + 1;
    const int f = blockIdx.z;
    const int y0 = blockIdx.y * kC16RowsPerCta, y1 = min(y0 + kC16RowsPerCta, p.h);
    if (threadIdx.x < kC16RowsPerCta) {
        const int y = y0 + threadIdx.x;
        float ya, ya1;
        axis_weight(y, p.inv_th, ya, ya1);
        const int cy = (int)floorf(__fsub_rn(__fmul_rn((float)y, p.inv_th), 0.5f)) + 1;   // 0 .. ty
        s_yw[threadIdx.x] = make_float2(ya1, ya);
        s_row[threadIdx.x] = (uint32_t)(min(cy, p.ty) * ncx) * (uint32_t)kBins16;
    }
    __syncthreads();
    const int x = blockIdx.x * kC16Threads + threadIdx.x;
    if (x >= p.w) return;
    const int z = c16_zero_bits(__ldg(p.ormask + f));
    float xa, xa1;
    axis_weight(x, p.inv_tw, xa, xa1);
    const int cx = (int)floorf(__fsub_rn(__fmul_rn((float)x, p.inv_tw), 0.5f)) + 1;   // 0 .. tx
    const uint2* cells = p.cells + (size_t)f * (p.ty + 1) * ncx * kBins16 + (size_t)cx * kBins16;
    const uint16_t* src = p.in + (unsigned long long)f * p.pitch + (size_t)y0 * p.stride + x;
    uint16_t* dst = p.out + (unsigned long long)f * p.pitch + (size_t)y0 * p.stride + x;
    const bool full = y0 + kC16RowsPerCta <= p.h;
#pragma unroll
    for (int b = 0; b < kC16RowsPerCta; b += kC16RowBatch) {
        const size_t off = (size_t)b * p.stride;
        if (full) c16_blend_rows<true>(src + off, dst + off, cells, (uint32_t)p.stride, kC16RowBatch, z, xa, xa1, s_yw + b, s_row + b);
        else if (y0 + b < y1) c16_blend_rows<false>(src + off, dst + off, cells, (uint32_t)p.stride, y1 - y0 - b, z, xa, xa1, s_yw + b, s_row + b);
    }
}

}  // namespace nv12eq
