// clahe.cuh -- CLAHE on the Y plane of a batch of NV12 frames, one launch per batch.
//
// Replaces cv::createCLAHE(clip, Size(tx,ty))->apply as called at clahevideo.cpp:184-195 together
// with the NV12 rebuild around it (clahevideo.cpp:200-201).
//
// Stages (all inside one kernel, scheduled with the same ticket-lag scheme as equalize.cuh):
//   tile item  (frame g, tile t): tile rows staged through a per-thread cp.async ring, 256-bin histogram of the tile in smem
//              hist[256][32] (conflict-free lane columns), then one warp clips at clipLimit, redistributes the excess exactly as OpenCV does (redistBatch to
//              every bin, then +1 to every residualStep-th bin while residual lasts), scans, and writes the tile's
//              256-byte LUT (cvRound(sum * lutScale), fp32, round-half-even).  The padding path
//              (copyMakeBorder BORDER_REFLECT_101 when the grid does not divide the image) is an index reflection
//              in the tile reader; the padded image is never materialised.
//   cell item  (frame f = g - lag, interpolation cell (i, j)): a cell is a rectangle of pixels that blend the same
//              four tile LUTs.  The CTA packs those four LUTs into table[v][reps] = {bf16 L11 | L21, bf16 L12 | L22}
//              (exact: 0..255 fit bf16 and widening bf16->fp32 is a shift; 16 or 32 8-byte replicas make a half-warp
//              gather conflict-free), pixel rows arrive through a per-thread cp.async ring (16 or 8 pixels per thread
//              and row), then every pixel does ONE shared gather and OpenCV's blend op for op in
//              unfused fp32, the top and bottom row of the 2x2 LUT neighbourhood side by side in packed fp32:
//                  (top, bot) = (L11, L21)*xa1 + (L12, L22)*xa        FMUL2, FMUL2, FADD2
//                  res        = top*ya1 + bot*ya                      FMUL2, FADD
//                  dst        = saturate(cvRound(res))                FADD2 with 1.5*2^23 on a pixel pair, PRMT
//              Table rows are 128 bytes apart by default (PRMT + multiply-add per address); the 256-byte variant turns a
//              pixel byte into the row offset with ONE PRMT (byte 1 = pixel value, byte 0 = lane offset) but needs a 64 KB
//              table, i.e. fewer CTAs per SM, and measured slower (profiles/r01_clahe_notes.md).
//   uv item    (frame f, chunk): chroma passthrough / 128 fill.
//
// Roofline: HBM, 3*W*H algorithmic bytes per frame (tile LUTs are 16 KB per frame).  Secondary limiters: shared
// atomics (tile items) and instruction issue (cell items: ~15 instructions per pixel).
#pragma once
#include "common.cuh"

// Dynamic shared memory of clahe_kernel, declared at global scope so that its PTX name is unmangled: the kernel takes
// its address with `mov.u32 r, nv12eq_smem_rows`, a plain shared-window offset (cvta would add the CTA's cluster-window
// bits), which keeps ring and table addresses simple register + immediate forms.
extern __shared__ __align__(256) uint32_t nv12eq_smem_rows[];

namespace nv12eq {

constexpr int kMaxCells = 4096;  // per axis (tiles + 1); plenty
constexpr int kMaxCellRows = 512;               // rows per interpolation cell (host cuts longer runs)
// CTA shape of clahe_kernel (compile-time; the Makefile's EXTRA can override for experiments):
//   256 threads x 4 (or 3) CTAs/SM with 128-byte table rows (32 KB table) -- the default: several independent items per SM
//       fill the bubbles of the per-item phases (table build, LUT warp, barriers); measured 8.1 us vs 9.1 us per 4K frame
//       and 2.6 us vs 3.6 us per 1080p frame against
//   512 threads x 2 CTAs/SM with 256-byte table rows (one-PRMT addressing, 64 KB table).
//   The host launches the <kClaheCtas - 1> instantiation (more registers per thread) for tiles of 256 K pixels and more.
#ifndef NV12EQ_CLAHE_THREADS
#define NV12EQ_CLAHE_THREADS 256
#endif
#ifndef NV12EQ_CLAHE_ROWSHIFT
#define NV12EQ_CLAHE_ROWSHIFT 7
#endif
#ifndef NV12EQ_CLAHE_CTAS
#define NV12EQ_CLAHE_CTAS 4
#endif
constexpr int kCT = NV12EQ_CLAHE_THREADS;       // threads per CTA
constexpr int kCWarps = kCT / 32;
constexpr int kClaheCtas = NV12EQ_CLAHE_CTAS;   // CTAs per SM the kernel is built for
constexpr int kRowShift = NV12EQ_CLAHE_ROWSHIFT;
constexpr int kRowBytes = 1 << kRowShift;       // bytes per table row: hist[bin][32 lanes] u32 (128 B used) or table[v][reps] uint2
constexpr int kCellReps = kRowBytes / 8;        // 8-byte replicas of a cell table entry (32 or 16: conflict-free per half-warp)
static_assert(kRowShift == 7 || kRowShift == 8, "table rows are 128 or 256 bytes");
static_assert(kCT % 256 == 0 && kCT >= 256 && kCT <= 1024, "table build maps threads to the 256 values");
constexpr int kRowTableBytes = 256 * kRowBytes;
constexpr int kRingDepth = 8;                   // pixel rows in flight per thread in the cell loop (power of two)
constexpr int kTileDepth = 4;                   // 16-byte tile row pieces in flight per thread in the tile loop
constexpr int kRingBytes = kRingDepth * kCT * 8;
static_assert(kTileDepth * kCT * 16 <= kRingBytes, "tile ring must fit");
constexpr int kClaheSmemBytes = kRowTableBytes + kRingBytes;  // dynamic shared memory of clahe_kernel

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
    uint32_t slot_magic;        // floor(2^32 / items per slot) + 1 if items * items_per_slot < 2^32 (exact division by multiply), else 0
    int debug_skip;             // developer tool: bit0 skip tile histogram, bit1 skip cell blend, bit2 skip uv, bit3 skip LUT build, bit4 skip table build
};

// n / d for 0 <= n < 65536, 1 <= d <= 65536 in five instructions (an integer division costs ~20): (n + 0.5) / d is never
// closer than 0.5 / d to an integer, i.e. 2^-17 relative, and the approximate reciprocal is good to 2^-21.
__device__ __forceinline__ int small_div(int n, int d) {
    if (n >= 65536) return n / d;
    return __float2int_rz(__fmul_rn(__fadd_rn(__int2float_rn(n), 0.5f), __fdividef(1.0f, __int2float_rn(d))));
}
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ---- 256-byte-row shared tables -----------------------------------------------------------------------------
// Row v of the table starts at tbase + v*kRowBytes.  `lane_off` is this lane's byte offset inside a row.  With 256-byte
// rows it lives in byte 0 of the register and byte K of a packed pixel word goes to byte 1: one PRMT = the whole offset.
// With 128-byte rows it is PRMT (extract) + one multiply-add onto the loop-invariant tbase + lane_off.
template <int K>
__device__ __forceinline__ uint32_t row_addr(uint32_t w, uint32_t tbase, uint32_t lane_off) {
    if (kRowShift == 8) return tbase + __byte_perm(w, lane_off, 0x5504u | (K << 4));
    return (tbase + lane_off) + (byte_of<K>(w) << kRowShift);
}

// histogram rows: hist[bin][lane] u32 in the first 128 bytes of row `bin`
__device__ __forceinline__ void hist256_byte(uint32_t v, uint32_t tbase, uint32_t lane4) { red_shared_inc(tbase + (v << kRowShift) + lane4); }
__device__ __forceinline__ void hist256_word(uint32_t w, uint32_t tbase, uint32_t lane4) {
    red_shared_inc(row_addr<0>(w, tbase, lane4));
    red_shared_inc(row_addr<1>(w, tbase, lane4));
    red_shared_inc(row_addr<2>(w, tbase, lane4));
    red_shared_inc(row_addr<3>(w, tbase, lane4));
}
__device__ __forceinline__ void hist256_vec(int4 v, uint32_t tbase, uint32_t lane4) {
    hist256_word((uint32_t)v.x, tbase, lane4);
    hist256_word((uint32_t)v.y, tbase, lane4);
    hist256_word((uint32_t)v.z, tbase, lane4);
    hist256_word((uint32_t)v.w, tbase, lane4);
}
// n contiguous bytes by one warp (general tile path): 16-byte vectors where alignment allows, bytes elsewhere
__device__ __forceinline__ void hist256_span_warp(const uint8_t* __restrict__ p, int n, int lane, uint32_t tbase, uint32_t lane4,
                                                  uint64_t pol) {
    const int mis = (int)((16 - ((uintptr_t)p & 15)) & 15);
    const int head = min(mis, n);
    for (int i = lane; i < head; i += 32) hist256_byte(p[i], tbase, lane4);
    const int nvec = (n - head) >> 4;
    const uint8_t* v = p + head;
    for (int i = lane; i < nvec; i += 32) hist256_vec(ldg128_hint(v + (size_t)i * 16, pol), tbase, lane4);
    for (int i = head + (nvec << 4) + lane; i < n; i += 32) hist256_byte(p[i], tbase, lane4);
}
__device__ __forceinline__ void hist256_zero(uint32_t* tab) {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 2048 / kCT; ++k) {
        const int i = threadIdx.x + k * kCT;  // 16-byte slot i of the 128 counter bytes of every row: row i>>3, column i&7
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(tab) + (i >> 3) * kRowBytes + (i & 7) * 16) = z;
    }
}
__device__ __forceinline__ uint32_t hist256_row_sum(const uint32_t* tab, int bin) {
    const uint8_t* row = reinterpret_cast<const uint8_t*>(tab) + bin * kRowBytes;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 q = *reinterpret_cast<const uint4*>(row + ((j + bin) & 7) * 16);
        s += q.x + q.y + q.z + q.w;
    }
    return s;
}

// Tile histogram -> clip -> redistribute -> scan -> LUT bytes (global), by the whole CTA: thread t < 256 owns bin t
// (count = its histogram value).  The per-bin work (clip, residual test, conversion) runs 256 wide and the two reductions
// cost one barrier each; a single-warp version kept the other warps of the CTA waiting ~1 us per tile.
// s_scratch: 2 * kCWarps words.
__device__ __forceinline__ void clahe_tile_lut_block(uint32_t count, int clip_limit, float lut_scale, uint8_t* __restrict__ glut,
                                                     uint32_t* s_scratch, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    int h = tid < 256 ? (int)count : 0;
    if (clip_limit > 0) {   // uniform over the CTA
        int excess = max(h - clip_limit, 0);
        h = min(h, clip_limit);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, d);
        if (lane == 0) s_scratch[warp] = (uint32_t)excess;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int w = 0; w < kCWarps; ++w) clipped += (int)s_scratch[w];
        const int batch = clipped >> 8, residual = clipped & 255;   // clipped >= 0
        h += batch;
        if (tid < 256 && residual != 0) {
            // for (i = 0; i < 256 && residual > 0; i += step, residual--) h[i]++
            const int step = max(256 / residual, 1);
            const int q = tid / step;
            if (q * step == tid && q < residual) h += 1;
        }
        if (tid >= 256) h = 0;
    }
    const uint32_t incl = warp_incl_scan((uint32_t)h, lane);
    if (lane == 31) s_scratch[kCWarps + warp] = incl;
    __syncthreads();
    uint32_t run = incl;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w)
        if (w < warp) run += s_scratch[kCWarps + w];
    if (tid < 256) glut[tid] = (uint8_t)round_sat_u8(__fmul_rn(__int2float_rn((int)run), lut_scale));
}

// ---- the blend ------------------------------------------------------------------------------------------------
// Table entry e = {bf16 L11 | bf16 L21 << 16, bf16 L12 | bf16 L22 << 16}; yw = (ya1, ya) packed.
// The sum of the two packed products is an add.rn.FTZ.f32x2: ptxas contracts a plain add.rn.f32x2 of mul.rn.f32x2
// results into FFMA2 (even with -fmad=false), which rounds once instead of twice and breaks bit-exactness with
// OpenCV's unfused arithmetic; it does not contract across an .ftz mismatch.  FTZ itself is value-neutral here: the
// operands are products of integers 0..255 and weights that are 0 or >= 2^-25, never subnormal.
// (tests/test_abi.py checks that the built library contains no FFMA2 at all.)
__device__ __forceinline__ uint64_t add_f2_nofuse(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// res of one pixel: 0 <= res < 255.5 (convex-ish combination of values in [0,255]), so adding 1.5*2^23 afterwards
// performs cvRound's round-half-to-even and leaves the integer in the low mantissa byte; saturate_cast is the identity.
__device__ __forceinline__ float clahe_blend_res(uint2 e, float xa, float xa1, uint64_t yw) {
    const uint64_t A = pack_f2(__uint_as_float(e.x << 16), __uint_as_float(e.x & 0xffff0000u));  // (L11, L21)
    const uint64_t B = pack_f2(__uint_as_float(e.y << 16), __uint_as_float(e.y & 0xffff0000u));  // (L12, L22)
    const uint64_t S = add_f2_nofuse(mul_f2(A, pack_f2(xa1, xa1)), mul_f2(B, pack_f2(xa, xa)));   // (top, bot)
    float r0, r1;
    unpack_f2(mul_f2(S, yw), r0, r1);  // (top*ya1, bot*ya)
    return __fadd_rn(r0, r1);
}
template <int K>
__device__ __forceinline__ float clahe_blend_px(uint32_t w, uint32_t tbase, uint32_t lane8, float xa, float xa1, uint64_t yw) {
    return clahe_blend_res(lds_u64(row_addr<K>(w, tbase, lane8)), xa, xa1, yw);
}
__device__ __forceinline__ void round_pair(float a, float b, uint32_t& oa, uint32_t& ob) {
    unpack_u2(add_f2(pack_f2(a, b), pack_f2(12582912.0f, 12582912.0f)), oa, ob);
}
// low bytes of four words -> one packed word (3 PRMT)
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// eight horizontally adjacent pixels (two packed words) -> two packed output words
__device__ __forceinline__ uint2 clahe_blend_8(uint2 px, uint32_t lane8, const float* xa, const float* xa1, uint64_t yw) {
    // re-read the table base here: a fresh uniform value lets ptxas address the gathers as [R + UR] instead of adding
    // a base held in a vector register to every offset
    const uint32_t tbase = smem_u32(nv12eq_smem_rows);
    float f[8];
    f[0] = clahe_blend_px<0>(px.x, tbase, lane8, xa[0], xa1[0], yw);
    f[1] = clahe_blend_px<1>(px.x, tbase, lane8, xa[1], xa1[1], yw);
    f[2] = clahe_blend_px<2>(px.x, tbase, lane8, xa[2], xa1[2], yw);
    f[3] = clahe_blend_px<3>(px.x, tbase, lane8, xa[3], xa1[3], yw);
    f[4] = clahe_blend_px<0>(px.y, tbase, lane8, xa[4], xa1[4], yw);
    f[5] = clahe_blend_px<1>(px.y, tbase, lane8, xa[5], xa1[5], yw);
    f[6] = clahe_blend_px<2>(px.y, tbase, lane8, xa[6], xa1[6], yw);
    f[7] = clahe_blend_px<3>(px.y, tbase, lane8, xa[7], xa1[7], yw);
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; k += 2) round_pair(f[k], f[k + 1], o[k], o[k + 1]);
    return make_uint2(pack_low_bytes(o[0], o[1], o[2], o[3]), pack_low_bytes(o[4], o[5], o[6], o[7]));
}

__device__ __forceinline__ void axis_weight(int pos, float inv, float& a, float& a1) {
    const float f = __fsub_rn(__fmul_rn((float)pos, inv), 0.5f);
    const float t1 = floorf(f);
    a = __fsub_rn(f, t1);
    a1 = __fsub_rn(1.0f, a);
}
// keeps a value in its register: stops the compiler from re-deriving xa1 = 1 - xa inside the pixel loop
__device__ __forceinline__ void pin_register(float& v) { asm volatile("" : "+f"(v)); }

// The rows of one thread inside a cell: G (8 or 16) horizontally adjacent pixels per row, rows tr, tr + rpp, ...
// Thread-private ring of D G-byte slots in shared memory, filled with cp.async: the rows D-1 ahead are in flight
// (about 50 KB per SM) without holding registers, and since a thread only ever reads its own slots no barrier is
// involved.  The row loop is unrolled by the ring depth, so every ring slot is a compile-time offset.  Thread 0
// (tr == 0, the most rows) draws the next ticket at the start of the last round, which hides the atomic's round trip.
// G = 16 halves the per-row overhead (ring, pointers, y weights, loop) per pixel; it needs 32 registers of x weights.
template <int G>
struct CellRows {
    static constexpr int D = (G == 16) ? kRingDepth / 2 : kRingDepth;   // same ring bytes either way
    static constexpr uint32_t kSlot = kCT * G;
    float xa[G], xa1[G];
    const uint8_t* spn;
    uint8_t* dp;
    size_t rstep;
    uint32_t ring0, yw_addr, yw_step;
    int nrows;

    __device__ __forceinline__ void issue(uint32_t slot, const uint8_t* g) const {
        if (G == 16) cp_async16(ring0 + slot * kSlot, g); else cp_async8(ring0 + slot * kSlot, g);
    }
    // Everything that does not depend on the tile LUTs: x weights and the first D-1 rows of the ring.  Called BEFORE the
    // dependency wait and the table build, so the pixel loads are in flight while the CTA waits for / packs the LUTs.
    __device__ __forceinline__ void start(const uint8_t* sp, uint8_t* dp_, size_t rstep_, int nrows_, int xg, float inv_tw, uint32_t ring0_,
                                          uint32_t yw_addr_, uint32_t yw_step_) {
        dp = dp_; rstep = rstep_; nrows = nrows_; ring0 = ring0_; yw_addr = yw_addr_; yw_step = yw_step_;
#pragma unroll
        for (int k = 0; k < G; ++k) axis_weight(xg + k, inv_tw, xa[k], xa1[k]);
#pragma unroll
        for (int j = 0; j < D - 1; ++j) {
            if (j < nrows) issue((uint32_t)j, sp + (size_t)j * rstep);
            cp_async_commit();
        }
        spn = sp + (size_t)(D - 1) * rstep;
    }
    __device__ __forceinline__ void run(TicketQueue& q, uint32_t lane8) {
#pragma unroll 1
        for (int i0 = 0; i0 < nrows; i0 += D) {
            if (i0 + D >= nrows) q.prefetch();
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const int i = i0 + j;
                if (i < nrows) {
                    if (i + D - 1 < nrows) issue((uint32_t)((j + D - 1) % D), spn);
                    cp_async_commit();
                    cp_async_wait<D - 1>();
                    const uint64_t yw = lds_b64(yw_addr);
                    if (G == 16) {
                        const int4 px = lds_s4(ring0 + (uint32_t)j * kSlot);
                        const uint2 o0 = clahe_blend_8(make_uint2((uint32_t)px.x, (uint32_t)px.y), lane8, xa, xa1, yw);
                        const uint2 o1 = clahe_blend_8(make_uint2((uint32_t)px.z, (uint32_t)px.w), lane8, xa + 8 * (G / 16), xa1 + 8 * (G / 16), yw);
                        __stcs(reinterpret_cast<uint4*>(dp), make_uint4(o0.x, o0.y, o1.x, o1.y));
                    } else {
                        const uint2 px = lds_u64(ring0 + (uint32_t)j * kSlot);
                        __stcs(reinterpret_cast<uint2*>(dp), clahe_blend_8(px, lane8, xa, xa1, yw));
                    }
                    spn += rstep; dp += rstep; yw_addr += yw_step;
                }
            }
        }
    }
};

template <int MIN_CTAS>
__global__ void __launch_bounds__(kCT, MIN_CTAS) clahe_kernel(const ClaheParams p) {
    uint32_t* const smem_rows = nv12eq_smem_rows;  // 64 KB of 256-byte rows: hist[bin][32] (tile items) / table[v][32] (cells); then the ring
    __shared__ uint32_t s_bins[256];
    __shared__ __align__(8) float2 s_yw[kMaxCellRows];  // (ya1, ya) of the rows of the current cell
    __shared__ uint32_t s_ticket[2];
    __shared__ int s_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.tx * p.ty;
    const int I = p.nxc * p.nyc;
    const int U = p.uv_chunks;
    const int per_slot = T + I + U;
    const uint32_t total_items = (uint32_t)(p.n_frames + p.lag) * (uint32_t)per_slot;
    uint32_t tbase;
    asm("mov.u32 %0, nv12eq_smem_rows;" : "=r"(tbase));
    const uint32_t rbase = tbase + kRowTableBytes;  // cp.async ring of the cell loop
    const uint32_t lane4 = (uint32_t)lane * 4u, lane8 = (uint32_t)(lane & (kCellReps - 1)) * 8u;

    TicketQueue q{p.ticket, s_ticket, 0u, 0u, false};
    q.start();
    for (;;) {
        const uint32_t item = q.current();
        if (item >= total_items) break;
        // item = g * per_slot + r; the host supplies floor(2^32 / per_slot) + 1 when that multiplier divides exactly
        const int g = p.slot_magic ? (int)__umulhi(item, p.slot_magic) : (int)(item / (uint32_t)per_slot);
        const int r = (int)(item - (uint32_t)g * (uint32_t)per_slot);
        const int f = g - p.lag;
        const ItemTrace tr_{p.trace};
        tr_.mark(item, 0);
        tr_.mark(item, 1);
        tr_.kind(item, r < T ? 1u : (r < T + I ? 2u : 3u));

        if (r < T) {
            if (g < p.n_frames) {
                // ------------------------- tile item: histogram -> clip -> LUT -------------------------
                const uint8_t* y = p.in + (unsigned long long)g * p.pitch;
                const int tyi = small_div(r, p.tx), txi = r - tyi * p.tx;
                const int x0 = txi * p.tw, y0 = tyi * p.th;
                const uint64_t keep = l2_policy_evict_last();
                const bool vec_ok = !(p.debug_skip & 1) && !p.padded && (p.tw & 15) == 0 && (p.stride & 15) == 0 &&
                                    (((uintptr_t)y + (uintptr_t)x0) & 15) == 0 && (p.tw >> 4) <= kCT;
                // threads form a (rows_per_pass x vectors_per_row) grid over the tile.  Every thread keeps
                // kTileDepth-1 of its 16-byte row pieces in flight through a private cp.async ring; the first ones are issued
                // before the table is zeroed so that their latency overlaps the set-up of the item.
                const int vpr = vec_ok ? (p.tw >> 4) : 1;
                const int rpp = small_div(kCT, vpr);
                const int tr = small_div(tid, vpr), tc = tid - tr * vpr;
                const int nrows = (vec_ok && tr < rpp && tr < p.th) ? small_div(p.th - tr + rpp - 1, rpp) : 0;
                const uint8_t* ptr = y + (size_t)(y0 + tr) * p.stride + x0 + tc * 16;
                const size_t rstep = (size_t)rpp * p.stride;
                const uint32_t ring0 = rbase + (uint32_t)tid * 16u;
                constexpr uint32_t kSlot = kCT * 16u;
                if (vec_ok) {
#pragma unroll
                    for (int j = 0; j < kTileDepth - 1; ++j) {
                        if (j < nrows) cp_async16(ring0 + (uint32_t)j * kSlot, ptr + (size_t)j * rstep);
                        cp_async_commit();
                    }
                }
                hist256_zero(smem_rows);
                __syncthreads();
                if (p.debug_skip & 1) {
                } else if (vec_ok) {
                    const uint8_t* pn = ptr + (size_t)(kTileDepth - 1) * rstep;
                    // unrolled by the ring depth: every slot is a compile-time offset
#pragma unroll 1
                    for (int i0 = 0; i0 < nrows; i0 += kTileDepth) {
#pragma unroll
                        for (int j = 0; j < kTileDepth; ++j) {
                            const int i = i0 + j;
                            if (i < nrows) {
                                if (i + kTileDepth - 1 < nrows) cp_async16(ring0 + (uint32_t)((j + kTileDepth - 1) % kTileDepth) * kSlot, pn);
                                cp_async_commit();
                                cp_async_wait<kTileDepth - 1>();
                                hist256_vec(lds_s4(ring0 + (uint32_t)j * kSlot), tbase, lane4);
                                pn += rstep;
                            }
                        }
                    }
                } else {
                    // general path: one warp per tile row, byte spans inside the image, reflected reads outside
                    for (int row = warp; row < p.th; row += kCWarps) {
                        const uint8_t* src_row = y + (size_t)reflect101(y0 + row, p.h) * p.stride;
                        const int xin = min(x0 + p.tw, p.w);  // end of the in-image part
                        if (x0 < xin) hist256_span_warp(src_row + x0, xin - x0, lane, tbase, lane4, keep);
                        for (int x = max(x0, p.w) + lane; x < x0 + p.tw; x += 32)
                            hist256_byte(src_row[reflect101(x, p.w)], tbase, lane4);
                    }
                }
                q.prefetch();  // the row sums and the LUT build hide the ticket round trip
                __syncthreads();
                if (!(p.debug_skip & 8))
                    clahe_tile_lut_block(tid < 256 ? hist256_row_sum(smem_rows, tid) : 0u, p.clip_limit, p.lut_scale,
                                         p.luts + ((size_t)g * T + r) * 256, s_bins, tid);
                __threadfence();   // every thread publishes its LUT byte before the tile is counted
                __syncthreads();
                if (tid == 0) atomicAdd(p.tiles_done + g, 1u);
            }
        } else if (f >= 0) {
            const uint8_t* src = p.in + (unsigned long long)f * p.pitch;
            uint8_t* dst = p.out + (unsigned long long)f * p.pitch;
            if (r < T + I) {
                // ------------------------- cell item: blend four tile LUTs -------------------------
                const int ci = r - T;
                const int cy = small_div(ci, p.nxc), cx = ci - cy * p.nxc;
                const int4 xc = p.xcells[cx], yc = p.ycells[cy];
                const int cw = xc.y - xc.x, ch = yc.y - yc.x;  // cell size in pixels (ch <= kMaxCellRows)
                // Geometry of the fast path: full groups of G = 16 (or 8) pixels, one 16- (8-) byte load / store per thread and
                // row.  Columns left over when the cell width is not a multiple of G (and everything when alignment does not
                // allow vector accesses) take the pixel-at-a-time path below.
                const uint32_t ywbase = smem_u32(s_yw);
                const uintptr_t align_or = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)p.stride | (uintptr_t)xc.x;
                const bool fast16 = ((align_or & 15) == 0) && (cw & 15) == 0 && (cw >> 4) >= 1 && (cw >> 4) <= kCT;
                const bool fast8 = !fast16 && ((align_or & 7) == 0) && (cw >> 3) >= 1 && (cw >> 3) <= kCT;
                const int G = fast16 ? 16 : 8;
                const int gpr = (p.debug_skip & 2) ? 0 : (fast16 ? (cw >> 4) : (fast8 ? (cw >> 3) : 0));   // G-pixel groups per row
                const int xslow = xc.x + gpr * G;                               // first column of the pixel-at-a-time path
                const int rpp = gpr ? small_div(kCT, gpr) : 1;
                const int tr = gpr ? small_div(tid, gpr) : 0, tc = tid - tr * gpr;
                const bool active = gpr > 0 && tr < rpp && tr < ch;
                const int xg = xc.x + tc * G;
                const size_t rstep = (size_t)rpp * p.stride;
                const int nrows = active ? small_div(ch - tr + rpp - 1, rpp) : 0;   // rows of this thread: tr, tr + rpp, ...
                const uint8_t* sp = src + (size_t)(yc.x + tr) * p.stride + xg;
                uint8_t* dp = dst + (size_t)(yc.x + tr) * p.stride + xg;
                const uint32_t yw_addr = ywbase + (uint32_t)tr * 8u, yw_step = (uint32_t)rpp * 8u;

                // dependency wait + table build; everything the row loop needs that does not depend on the LUTs has been
                // started by then (CellRows::start), so pixel loads overlap the wait and the LUT loads
                auto wait_and_build_table = [&]() -> bool {
                    // y weights of the cell's rows
                    for (int i = tid; i < ch; i += kCT) {
                        float ya, ya1;
                        axis_weight(yc.x + i, p.inv_th, ya, ya1);
                        s_yw[i] = make_float2(ya1, ya);
                    }
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
                    if (!s_flag) return false;
                    tr_.mark(item, 1);
                    if (!(p.debug_skip & 16))
                    // pack the four LUTs: row v = replicas of {bf16 L11 | L21 << 16, bf16 L12 | L22 << 16}
                    {
                        constexpr int kShare = kCT / 256, kPer = kCellReps / kShare;
                        const uint8_t* L = p.luts + (size_t)f * T * 256;
                        const int v = tid & 255, part = tid >> 8;
                        const uint32_t l11 = __ldcg(L + (size_t)(yc.z * p.tx + xc.z) * 256 + v);
                        const uint32_t l12 = __ldcg(L + (size_t)(yc.z * p.tx + xc.w) * 256 + v);
                        const uint32_t l21 = __ldcg(L + (size_t)(yc.w * p.tx + xc.z) * 256 + v);
                        const uint32_t l22 = __ldcg(L + (size_t)(yc.w * p.tx + xc.w) * 256 + v);
                        uint2 e;
                        e.x = (__float_as_uint((float)l11) >> 16) | (__float_as_uint((float)l21) & 0xffff0000u);
                        e.y = (__float_as_uint((float)l12) >> 16) | (__float_as_uint((float)l22) & 0xffff0000u);
                        uint2* row = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(smem_rows) + v * kRowBytes);
#pragma unroll
                        for (int j = 0; j < kPer; ++j) row[(part * kPer + j + v) & (kCellReps - 1)] = e;
                    }
                    __syncthreads();
                    return true;
                };
                bool ok;
                if (fast16) {
                    CellRows<16> rows;
                    rows.start(sp, dp, rstep, nrows, xg, p.inv_tw, rbase + (uint32_t)tid * 16u, yw_addr, yw_step);
                    ok = wait_and_build_table();
                    if (ok) rows.run(q, lane8);
                } else {
                    CellRows<8> rows;
                    rows.start(sp, dp, rstep, nrows, xg, p.inv_tw, rbase + (uint32_t)tid * 8u, yw_addr, yw_step);
                    ok = wait_and_build_table();
                    if (ok) rows.run(q, lane8);
                }
                if (!ok) break;
                if (xslow < xc.y && !(p.debug_skip & 2)) {
                    // pixel-at-a-time path
                    const int sw = xc.y - xslow;
                    const long long npix = (long long)sw * ch;
                    for (long long i = tid; i < npix; i += kCT) {
                        const int ry = (int)(i / sw), rx = (int)(i - (long long)ry * sw);
                        const int x = xslow + rx, yy = yc.x + ry;
                        float xa, xa1;
                        axis_weight(x, p.inv_tw, xa, xa1);
                        const uint32_t v = src[(size_t)yy * p.stride + x];
                        const float res = clahe_blend_res(lds_u64(tbase + (v << kRowShift) + lane8), xa, xa1, lds_b64(ywbase + (uint32_t)ry * 8u));
                        dst[(size_t)yy * p.stride + x] = (uint8_t)__float_as_uint(__fadd_rn(res, 12582912.0f));
                    }
                }
            } else {
                // ------------------------- uv item -------------------------
                q.prefetch();  // short item: draw the next ticket right away
                if (!(p.debug_skip & 4)) {
                const int c = r - T - I;
                const bool copy_uv = p.uv_mode == UV_COPY && src != dst;
                const size_t uv_off = (size_t)p.stride * p.h;
                if (p.flat) {
                    const unsigned long long b0 = min((unsigned long long)c * p.uv_chunk, p.uv_bytes);
                    const unsigned long long b1 = min(b0 + p.uv_chunk, p.uv_bytes);
                    if (copy_uv) copy_span(src + uv_off + b0, dst + uv_off + b0, (size_t)(b1 - b0), tid, kCT);
                    else if (p.uv_mode == UV_GRAY128) fill_span(dst + uv_off + b0, (size_t)(b1 - b0), tid, kCT, 128);
                } else {
                    const int rows = p.h / 2;
                    const int r0 = min(c * p.uv_rows_chunk, rows), r1 = min(r0 + p.uv_rows_chunk, rows);
                    for (int row = r0 + warp; row < r1; row += kCWarps) {
                        const size_t off = uv_off + (size_t)row * p.stride;
                        if (copy_uv) copy_span(src + off, dst + off, (size_t)p.w, lane, 32);
                        else if (p.uv_mode == UV_GRAY128) fill_span(dst + off, (size_t)p.w, lane, 32, 128);
                    }
                }
                }
            }
        }
        tr_.mark(item, 2);
        q.advance();
    }
    // the last CTA out returns the per-frame counters to zero for the next launch
    if (q.finish())
        for (int i = tid; i < p.n_frames; i += kCT) p.tiles_done[i] = 0;
}

}  // namespace nv12eq
