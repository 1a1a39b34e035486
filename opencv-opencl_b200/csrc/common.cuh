// common.cuh -- shared device helpers for libnv12eq (sm_100a).
//
// Data layout conventions used by every kernel in this directory:
//   * A frame batch lives in HBM as n_frames NV12 frames, frame k at base + k*pitch; inside a frame H rows of Y
//     then H/2 rows of interleaved UV, rows `stride` bytes apart (stride == width is the "flat" fast case: the
//     whole Y plane is one contiguous byte span, cut into 16-byte vectors).
//   * 256-bin histograms are accumulated in shared memory as hist[bin][lane] (uint32, 32 KB): lane L of every
//     warp only ever touches column L, so a warp-wide shared atomic hits 32 distinct banks regardless of the
//     pixel values -- no bank conflicts, no same-address serialisation, for flat patches and noise alike.
//   * Look-up tables are expanded in shared memory as table[value][lane] (uint32, byte-replicated, 32 KB) for the
//     same reason: the gather `table[pixel][lane]` is conflict-free for arbitrary pixel values.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nv12eq {

constexpr int kThreads = 256;            // threads per CTA for every kernel here
constexpr int kWarps = kThreads / 32;
constexpr int kLaneTableWords = 256 * 32; // one [256][32] uint32 table
constexpr int kLaneTableBytes = kLaneTableWords * 4;

enum : int { UV_COPY = 0, UV_GRAY128 = 1, UV_SKIP = 2 };

// ---- memory access helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// no-return shared atomic increment on a 32-bit shared address
__device__ __forceinline__ void red_shared_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Y plane first read: keep it in L2 (it is read again by the apply / interpolation pass).
__device__ __forceinline__ int4 ld_keep(const int4* p) { return __ldg(p); }
// Last read of a line / write-once output: streaming (evict-first) so it does not push the Y plane out of L2.
__device__ __forceinline__ int4 ld_stream(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(int4* p, int4 v) { __stcs(p, v); }

// ---- byte-span walkers ----------------------------------------------------------------------------------
// A span is n contiguous bytes.  `tid`/`nthr` select the participating threads (a whole CTA for flat planes, one
// warp for one row of a strided plane).  Unaligned heads/tails are handled with byte accesses.
struct SpanSplit {
    size_t head;  // bytes before the first 16-byte aligned address
    size_t nvec;  // number of 16-byte vectors
    size_t tail0; // offset of the first tail byte
};
__device__ __forceinline__ SpanSplit split_span(const void* p, size_t n) {
    SpanSplit s;
    size_t mis = (size_t)((16 - ((uintptr_t)p & 15)) & 15);
    s.head = mis < n ? mis : n;
    s.nvec = (n - s.head) >> 4;
    s.tail0 = s.head + (s.nvec << 4);
    return s;
}

// A lane-private table is addressed as base + ((value << 7) | lane4): `base` is a CTA-uniform pointer into shared
// memory (ends up in a uniform register: ATOMS/LDS [R + UR]), lane4 = lane * 4.  Extracting byte k of a packed
// word, scaling it by 128 and OR-ing in the lane offset is one shift plus one LOP3 per pixel.
struct LaneTable {
    char* base;       // CTA-uniform, points into shared memory
    uint32_t lane4;   // (threadIdx.x & 31) * 4
};
__device__ __forceinline__ uint32_t* lt_ptr(const LaneTable& t, uint32_t scaled_value) {
    return reinterpret_cast<uint32_t*>(t.base + (scaled_value | t.lane4));
}
__device__ __forceinline__ void hist_byte(uint32_t v, const LaneTable& t) { atomicAdd(lt_ptr(t, v << 7), 1u); }
__device__ __forceinline__ void hist_word(uint32_t w, const LaneTable& t) {
    atomicAdd(lt_ptr(t, (w << 7) & 0x7f80u), 1u);
    atomicAdd(lt_ptr(t, (w >> 1) & 0x7f80u), 1u);
    atomicAdd(lt_ptr(t, (w >> 9) & 0x7f80u), 1u);
    atomicAdd(lt_ptr(t, (w >> 17) & 0x7f80u), 1u);
}
__device__ __forceinline__ void hist_vec(int4 v, const LaneTable& t) {
    hist_word((uint32_t)v.x, t);
    hist_word((uint32_t)v.y, t);
    hist_word((uint32_t)v.z, t);
    hist_word((uint32_t)v.w, t);
}

template <int UNROLL>
__device__ __forceinline__ void hist_span(const uint8_t* __restrict__ p, size_t n, int tid, int nthr,
                                          const LaneTable& hist) {
    SpanSplit s = split_span(p, n);
    for (size_t i = tid; i < s.head; i += nthr) hist_byte(p[i], hist);
    const int4* v = reinterpret_cast<const int4*>(p + s.head);
    size_t i = tid;
    const size_t step = (size_t)nthr * UNROLL;
    for (; i + step - nthr < s.nvec; i += step) {
        int4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) r[u] = ld_keep(v + i + (size_t)u * nthr);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) hist_vec(r[u], hist);
    }
    for (; i < s.nvec; i += nthr) hist_vec(ld_keep(v + i), hist);
    for (size_t j = s.tail0 + tid; j < n; j += nthr) hist_byte(p[j], hist);
}

// Gather through a byte-replicated lane table (entries are lut * 0x01010101).
__device__ __forceinline__ uint32_t lut_word(uint32_t w, const LaneTable& t) {
    const uint32_t a = *lt_ptr(t, (w << 7) & 0x7f80u);
    const uint32_t b = *lt_ptr(t, (w >> 1) & 0x7f80u);
    const uint32_t c = *lt_ptr(t, (w >> 9) & 0x7f80u);
    const uint32_t d = *lt_ptr(t, (w >> 17) & 0x7f80u);
    const uint32_t lo = __byte_perm(a, b, 0x5140);  // a.b0, b.b0 in the low half (all bytes of a, b are equal)
    const uint32_t hi = __byte_perm(c, d, 0x5140);
    return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ int4 lut_vec(int4 v, const LaneTable& t) {
    int4 o;
    o.x = (int)lut_word((uint32_t)v.x, t);
    o.y = (int)lut_word((uint32_t)v.y, t);
    o.z = (int)lut_word((uint32_t)v.z, t);
    o.w = (int)lut_word((uint32_t)v.w, t);
    return o;
}
__device__ __forceinline__ uint8_t lut_byte(uint8_t v, const LaneTable& t) { return (uint8_t)*lt_ptr(t, (uint32_t)v << 7); }

// dst[i] = table[src[i]] over a span.  src and dst must be congruent mod 16 for the vector path; otherwise the
// whole span goes through the byte path (correct, slow, only hit by exotic pointer/pitch combinations).
template <int UNROLL>
__device__ __forceinline__ void lut_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                         int nthr, const LaneTable& table) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = lut_byte(src[i], table);
        return;
    }
    SpanSplit s = split_span(src, n);
    for (size_t i = tid; i < s.head; i += nthr) dst[i] = lut_byte(src[i], table);
    const int4* v = reinterpret_cast<const int4*>(src + s.head);
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    size_t i = tid;
    const size_t step = (size_t)nthr * UNROLL;
    for (; i + step - nthr < s.nvec; i += step) {
        int4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) r[u] = ld_stream(v + i + (size_t)u * nthr);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(o + i + (size_t)u * nthr, lut_vec(r[u], table));
    }
    for (; i < s.nvec; i += nthr) st_stream(o + i, lut_vec(ld_stream(v + i), table));
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = lut_byte(src[j], table);
}

template <int UNROLL>
__device__ __forceinline__ void copy_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                          int nthr) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = src[i];
        return;
    }
    SpanSplit s = split_span(src, n);
    for (size_t i = tid; i < s.head; i += nthr) dst[i] = src[i];
    const int4* v = reinterpret_cast<const int4*>(src + s.head);
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    size_t i = tid;
    const size_t step = (size_t)nthr * UNROLL;
    for (; i + step - nthr < s.nvec; i += step) {
        int4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) r[u] = ld_stream(v + i + (size_t)u * nthr);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(o + i + (size_t)u * nthr, r[u]);
    }
    for (; i < s.nvec; i += nthr) st_stream(o + i, ld_stream(v + i));
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = src[j];
}

__device__ __forceinline__ void fill_span(uint8_t* __restrict__ dst, size_t n, int tid, int nthr, uint8_t value) {
    SpanSplit s = split_span(dst, n);
    for (size_t i = tid; i < s.head; i += nthr) dst[i] = value;
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    const int w = (int)(value * 0x01010101u);
    const int4 vv = make_int4(w, w, w, w);
    for (size_t i = tid; i < s.nvec; i += nthr) st_stream(o + i, vv);
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = value;
}

// ---- lane-private shared-memory tables ------------------------------------------------------------------
// Zero a [256][32] uint32 table with 16-byte stores.
__device__ __forceinline__ void lane_table_zero(uint32_t* tab) {
    uint4* t4 = reinterpret_cast<uint4*>(tab);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < kLaneTableWords / 4 / kThreads; ++i) t4[threadIdx.x + i * kThreads] = z;
}
// Sum row `bin` (32 lane columns) of a [256][32] table.  Rotated 16-byte reads keep the 8 threads of a
// quarter-warp phase on 8 different bank groups.
__device__ __forceinline__ uint32_t lane_table_row_sum(const uint32_t* tab, int bin) {
    const uint4* row = reinterpret_cast<const uint4*>(tab + bin * 32);
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint4 q = row[(j + bin) & 7];
        s += q.x + q.y + q.z + q.w;
    }
    return s;
}
// Expand a 256-byte LUT (shared memory) into table[v][lane] = lut[v] * 0x01010101.
__device__ __forceinline__ void lane_table_fill_from_lut(uint32_t* tab, const uint8_t* lut256) {
    const int v = threadIdx.x;  // kThreads == 256
    const uint32_t w = (uint32_t)lut256[v] * 0x01010101u;
    uint4* row = reinterpret_cast<uint4*>(tab + v * 32);
    const uint4 q = make_uint4(w, w, w, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) row[(j + v) & 7] = q;
}

// ---- warp scan ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// cv::saturate_cast<uchar>(cvRound(x)): round-half-even, clamp.
__device__ __forceinline__ uint32_t round_sat_u8(float x) {
    int r = __float2int_rn(x);
    return (uint32_t)min(max(r, 0), 255);
}

// Appendix B hash
__device__ __forceinline__ uint32_t fmix32(uint32_t k) {
    k ^= k >> 16; k *= 0x85EBCA6Bu; k ^= k >> 13; k *= 0xC2B2AE35u; k ^= k >> 16;
    return k;
}

// Bounded spin on a global counter (thread 0 only); returns false on timeout (~2 s) so a logic error can never
// hang the GPU.  Progress is guaranteed by construction: the counter is only waited on by CTAs whose work ticket
// is larger than the tickets of all CTAs that feed it, and those are resident or finished.
__device__ __forceinline__ bool spin_until_ge(const uint32_t* ctr, uint32_t target) {
    if (ld_acquire_u32(ctr) >= target) return true;
    const long long t0 = clock64();
    unsigned ns = 32;
    while (ld_acquire_u32(ctr) < target) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (clock64() - t0 > (4ll << 30)) return false;
    }
    return true;
}

}  // namespace nv12eq
