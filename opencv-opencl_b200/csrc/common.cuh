// common.cuh -- shared device helpers for libnv12eq (sm_100a).
//
// Data layout conventions used by every kernel in this directory:
//   * A frame batch lives in HBM as n_frames NV12 frames, frame k at base + k*pitch; inside a frame H rows of Y
//     then H/2 rows of interleaved UV, rows `stride` bytes apart (stride == width is the "flat" fast case: the
//     whole Y plane is one contiguous byte span, cut into 16-byte vectors).
//   * 256-bin histograms are accumulated in shared memory as hist[bin][lane] (uint32, 32 KB): lane L of every
//     warp only ever touches column L, so a warp-wide shared atomic hits 32 distinct banks regardless of the
//     pixel values -- no bank conflicts, no same-address serialisation, for flat patches and noise alike
//     (ncu: 1.1 shared wavefronts per ATOMS, the same cost as a conflict-free load).
//   * Look-up tables are expanded in shared memory as table[value][lane] (uint32, byte-replicated, 32 KB) for the
//     same reason: the gather `table[pixel][lane]` is conflict-free for arbitrary pixel values.
//   * One CTA works on one item at a time and an item needs only one of the two tables, so they share the same
//     32 KB of dynamic shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nv12eq {

constexpr int kThreads = 512;            // threads per CTA of the hot kernels (2 CTAs/SM: 32 warps, 64 registers/thread)
constexpr int kWarps = kThreads / 32;
constexpr int kLaneTableWords = 256 * 32; // one [256][32] uint32 table
constexpr int kLaneTableBytes = kLaneTableWords * 4;

enum : int { UV_COPY = 0, UV_GRAY128 = 1, UV_SKIP = 2 };

// ---- memory access helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// no-return shared atomic increment on a 32-bit shared address (SASS: ATOMS.POPC.INC)
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

// Byte k of a packed word, zero extended (one PRMT).
template <int K>
__device__ __forceinline__ uint32_t byte_of(uint32_t w) { return __byte_perm(w, 0u, 0x4440u + K); }

// ---- byte-span walkers ----------------------------------------------------------------------------------
// A span is n contiguous bytes.  `tid`/`nthr` select the participating threads (a whole CTA for flat planes, one
// warp for one row of a strided plane).  Unaligned heads/tails are handled with byte accesses.
struct SpanSplit {
    uint32_t head;  // bytes before the first 16-byte aligned address
    uint32_t nvec;  // number of 16-byte vectors (a span is one chunk or one row: far below 2^32 vectors)
    size_t tail0;   // offset of the first tail byte
};
__device__ __forceinline__ SpanSplit split_span(const void* p, size_t n) {
    SpanSplit s;
    const size_t mis = (size_t)((16 - ((uintptr_t)p & 15)) & 15);
    s.head = (uint32_t)(mis < n ? mis : n);
    s.nvec = (uint32_t)((n - s.head) >> 4);
    s.tail0 = (size_t)s.head + ((size_t)s.nvec << 4);
    return s;
}

// Software-pipelined walk over nvec 16-byte vectors: the loads of round r+1 are issued before round r is consumed,
// so every thread always has U vector loads in flight while it works (ncu showed the unpipelined loop stalled on
// long_scoreboard for a third of its samples).
template <int U, class Load, class Use>
__device__ __forceinline__ void pipelined_vectors(uint32_t nvec, int tid, int nthr, Load load, Use use) {
    const uint32_t step = (uint32_t)nthr * U;
    uint32_t i = tid;
    // rounds in which all U vectors of this thread are in range
    uint32_t rounds = (i + step - nthr < nvec) ? (nvec - (i + step - nthr) + step - 1) / step : 0u;
    int4 cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = make_int4(0, 0, 0, 0);
    if (rounds) {
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = load(i + (uint32_t)u * nthr);
    }
    for (; rounds; --rounds) {
        int4 nxt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = make_int4(0, 0, 0, 0);
        if (rounds > 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = load(i + step + (uint32_t)u * nthr);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) use(i + (uint32_t)u * nthr, cur[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        i += step;
    }
    for (; i < nvec; i += nthr) use(i, load(i));
}

// A lane-private table lives at 32-bit shared address `lane_base` (= table base + lane*4); the entry of value v
// is at lane_base + v*128.  Per pixel: one PRMT (byte extract), one shift-add, one shared access.
__device__ __forceinline__ void hist_byte(uint32_t v, uint32_t lane_base) { red_shared_inc(lane_base + (v << 7)); }
__device__ __forceinline__ void hist_word(uint32_t w, uint32_t lane_base) {
    red_shared_inc(lane_base + (byte_of<0>(w) << 7));
    red_shared_inc(lane_base + (byte_of<1>(w) << 7));
    red_shared_inc(lane_base + (byte_of<2>(w) << 7));
    red_shared_inc(lane_base + (byte_of<3>(w) << 7));
}
__device__ __forceinline__ void hist_vec(int4 v, uint32_t lane_base) {
    hist_word((uint32_t)v.x, lane_base);
    hist_word((uint32_t)v.y, lane_base);
    hist_word((uint32_t)v.z, lane_base);
    hist_word((uint32_t)v.w, lane_base);
}

template <int U>
__device__ __forceinline__ void hist_span(const uint8_t* __restrict__ p, size_t n, int tid, int nthr, uint32_t lane_base) {
    const SpanSplit s = split_span(p, n);
    for (uint32_t i = tid; i < s.head; i += nthr) hist_byte(p[i], lane_base);
    const int4* v = reinterpret_cast<const int4*>(p + s.head);
    pipelined_vectors<U>(
        s.nvec, tid, nthr, [&](uint32_t i) { return ld_keep(v + i); }, [&](uint32_t, int4 x) { hist_vec(x, lane_base); });
    for (size_t j = s.tail0 + tid; j < n; j += nthr) hist_byte(p[j], lane_base);
}

// Gather through a byte-replicated lane table (entries are lut * 0x01010101).
__device__ __forceinline__ uint32_t lut_word(uint32_t w, uint32_t lane_base) {
    const uint32_t a = lds_u32(lane_base + (byte_of<0>(w) << 7));
    const uint32_t b = lds_u32(lane_base + (byte_of<1>(w) << 7));
    const uint32_t c = lds_u32(lane_base + (byte_of<2>(w) << 7));
    const uint32_t d = lds_u32(lane_base + (byte_of<3>(w) << 7));
    const uint32_t lo = __byte_perm(a, b, 0x5140);  // a.b0, b.b0 in the low half (all bytes of a, b are equal)
    const uint32_t hi = __byte_perm(c, d, 0x5140);
    return __byte_perm(lo, hi, 0x5410);
}
__device__ __forceinline__ int4 lut_vec(int4 v, uint32_t lane_base) {
    int4 o;
    o.x = (int)lut_word((uint32_t)v.x, lane_base);
    o.y = (int)lut_word((uint32_t)v.y, lane_base);
    o.z = (int)lut_word((uint32_t)v.z, lane_base);
    o.w = (int)lut_word((uint32_t)v.w, lane_base);
    return o;
}
__device__ __forceinline__ uint8_t lut_byte(uint8_t v, uint32_t lane_base) { return (uint8_t)lds_u32(lane_base + ((uint32_t)v << 7)); }

// dst[i] = table[src[i]] over a span.  src and dst must be congruent mod 16 for the vector path; otherwise the
// whole span goes through the byte path (correct, slow, only hit by exotic pointer/pitch combinations).
template <int U>
__device__ __forceinline__ void lut_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                         int nthr, uint32_t lane_base) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = lut_byte(src[i], lane_base);
        return;
    }
    const SpanSplit s = split_span(src, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = lut_byte(src[i], lane_base);
    const int4* v = reinterpret_cast<const int4*>(src + s.head);
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    pipelined_vectors<U>(
        s.nvec, tid, nthr, [&](uint32_t i) { return ld_stream(v + i); },
        [&](uint32_t i, int4 x) { st_stream(o + i, lut_vec(x, lane_base)); });
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = lut_byte(src[j], lane_base);
}

template <int U>
__device__ __forceinline__ void copy_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                          int nthr) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = src[i];
        return;
    }
    const SpanSplit s = split_span(src, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = src[i];
    const int4* v = reinterpret_cast<const int4*>(src + s.head);
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    pipelined_vectors<U>(
        s.nvec, tid, nthr, [&](uint32_t i) { return ld_stream(v + i); }, [&](uint32_t i, int4 x) { st_stream(o + i, x); });
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = src[j];
}

__device__ __forceinline__ void fill_span(uint8_t* __restrict__ dst, size_t n, int tid, int nthr, uint8_t value) {
    const SpanSplit s = split_span(dst, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = value;
    int4* o = reinterpret_cast<int4*>(dst + s.head);
    const int w = (int)(value * 0x01010101u);
    const int4 vv = make_int4(w, w, w, w);
    for (uint32_t i = tid; i < s.nvec; i += nthr) st_stream(o + i, vv);
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
// Sum row `bin` (32 lane columns) of a [256][32] table (call with bin = tid for tid < 256).  Rotated 16-byte reads
// keep the 8 threads of a quarter-warp phase on 8 different bank groups.
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
// Expand a 256-byte LUT (shared memory) into table[v][lane] = lut[v] * 0x01010101.  kThreads / 256 threads share a row.
__device__ __forceinline__ void lane_table_fill_from_lut(uint32_t* tab, const uint8_t* lut256) {
    constexpr int kShare = kThreads / 256, kPer = 8 / kShare;
    const int v = threadIdx.x & 255, part = threadIdx.x >> 8;
    const uint32_t w = (uint32_t)lut256[v] * 0x01010101u;
    uint4* row = reinterpret_cast<uint4*>(tab + v * 32);
    const uint4 q = make_uint4(w, w, w, w);
#pragma unroll
    for (int j = 0; j < kPer; ++j) row[(part * kPer + j + v) & 7] = q;
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

// Spin budget for dependency waits (~2 s): a logic error can never hang the GPU.  Progress is guaranteed by
// construction: an item only waits on items with smaller tickets, and those are resident or finished.
constexpr long long kSpinCycles = 4ll << 30;

// Work tickets.  Thread 0 draws the NEXT ticket at the start of an item and publishes it at the end, so the
// global-atomic round trip is hidden behind the item's work.  Every CTA draws exactly one ticket >= total; the
// CTA that draws the last of those resets the counter for the next launch.
struct TicketQueue {
    uint32_t* counter;
    uint32_t* slots;  // shared uint32[2]
    uint32_t pending;
    uint32_t round;
    __device__ __forceinline__ void start() {
        round = 0;
        if (threadIdx.x == 0) slots[0] = atomicAdd(counter, 1u);
        __syncthreads();
    }
    __device__ __forceinline__ uint32_t current() const { return slots[round & 1]; }
    __device__ __forceinline__ void prefetch() {
        if (threadIdx.x == 0) pending = atomicAdd(counter, 1u);
    }
    // also the end-of-item barrier that protects the shared tables of the next item
    __device__ __forceinline__ void advance() {
        if (threadIdx.x == 0) slots[(round + 1) & 1] = pending;
        ++round;
        __syncthreads();
    }
    __device__ __forceinline__ void finish(uint32_t item, uint32_t total) {
        if (threadIdx.x == 0 && item == total + gridDim.x - 1) atomicExch(counter, 0u);
    }
};

}  // namespace nv12eq
