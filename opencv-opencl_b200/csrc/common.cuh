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

#ifndef NV12EQ_EQ_THREADS
#define NV12EQ_EQ_THREADS 512
#endif
#ifndef NV12EQ_EQ_CTAS
#define NV12EQ_EQ_CTAS 2
#endif
constexpr int kEqCtas = NV12EQ_EQ_CTAS;  // CTAs per SM equalize_kernel is built for
constexpr int kThreads = NV12EQ_EQ_THREADS;  // threads per CTA of equalize_kernel (512 x 2 CTAs/SM: 32 warps, 64 registers/thread)
constexpr int kWarps = kThreads / 32;
static_assert(kThreads == 256 || kThreads == 512 || kThreads == 1024, "lane-table helpers divide 2048 vectors / 256 values evenly over the CTA");
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
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ int4 lds_s4(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint64_t lds_b64(uint32_t saddr) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// L2 residency is managed explicitly (sm_100 LDG/STG carry an L2 eviction priority; SASS: LDG.E.NA.ELL2.256 etc.):
//   first read of the Y plane  -> evict_last  : it is read again by the apply / interpolation pass a few frames later
//   second read, chroma, output-> evict_first : streaming data must not push the Y planes out of the 126 MB L2
// 256-bit accesses (one 32-byte sector-pair per thread) halve the number of load/store instructions.
// The loads are coherent ld.global (no .nc): in-place calls (in == out, supported and tested) store to the very addresses
// the same kernel loads from, and PTX only allows .nc for data that is read-only for the whole kernel.  The L1 bypass and
// the L2 eviction priorities are what matter for speed, not the non-coherent path (A/B: -DNV12EQ_NC='".nc"').
#ifndef NV12EQ_NC
#define NV12EQ_NC ""
#endif
struct V8 {
    uint32_t r[8];
};
__device__ __forceinline__ V8 ldg256_keep(const void* p) {
    V8 v;
    asm volatile("ld.global" NV12EQ_NC ".L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
                 : "l"(p));
    return v;
}
__device__ __forceinline__ V8 ldg256_stream(const void* p) {
    V8 v;
    asm volatile("ld.global" NV12EQ_NC ".L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256_stream(void* p, const V8& v) {
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v.r[0]), "r"(v.r[1]),
                 "r"(v.r[2]), "r"(v.r[3]), "r"(v.r[4]), "r"(v.r[5]), "r"(v.r[6]), "r"(v.r[7]), "l"(p)
                 : "memory");
}
// Pull a line into L2 ahead of its use without tying up a register (row-strided walkers issue this many rows ahead).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 128-bit / 64-bit forms (tile rows and cell rows of CLAHE) take an explicit policy word.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int4 ldg128_hint(const void* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global" NV12EQ_NC ".L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint2 ldg64_hint(const void* p, uint64_t pol) {
    uint2 r;
    asm volatile("ld.global" NV12EQ_NC ".L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg64_hint(void* p, uint2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

// Asynchronous global -> shared copies (LDGSTS): the data never occupies a register while it is in flight, so a
// thread can keep many rows outstanding.  Groups are per thread: wait_group<N> returns when all but the newest N
// committed groups of THIS thread have landed -- a thread that only reads its own slots needs no barrier.
// NOTE: no .L2::cache_hint here.  ptxas 12.9 miscompiles a cache-hinted LDGSTS whose shared address it splits into
// [R + UR + imm]: the emitted instruction names uniform registers that are never written (seen as
// `LDGSTS.E.64 [R9+UR0+0x8000], desc[UR1][...]` with UR0/UR1 undefined) and the kernel dies with "illegal instruction".
// Whether the split happens depends on the surrounding code, so the hint is not worth the risk.
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Byte k of a packed word, zero extended (one PRMT).
template <int K>
__device__ __forceinline__ uint32_t byte_of(uint32_t w) { return __byte_perm(w, 0u, 0x4440u + K); }

// ---- byte-span walkers ----------------------------------------------------------------------------------
// A span is n contiguous bytes.  `tid`/`nthr` select the participating threads (a whole CTA for flat planes, one
// warp for one row of a strided plane).  Unaligned heads/tails are handled with byte accesses.
struct SpanSplit {
    uint32_t head;  // bytes before the first 32-byte aligned address
    uint32_t nvec;  // number of 32-byte vectors (a span is one chunk or one row: far below 2^32 vectors)
    size_t tail0;   // offset of the first tail byte
};
__device__ __forceinline__ SpanSplit split_span(const void* p, size_t n) {
    SpanSplit s;
    const size_t mis = (size_t)((32 - ((uintptr_t)p & 31)) & 31);
    s.head = (uint32_t)(mis < n ? mis : n);
    s.nvec = (uint32_t)((n - s.head) >> 5);
    s.tail0 = (size_t)s.head + ((size_t)s.nvec << 5);
    return s;
}

// Software-pipelined walk over nvec 32-byte vectors, thread `tid` of `nthr` taking vectors tid, tid+nthr, ...
// Three register slots per thread (NV12EQ_PIPE_SLOTS): a slot is refilled (load of the vector three steps ahead) as soon
// as it has been consumed, so a thread always has two or three 32-byte loads in flight while it works: 64-96 KB per SM
// at 1024 threads, which covers the loaded HBM latency at the per-SM share of the bandwidth (two slots: 1 % slower).
#ifndef NV12EQ_PIPE_SLOTS
#define NV12EQ_PIPE_SLOTS 3
#endif
template <class Load, class Use>
__device__ __forceinline__ void pipelined_vectors(uint32_t nvec, int tid, int nthr, Load load, Use use) {
    const uint32_t K = nvec > (uint32_t)tid ? (nvec - tid + nthr - 1) / nthr : 0u;  // vectors of this thread
    const uint32_t step = (uint32_t)nthr;
    uint32_t i = tid;
#if NV12EQ_PIPE_SLOTS == 3
    V8 a, b, c;
#pragma unroll
    for (int j = 0; j < 8; ++j) a.r[j] = b.r[j] = c.r[j] = 0;
    if (K > 0) a = load(i);
    if (K > 1) b = load(i + step);
    if (K > 2) c = load(i + 2 * step);
    uint32_t k = 0;
    for (; k + 3 <= K; k += 3) {
        use(i, a);
        if (k + 3 < K) a = load(i + 3 * step);
        use(i + step, b);
        if (k + 4 < K) b = load(i + 4 * step);
        use(i + 2 * step, c);
        if (k + 5 < K) c = load(i + 5 * step);
        i += 3 * step;
    }
    if (k < K) use(i, a);
    if (k + 1 < K) use(i + step, b);
#else
    V8 a, b;
#pragma unroll
    for (int j = 0; j < 8; ++j) a.r[j] = b.r[j] = 0;
    if (K > 0) a = load(i);
    if (K > 1) b = load(i + step);
    uint32_t k = 0;
    for (; k + 2 <= K; k += 2) {
        use(i, a);
        if (k + 2 < K) a = load(i + 2 * step);
        use(i + step, b);
        if (k + 3 < K) b = load(i + 3 * step);
        i += 2 * step;
    }
    if (k < K) use(i, a);
#endif
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
__device__ __forceinline__ void hist_v8(const V8& v, uint32_t lane_base) {
#pragma unroll
    for (int j = 0; j < 8; ++j) hist_word(v.r[j], lane_base);
}

__device__ __forceinline__ void hist_span(const uint8_t* __restrict__ p, size_t n, int tid, int nthr, uint32_t lane_base) {
    const SpanSplit s = split_span(p, n);
    for (uint32_t i = tid; i < s.head; i += nthr) hist_byte(p[i], lane_base);
    const uint8_t* v = p + s.head;
    pipelined_vectors(
        s.nvec, tid, nthr, [&](uint32_t i) { return ldg256_keep(v + (size_t)i * 32); },
        [&](uint32_t, const V8& x) { hist_v8(x, lane_base); });
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

__device__ __forceinline__ V8 lut_v8(const V8& v, uint32_t lane_base) {
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.r[j] = lut_word(v.r[j], lane_base);
    return o;
}

// dst[i] = table[src[i]] over a span.  src and dst must be congruent mod 32 for the vector path; otherwise the
// whole span goes through the byte path (correct, slow, only hit by exotic pointer/pitch combinations).
__device__ __forceinline__ void lut_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                         int nthr, uint32_t lane_base) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 31) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = lut_byte(src[i], lane_base);
        return;
    }
    const SpanSplit s = split_span(src, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = lut_byte(src[i], lane_base);
    const uint8_t* v = src + s.head;
    uint8_t* o = dst + s.head;
    pipelined_vectors(
        s.nvec, tid, nthr, [&](uint32_t i) { return ldg256_stream(v + (size_t)i * 32); },
        [&](uint32_t i, const V8& x) { stg256_stream(o + (size_t)i * 32, lut_v8(x, lane_base)); });
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = lut_byte(src[j], lane_base);
}

__device__ __forceinline__ void copy_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int tid,
                                          int nthr) {
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 31) != 0) {
        for (size_t i = tid; i < n; i += nthr) dst[i] = src[i];
        return;
    }
    const SpanSplit s = split_span(src, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = src[i];
    const uint8_t* v = src + s.head;
    uint8_t* o = dst + s.head;
    pipelined_vectors(
        s.nvec, tid, nthr, [&](uint32_t i) { return ldg256_stream(v + (size_t)i * 32); },
        [&](uint32_t i, const V8& x) { stg256_stream(o + (size_t)i * 32, x); });
    for (size_t j = s.tail0 + tid; j < n; j += nthr) dst[j] = src[j];
}

__device__ __forceinline__ void fill_span(uint8_t* __restrict__ dst, size_t n, int tid, int nthr, uint8_t value) {
    const SpanSplit s = split_span(dst, n);
    for (uint32_t i = tid; i < s.head; i += nthr) dst[i] = value;
    V8 vv;
#pragma unroll
    for (int j = 0; j < 8; ++j) vv.r[j] = value * 0x01010101u;
    for (uint32_t i = tid; i < s.nvec; i += nthr) stg256_stream(dst + s.head + (size_t)i * 32, vv);
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

// ---- packed fp32 (sm_100 FMUL2 / FADD2): two IEEE round-to-nearest operations per issue slot ------------------
// NOTE: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry .rn, which would break bit-exactness
// with OpenCV's unfused arithmetic.  So products are packed (FMUL2) but the sums that consume them stay scalar
// add.rn.f32 (never fused); only the final magic-number add, whose inputs are sums, is packed again (FADD2).
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void unpack_u2(uint64_t v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul_f2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- mbarrier + TMA (cp.async.bulk.tensor) ---------------------------------------------------------------------
// `bar` / `dst` are shared-window addresses (smem_u32).  One elected thread arms a barrier with the byte count of the boxes it
// requests (arrive.expect_tx); the TMA unit completes the transaction bytes as the boxes land; consumers wait on the phase
// parity.  No registers and no issue slots of the consuming warps are spent on the copy itself.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    // the time hint lets the hardware park the warp until the phase completes (no issue slots burnt on polling)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
// Bounded wait (a lost transaction must not hang the device): false after ~2 s.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > (4ll << 30)) return false;
    return true;
}
// One box of a rank-3 tensor map (x = bytes along a row, y = row, z = frame) into shared memory
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, int x, int y, int z, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar), "l"(policy) : "memory");
}

// 1-D bulk copy global -> shared through the TMA unit (src, dst and size multiples of 16 bytes), with an L2 eviction policy
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// L2 prefetch of one box (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const void* map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
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

// Optional per-item trace (developer tool, tools/trace_items.py): {start ns, dependency-ready ns, end ns, kind | smid<<8}.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t sm_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
// Per-item timestamps (developer tool, tools/trace_items.py).  Compiled in only with -DNV12EQ_ITEM_TRACE=1: the four
// `if (buf && tid == 0)` tests per item were 1.8 % of the 1080p CLAHE kernel's instructions.
#ifndef NV12EQ_ITEM_TRACE
#define NV12EQ_ITEM_TRACE 0
#endif
struct ItemTrace {
    unsigned long long* buf;  // [items][4] or null
    int tid;                  // thread index inside the CTA / work group that processes the item
    __device__ __forceinline__ void mark(uint32_t item, int slot) const {
#if NV12EQ_ITEM_TRACE
        if (buf && tid == 0) buf[(size_t)item * 4 + slot] = global_ns();
#endif
    }
    __device__ __forceinline__ void kind(uint32_t item, uint32_t k) const {
#if NV12EQ_ITEM_TRACE
        if (buf && tid == 0) buf[(size_t)item * 4 + 3] = (unsigned long long)k | ((unsigned long long)sm_id() << 8);
#endif
    }
};

// Work tickets.  Thread 0 draws the NEXT ticket shortly before the end of an item (prefetch(), so that the
// global-atomic round trip is hidden behind the tail of the item's work) and publishes it at the end.  Drawing it
// earlier would queue the next item behind the current one while other CTAs may be idle: with items of very
// different lengths that head-of-line blocking delays the items other CTAs are waiting for.  Every CTA draws exactly
// one ticket >= total and then checks out on an exit counter; the last CTA to check out (all work in the grid is
// finished by then) returns the counters to zero for the next launch -- see the `last_out` result of finish().
struct TicketQueue {
    uint32_t* counter;   // [0] ticket, [2] exit count  (misc workspace words)
    uint32_t* slots;     // shared uint32[2]
    uint32_t pending;
    uint32_t round;
    bool drawn;          // thread 0: the next ticket has been drawn
    __device__ __forceinline__ void start() {
        round = 0;
        drawn = false;
        if (threadIdx.x == 0) slots[0] = atomicAdd(counter, 1u);
        __syncthreads();
    }
    __device__ __forceinline__ uint32_t current() const { return slots[round & 1]; }
    // thread 0 only (other threads: no-op); idempotent within an item
    __device__ __forceinline__ void prefetch() {
        if (threadIdx.x == 0 && !drawn) {
            pending = atomicAdd(counter, 1u);
            drawn = true;
        }
    }
    // also the end-of-item barrier that protects the shared tables of the next item
    __device__ __forceinline__ void advance() {
        prefetch();
        if (threadIdx.x == 0) slots[(round + 1) & 1] = pending;
        drawn = false;
        ++round;
        __syncthreads();
    }
    // Call once, by all threads, when the CTA has no more work.  Returns true (to every thread) in the last CTA.
    __device__ __forceinline__ bool finish() {
        __shared__ int s_last;
        return finish(&s_last);
    }
    // `flag`: one shared int (kernels that own all of their shared memory pass a slot of their dynamic block)
    __device__ __forceinline__ bool finish(int* flag) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const bool last = atomicAdd(counter + 2, 1u) == gridDim.x - 1;
            if (last) {
                counter[0] = 0;
                counter[2] = 0;
            }
            *flag = last;
        }
        __syncthreads();
        return *flag != 0;
    }
};

}  // namespace nv12eq
