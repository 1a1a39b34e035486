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
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"

// Dynamic shared memory of clahe_kernel.  The kernel owns ALL of its shared memory through this one array (no static
// __shared__ variables), so the array starts at a fixed offset of the CTA's shared window: kSmemBase, the 1 KB the system
// reserves at the bottom of the window on sm_90+/sm_100.  Knowing the base at compile time lets every table access be
// `[register + immediate]` with the register formed by ONE PRMT (see row_off): the kernel verifies the assumption when it
// starts and refuses to run otherwise (status word 2 -> NV12EQ_ERR_CUDA on the host).
extern __shared__ __align__(256) uint32_t nv12eq_smem_rows[];

namespace nv12eq {

constexpr int kMaxCells = 4096;  // per axis (tiles + 1); plenty
constexpr int kMaxCellRows = 512;               // rows per interpolation cell (host cuts longer runs)
constexpr int kParamCells = 20;                 // cells per axis that fit the kernel parameters (an 8x8 .. 16x16 grid has 9 .. 17)
// CTA shape of clahe_kernel (compile-time; the Makefile's EXTRA can override for experiments):
//   A CTA is kGroups = 2 independent WORK GROUPS of kCT = 256 threads ("virtual CTAs"): each group draws its own tickets and
//   synchronises on its own named barrier, exactly as a 256-thread CTA would.  Two CTAs per SM = four groups per SM = 32
//   warps at 64 registers per thread: several independent items per SM fill the bubbles of the per-item phases (table build,
//   LUT build, dependency waits, barriers) -- the reason round 1 preferred four small CTAs to two large ones.
//   What the two groups SHARE is one set of 256 table rows of 256 bytes:
//     bytes [g * 128, g * 128 + 128) of row v: group g's hist[v][32 lanes] u32 (tile item) or table[v][16 replicas] of 8-byte
//     entries (cell item).
//   A 256-byte pitch turns a pixel byte into its row offset with ONE PRMT (byte 1 = pixel value, byte 0 = the lane's offset
//   inside the row, which carries the group's half); with 128-byte rows every access needs an extra multiply-add (round 1:
//   24 thread-instructions per pixel, issue-bound).  Four real 256-thread CTAs per SM could not afford 64 KB of rows each.
#ifndef NV12EQ_CLAHE_CTAS
#define NV12EQ_CLAHE_CTAS 2
#endif
#ifndef NV12EQ_CLAHE_SMEM_BASE
#define NV12EQ_CLAHE_SMEM_BASE 1024
#endif
#ifndef NV12EQ_CLAHE_DEPTH
#define NV12EQ_CLAHE_DEPTH 4
#endif
#define NV12EQ_STR2(x) #x
#define NV12EQ_STR(x) NV12EQ_STR2(x)
#define NV12EQ_SBASE NV12EQ_STR(NV12EQ_CLAHE_SMEM_BASE)   // the immediate of every [reg + imm] shared access below
constexpr int kCT = 256;                        // threads of one work group
constexpr int kGroups = 2;                      // work groups per CTA (one per half of the table rows)
constexpr int kBlockThreads = kCT * kGroups;
constexpr int kCWarps = kCT / 32;
constexpr int kClaheCtas = NV12EQ_CLAHE_CTAS;   // CTAs per SM the kernel is built for
constexpr uint32_t kSmemBase = NV12EQ_CLAHE_SMEM_BASE;
constexpr int kRowShift = 8;
constexpr int kRowBytes = 1 << kRowShift;       // pitch of a table row
constexpr int kHalfBytes = kRowBytes / kGroups; // a group's part of every row
constexpr int kCellReps = 16;                   // 8-byte replicas of a cell table entry (conflict-free per half-warp)
constexpr int kRowTableBytes = 256 * kRowBytes;
constexpr int kRingBytesPerThread = 64;         // cp.async ring of the cell loop: 4 x 16 or 8 x 8 bytes in flight per thread
constexpr int kIterDepth = NV12EQ_CLAHE_DEPTH;  // row iterations in flight per thread in the cell loop
#ifndef NV12EQ_CLAHE_G16
#define NV12EQ_CLAHE_G16 1
#endif
constexpr bool kUseG16 = NV12EQ_CLAHE_G16 != 0;
#ifndef NV12EQ_CLAHE_ROWS16
#define NV12EQ_CLAHE_ROWS16 1
#endif
constexpr int kRows16 = NV12EQ_CLAHE_ROWS16;      // rows per loop iteration of the 16-pixel path (1: four iterations in flight, 2: two)   // 16 pixels per thread-row where alignment allows (32 registers of x weights); else 8 pixels x 2 rows
// shared memory map (byte offsets from the start of nv12eq_smem_rows); per-group areas are indexed by the group
constexpr int kRingOff = kRowTableBytes;                                   // [kGroups][kCT * kRingBytesPerThread]
constexpr int kRingGroupBytes = kCT * kRingBytesPerThread;
constexpr int kYwOff = kRingOff + kGroups * kRingGroupBytes;               // [kGroups][kMaxCellRows] float2 (ya1, ya)
constexpr int kScratchOff = kYwOff + kGroups * kMaxCellRows * 8;           // [kGroups][2 * kCWarps] words of the LUT build
constexpr int kMiscOff = kScratchOff + kGroups * 2 * kCWarps * 4;          // [kGroups][kMiscWords]
// Two TMA experiments are kept behind compile-time switches (measured A/B in profiles/r02_clahe_notes.md; both lose):
#ifndef NV12EQ_CLAHE_TMA
#define NV12EQ_CLAHE_TMA 0   // 1: tile rows staged by the TMA unit (cp.async.bulk.tensor + mbarrier stages) instead of the per-thread cp.async ring
#endif
#ifndef NV12EQ_CLAHE_PF
#define NV12EQ_CLAHE_PF 0    // 1: TMA L2 prefetch (cp.async.bulk.prefetch.tensor) of the next item's pixels when its ticket is drawn
#endif
constexpr bool kTmaTiles = NV12EQ_CLAHE_TMA != 0;
constexpr bool kTmaPrefetch = NV12EQ_CLAHE_PF != 0;
constexpr int kStages = 4;                                                 // TMA stages of a tile item (they live in the group's ring bytes)
constexpr int kStageBytes = kRingGroupBytes / kStages;                     // 4 KB: at most one 16-byte piece per thread
constexpr int kMiscWords = 8 + 4 * kStages;                                // ticket slots [2], look-ahead flags [2], flag, last, pad [2]; full / empty mbarriers
constexpr int kClaheSmemBytes = kMiscOff + kGroups * kMiscWords * 4;       // dynamic shared memory of clahe_kernel

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
    int cells_in_params; // nxc, nyc <= kParamCells: the tables below are used (a constant-bank read instead of a global load
                         // at the head of every cell item's dependency chain)
    int4 xc_small[kParamCells], yc_small[kParamCells];
    int uv_chunks;     // chroma items per frame (0 when nothing to do)
    unsigned long long uv_bytes, uv_chunk;  // flat
    int uv_rows_chunk;                      // strided
    int lag;
    // Row-band mode (spatial split of one frame over GPUs, sharding.SpatialSplitClahe): the plane passed in is rows
    // [y_origin, y_origin + h) of a taller frame.  tile_items = 0 runs the interpolation only, on a caller-supplied LUT grid that
    // holds one halo row of tile LUTs above and below the band's own (y cells index into it); cells_off = 1 builds the LUTs only.
    int tile_items;        // tile items per frame: tx * ty, or 0 (apply only)
    int cells_off;         // 1: no cell and no uv items
    int lut_tiles;         // LUT tables per frame in `luts`
    int y_origin;          // frame row of the plane's first row (the y weights are those of the whole frame)
    // interpolation weights, computed once per geometry on the host with the device's operation order (separately rounded
    // multiply, subtract, floor, subtract): xw[x] = xa(x) * kXScale, yw[y] = {(1 - ya(y)) * kYScale, ya(y) * kYScale}
    const float* xw;       // [w]
    const float2* yw;      // [frame height]
    uint8_t* luts;         // [n_frames][lut_tiles][256]
    uint32_t* tiles_done;  // [n_frames] published tile LUTs per frame; self-cleaned
    uint32_t* ticket;      // [1] self-cleaned
    uint32_t* status;      // [1]
    unsigned long long* trace;  // optional [items][4] (developer tool)
    // TMA staging of the tile rows (tile items, NV12EQ_CLAHE_TMA): rank-3 maps (x bytes, row, frame) over the input batch with
    // boxes of tma_bw x tma_bh bytes ([0]) and tma_bw x tma_bh_tail ([1], last stage of a tile when th % tma_bh != 0)
    alignas(64) CUtensorMap tile_map[2];
    int tma_tiles;              // 0: per-thread cp.async ring (any alignment / padding)
    int tma_bw, tma_nb;         // box width in bytes, boxes per row block (tw = tma_bw * tma_nb)
    int tma_bh, tma_bh_tail;    // rows per stage, rows of the last stage
    int tma_nst;                // stages per tile
    // L2 prefetch of the next item's pixels by the TMA unit (NV12EQ_CLAHE_PF): boxes of pf_bw x pf_bh bytes of the same rank-3
    // view, issued by the group's thread 0 when it draws the next ticket, one item ahead of the loads
    alignas(64) CUtensorMap pf_map;
    int pf_on, pf_bw, pf_bh;
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

// ---- shared memory accessed as [register + kSmemBase] ----------------------------------------------------------
// `off` is a byte offset from the start of nv12eq_smem_rows.
#ifdef NV12EQ_CLAHE_ATOMS_ADD
// variant: the increment in a register (SASS ATOMS.ADD instead of ATOMS.POPC.INC)
__device__ __forceinline__ void red_inc_rel(uint32_t off) {
    uint32_t one;
    asm("mov.u32 %0, %%nctaid.y;" : "=r"(one));   // 1 for this kernel's 1-D grid, unknown to ptxas; not volatile: hoisted out of the loops
    asm volatile("red.shared.add.u32 [%0+" NV12EQ_SBASE "], %1;" ::"r"(off), "r"(one) : "memory");
}
#else
__device__ __forceinline__ void red_inc_rel(uint32_t off) { asm volatile("red.shared.add.u32 [%0+" NV12EQ_SBASE "], 1;" ::"r"(off) : "memory"); }
#endif
__device__ __forceinline__ uint2 lds64_rel(uint32_t off) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+" NV12EQ_SBASE "];" : "=r"(v.x), "=r"(v.y) : "r"(off));
    return v;
}
__device__ __forceinline__ uint64_t lds_b64_rel(uint32_t off) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1+" NV12EQ_SBASE "];" : "=l"(v) : "r"(off));
    return v;
}
__device__ __forceinline__ int4 lds128_rel(uint32_t off) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4+" NV12EQ_SBASE "];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(off));
    return v;
}
__device__ __forceinline__ void cp_async8_rel(uint32_t off, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0+" NV12EQ_SBASE "], [%1], 8;" ::"r"(off), "l"(g) : "memory");
}
#ifndef NV12EQ_CLAHE_L2HINT
#define NV12EQ_CLAHE_L2HINT 1   // tile pass loads carry L2::evict_last (the cell pass reads the plane again `lag` frames later)
#endif
__device__ __forceinline__ void cp_async16_rel_hint(uint32_t off, const void* g, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0+" NV12EQ_SBASE "], [%1], 16, %2;" ::"r"(off), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_rel(uint32_t off, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0+" NV12EQ_SBASE "], [%1], 16;" ::"r"(off), "l"(g) : "memory");
}
// ---- 256-byte-pitch shared tables ------------------------------------------------------------------------------
// Row v of the table is at byte offset v << 8.  `lane_off` is this lane's byte offset inside a row (< 128, bytes 1..3 of
// the register zero): byte K of a packed pixel word goes to byte 1, the lane offset stays in byte 0 -- one PRMT is the
// whole offset, the base is the instruction's immediate.
template <int K>
__device__ __forceinline__ uint32_t row_off(uint32_t w, uint32_t lane_off) { return __byte_perm(w, lane_off, 0x5504u | (K << 4)); }

// histogram rows: hist[bin][lane] u32 in the first 128 bytes of row `bin`
__device__ __forceinline__ void hist256_byte(uint32_t v, uint32_t lane4) { red_inc_rel((v << kRowShift) + lane4); }
__device__ __forceinline__ void hist256_word(uint32_t w, uint32_t lane4) {
    red_inc_rel(row_off<0>(w, lane4));
    red_inc_rel(row_off<1>(w, lane4));
    red_inc_rel(row_off<2>(w, lane4));
    red_inc_rel(row_off<3>(w, lane4));
}
__device__ __forceinline__ void hist256_vec(int4 v, uint32_t lane4) {
    hist256_word((uint32_t)v.x, lane4);
    hist256_word((uint32_t)v.y, lane4);
    hist256_word((uint32_t)v.z, lane4);
    hist256_word((uint32_t)v.w, lane4);
}
// n contiguous bytes by one warp (general tile path): 16-byte vectors where alignment allows, bytes elsewhere
__device__ __forceinline__ void hist256_span_warp(const uint8_t* __restrict__ p, int n, int lane, uint32_t lane4, uint64_t pol) {
    const int mis = (int)((16 - ((uintptr_t)p & 15)) & 15);
    const int head = min(mis, n);
    for (int i = lane; i < head; i += 32) hist256_byte(p[i], lane4);
    const int nvec = (n - head) >> 4;
    const uint8_t* v = p + head;
    for (int i = lane; i < nvec; i += 32) hist256_vec(ldg128_hint(v + (size_t)i * 16, pol), lane4);
    for (int i = head + (nvec << 4) + lane; i < n; i += 32) hist256_byte(p[i], lane4);
}
// named barrier of one work group (ids 1..kGroups; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kCT) : "memory"); }

// `half` = rows + group * kHalfBytes: the group's 128 bytes of every row
__device__ __forceinline__ void hist256_zero(uint8_t* half, int tid) {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 2048 / kCT; ++k) {
        const int i = tid + k * kCT;  // 16-byte slot i of the 128 counter bytes of every row: row i>>3, column i&7
        *reinterpret_cast<uint4*>(half + (i >> 3) * kRowBytes + (i & 7) * 16) = z;
    }
}
__device__ __forceinline__ uint32_t hist256_row_sum(const uint8_t* half, int bin) {
    const uint8_t* row = half + bin * kRowBytes;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 q = *reinterpret_cast<const uint4*>(row + ((j + bin) & 7) * 16);   // 256-byte pitch: every row starts in bank 0, rotate by the row
        s += q.x + q.y + q.z + q.w;
    }
    return s;
}

__device__ __forceinline__ int4 cell_x(const ClaheParams& p, int cx) { return p.cells_in_params ? p.xc_small[cx] : p.xcells[cx]; }
__device__ __forceinline__ int4 cell_y(const ClaheParams& p, int cy) { return p.cells_in_params ? p.yc_small[cy] : p.ycells[cy]; }

// Work tickets of one work group (see TicketQueue in common.cuh for the protocol): the group's thread 0 draws the next ticket
// shortly before the end of an item and publishes it through the group's shared slots at the group's barrier.
// Look-ahead: when the ticket just drawn is a cell item, thread 0 also reads (acquire) the tile counter of that item's frame.
// In steady state the tiles are long complete, so the next item starts with the answer in hand: no polling round trip and no
// extra barrier at the head of the cell item's dependency chain.
struct GroupTickets {
    uint32_t* counter;   // [0] ticket, [2] exit count  (misc workspace words)
    uint32_t* slots;     // shared uint32[4] of this group: tickets [2], look-ahead flags [2]
    int tid, bar;
    // decoding of a ticket (same values the kernel uses)
    const ClaheParams& p;
    uint32_t total_items, per_slot;
    int T, I;
    uint32_t pending, pending_ready;
    uint32_t round;
    bool drawn;          // thread 0: the next ticket has been drawn
    __device__ __forceinline__ void start() {
        round = 0;
        drawn = false;
        pending_ready = 0;
        if (tid == 0) {
            slots[0] = atomicAdd(counter, 1u);
            slots[2] = 0;
        }
        group_sync(bar);
    }
    __device__ __forceinline__ uint32_t current() const { return slots[round & 1]; }
    __device__ __forceinline__ bool current_ready() const { return slots[2 + (round & 1)] != 0; }
    __device__ __forceinline__ void prefetch() {   // thread 0 only (other threads: no-op); idempotent within an item
        if (tid == 0 && !drawn) {
            pending = atomicAdd(counter, 1u);
            drawn = true;
            pending_ready = 0;
            if (pending < total_items) {
                const uint32_t g = p.slot_magic ? __umulhi(pending, p.slot_magic) : pending / per_slot;
                const int r = (int)(pending - g * per_slot), f = (int)g - p.lag;
                if (r >= T && r < T + I && f >= 0) pending_ready = ld_acquire_u32(p.tiles_done + f) >= (uint32_t)T;
                if (kTmaPrefetch && p.pf_on) {
                    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, z = -1;
                    if (r < T) {
                        if ((int)g < p.n_frames) {
                            const int tyi = small_div(r, p.tx), txi = r - tyi * p.tx;
                            x0 = txi * p.tw; y0 = tyi * p.th; x1 = x0 + p.tw; y1 = y0 + p.th; z = (int)g;
                        }
                    } else if (r < T + I && f >= 0) {
                        const int ci = r - T;
                        const int cy = small_div(ci, p.nxc), cx = ci - cy * p.nxc;
                        const int4 xc = p.cells_in_params ? p.xc_small[cx] : p.xcells[cx], yc = p.cells_in_params ? p.yc_small[cy] : p.ycells[cy];
                        x0 = xc.x; x1 = xc.y; y0 = yc.x; y1 = yc.y; z = f;
                    }
                    if (z >= 0)
                        for (int y = y0; y < y1; y += p.pf_bh)
                            for (int x = x0; x < x1; x += p.pf_bw) tma_prefetch_3d(&p.pf_map, x, y, z);
                }
            }
        }
    }
    __device__ __forceinline__ void advance() {    // also the end-of-item barrier that protects the group's tables
        prefetch();
        if (tid == 0) {
            slots[(round + 1) & 1] = pending;
            slots[2 + ((round + 1) & 1)] = pending_ready;
        }
        drawn = false;
        ++round;
        group_sync(bar);
    }
    // Call once, by all threads of the group, when it has no more work.  True (to every thread) in the last group of the grid.
    __device__ __forceinline__ bool finish(int* flag) {
        group_sync(bar);
        if (tid == 0) {
            __threadfence();
            const bool last = atomicAdd(counter + 2, 1u) == gridDim.x * kGroups - 1;
            if (last) {
                counter[0] = 0;
                counter[2] = 0;
            }
            *flag = last;
        }
        group_sync(bar);
        return *flag != 0;
    }
};

// One tile's pixel rows -> histogram.  Threads form a (rows per pass x 16-byte pieces per row) grid over the tile; a thread
// walks down its column of pieces, staged through its slots of the group's cp.async ring (kRingBytesPerThread / 16 - 1 pieces
// in flight without holding registers).  Loading the pieces straight into registers instead (256-bit or 128-bit LDG with
// L2::evict_last, four deep) measured 5 % slower per frame on B200 (profiles/r02_clahe_notes.md): the tile pass is bound by the
// shared-atomic rate (~16 lanes per clock per SM), and what matters beside it is that waiting warps cost no registers.
struct TileRowsRing {
    static constexpr int D = kRingBytesPerThread / 16;
    static constexpr uint32_t kSlotStride = kCT * 16;
    const uint8_t* pn;
    size_t rstep;
    int nrows;
    uint32_t ring0;
    uint64_t pol;   // L2 eviction policy of the loads
    __device__ __forceinline__ void load(uint32_t off, const uint8_t* g) const {
#if NV12EQ_CLAHE_L2HINT
        cp_async16_rel_hint(off, g, pol);
#else
        cp_async16_rel(off, g);
#endif
    }
    __device__ __forceinline__ void start(const uint8_t* __restrict__ ptr, size_t rstep_, int nrows_, int tid, int group, uint64_t pol_) {
        rstep = rstep_; nrows = nrows_; ring0 = (uint32_t)(kRingOff + group * kRingGroupBytes + tid * 16); pol = pol_;
#pragma unroll
        for (int j = 0; j < D - 1; ++j) {
            if (j < nrows) load(ring0 + (uint32_t)j * kSlotStride, ptr + (size_t)j * rstep);
            cp_async_commit();
        }
        pn = ptr + (size_t)(D - 1) * rstep;
    }
    __device__ __forceinline__ void run(uint32_t lane4, GroupTickets& q) {
#pragma unroll 1
        for (int i0 = 0; i0 < nrows; i0 += D) {
            if (i0 + 2 * D >= nrows) q.prefetch();
#pragma unroll
            for (int j = 0; j < D; ++j) {
                if (i0 + j < nrows) {
                    if (i0 + j + D - 1 < nrows) load(ring0 + (uint32_t)((j + D - 1) % D) * kSlotStride, pn);
                    cp_async_commit();
                    cp_async_wait<D - 1>();
                    hist256_vec(lds128_rel(ring0 + (uint32_t)j * kSlotStride), lane4);
                    pn += rstep;
                }
            }
        }
    }
};

// Tile histogram -> clip -> redistribute -> scan -> LUT bytes (global), by the whole CTA: thread t < 256 owns bin t
// (count = its histogram value).  The per-bin work (clip, residual test, conversion) runs 256 wide and the two reductions
// cost one barrier each; a single-warp version kept the other warps of the CTA waiting ~1 us per tile.
// s_scratch: 2 * kCWarps words.
__device__ __forceinline__ void clahe_tile_lut_block(uint32_t count, int clip_limit, float lut_scale, uint8_t* __restrict__ glut,
                                                     uint32_t* s_scratch, int tid, int bar) {
    const int lane = tid & 31, warp = tid >> 5;
    int h = tid < 256 ? (int)count : 0;
    if (clip_limit > 0) {   // uniform over the CTA
        int excess = max(h - clip_limit, 0);
        h = min(h, clip_limit);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, d);
        if (lane == 0) s_scratch[warp] = (uint32_t)excess;
        group_sync(bar);
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
    group_sync(bar);
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
#ifndef NV12EQ_CLAHE_UNPACK_PRMT
#define NV12EQ_CLAHE_UNPACK_PRMT 1
#endif
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint64_t pack_u2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ uint64_t sub_f2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// bf16 pair -> packed fp32 pair.  With PRMT both halves are alu-pipe operations; left to the compiler, the low half becomes
// an IMAD.U32 on the fma pipe, which the blend's FMUL2 / FADD2 already load the most (tools/blend_bench.cu, pipe_rates.cu:
// f32x2 operations and all alu-pipe operations take two cycles per warp on a sub-partition, scalar fp32 one).
__device__ __forceinline__ uint64_t bf16x2_to_f2(uint32_t e) {
#if NV12EQ_CLAHE_UNPACK_PRMT
    return pack_u2(prmt_b32(e, 0u, 0x1044u), prmt_b32(e, 0u, 0x3244u));
#else
    return pack_f2(__uint_as_float(e << 16), __uint_as_float(e & 0xffff0000u));
#endif
}
// Table entry of one pixel value.  Two formats:
//   bf16 pairs (8 bytes, 16 replicas per row): {bf16 L11 | bf16 L21 << 16, bf16 L12 | bf16 L22 << 16}
//   bytes      (4 bytes, 32 replicas per row): L11 | L21 << 8 | L12 << 16 | L22 << 24.  A byte moved to the low end of a zero
//              word IS the fp32 subnormal L * 2^-149; the x weights carry 2^100 and the y weights 2^49, so every product and
//              sum is the reference's value times an exact power of two (same significand, same rounding; all intermediate
//              values are normal or zero: L * xa * 2^-49 >= 2^-75) and the final sum comes out unscaled.  One shared
//              wavefront per warp-wide gather instead of two (tools/blend_bench.cu checks both against each other).
#ifndef NV12EQ_CLAHE_BYTE_TABLE
#define NV12EQ_CLAHE_BYTE_TABLE 1
#endif
constexpr bool kByteTable = NV12EQ_CLAHE_BYTE_TABLE != 0;
constexpr float kXScale = kByteTable ? 1.2676506002282294e30f /* 2^100 */ : 1.0f;
constexpr float kYScale = kByteTable ? 5.62949953421312e14f /* 2^49 */ : 1.0f;
struct TableEntry {
    uint32_t x, y;   // bytes format: x only
};
__device__ __forceinline__ TableEntry lds_entry_rel(uint32_t off) {
    TableEntry e;
    if (kByteTable) {
        asm volatile("ld.shared.u32 %0, [%1+" NV12EQ_SBASE "];" : "=r"(e.x) : "r"(off));
        e.y = 0;
    } else {
        const uint2 v = lds64_rel(off);
        e.x = v.x; e.y = v.y;
    }
    return e;
}
__device__ __forceinline__ float clahe_blend_res(TableEntry e, float xa, float xa1, uint64_t yw) {
    const uint64_t A = kByteTable ? pack_u2(prmt_b32(e.x, 0u, 0x4440u), prmt_b32(e.x, 0u, 0x4441u)) : bf16x2_to_f2(e.x);  // (L11, L21)
    const uint64_t B = kByteTable ? pack_u2(prmt_b32(e.x, 0u, 0x4442u), prmt_b32(e.x, 0u, 0x4443u)) : bf16x2_to_f2(e.y);  // (L12, L22)
    const uint64_t S = add_f2_nofuse(mul_f2(A, pack_f2(xa1, xa1)), mul_f2(B, pack_f2(xa, xa)));   // (top, bot)
    float r0, r1;
    unpack_f2(mul_f2(S, yw), r0, r1);  // (top*ya1, bot*ya)
    return __fadd_rn(r0, r1);
}
template <int K>
__device__ __forceinline__ float clahe_blend_px(uint32_t w, uint32_t lane8, float xa, float xa1, uint64_t yw) {
    return clahe_blend_res(lds_entry_rel(row_off<K>(w, lane8)), xa, xa1, yw);
}
__device__ __forceinline__ void round_pair(float a, float b, uint32_t& oa, uint32_t& ob) {
    unpack_u2(add_f2(pack_f2(a, b), pack_f2(12582912.0f, 12582912.0f)), oa, ob);
}
// low bytes of four words -> one packed word (3 PRMT)
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// Pixels 2P and 2P+1 of word q of R rows: the x weights of a pixel pair travel as one packed register pair, and the pair
// (1 - xa) is ONE packed subtraction shared by the R rows -- the compiler otherwise re-derives 1 - xa per pixel and row (a
// 64-register kernel cannot hold 32 x weights) with scalar FADDs.
template <int P, int R>
__device__ __forceinline__ void clahe_blend_pair(const uint32_t (&w)[R], uint32_t lane8, uint64_t xap, const uint64_t (&yw)[R], uint32_t (&o)[R][4]) {
    const uint64_t xa1p = sub_f2(pack_f2(kXScale, kXScale), xap);   // (1 - xa) * kXScale, exactly
    float xa0, xa1, xb0, xb1;
    unpack_f2(xap, xa0, xa1);
    unpack_f2(xa1p, xb0, xb1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float f0 = clahe_blend_px<2 * P>(w[r], lane8, xa0, xb0, yw[r]);
        const float f1 = clahe_blend_px<2 * P + 1>(w[r], lane8, xa1, xb1, yw[r]);
        round_pair(f0, f1, o[r][2 * P], o[r][2 * P + 1]);
    }
}
// NW packed words (4 pixels each) of R rows that share their x weights xw[2 * NW] (pixel pairs)
template <int NW, int R>
__device__ __forceinline__ void clahe_blend_rows(const uint32_t (&px)[R][NW], uint32_t lane8, const uint64_t* xw, const uint64_t (&yw)[R], uint32_t (&out)[R][NW]) {
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        uint32_t w[R], o[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = px[r][q];
        clahe_blend_pair<0, R>(w, lane8, xw[2 * q], yw, o);
        clahe_blend_pair<1, R>(w, lane8, xw[2 * q + 1], yw, o);
#pragma unroll
        for (int r = 0; r < R; ++r) out[r][q] = pack_low_bytes(o[r][0], o[r][1], o[r][2], o[r][3]);
    }
}

__device__ __forceinline__ void axis_weight(int pos, float inv, float& a, float& a1) {
    const float f = __fsub_rn(__fmul_rn((float)pos, inv), 0.5f);
    const float t1 = floorf(f);
    a = __fsub_rn(f, t1);
    a1 = __fsub_rn(1.0f, a);
}

// The rows of one thread inside a cell: G (8 or 16) horizontally adjacent pixels per row, rows tr, tr + rpp, ...; R of
// them per loop iteration (the x weights are shared by all rows, so R = 2 halves the per-iteration overhead and the
// re-derivation of 1 - xa per pixel).  Thread-private ring of D * R G-byte slots, filled with cp.async: the rows of D - 1
// iterations ahead are in flight without holding registers, and since a thread only ever reads its own slots no barrier is
// involved.  The loop is unrolled by D, so every ring slot is a compile-time offset.  Thread 0 (tr == 0, the most rows) draws
// the next ticket at the start of the last round, which hides the atomic's round trip.
template <int G, int R, int D>
struct CellRows {
    static_assert(D * R * G <= kRingBytesPerThread, "ring must fit the thread's slots");
    static constexpr int NW = G / 4;   // packed pixel words per row
    uint64_t xw[G / 2];                // x weights (xa) of pixel pairs
    const uint8_t* spn;
    uint8_t* dp;
    size_t rstep;
    uint32_t ring0, yw_off, yw_step;
    int nrows;

    // Slot j of thread t is at ring0 + j * kSlotStride with ring0 = group ring + t * G: consecutive threads sit G bytes
    // apart, so a warp-wide access is one contiguous span (no bank conflicts) and every slot is the instruction's immediate.
    static constexpr uint32_t kSlotStride = kCT * G;
    static __device__ __forceinline__ uint32_t ring_base(int tid, int group) {
        return (uint32_t)(kRingOff + group * kRingGroupBytes + tid * G);
    }
    __device__ __forceinline__ void issue(int slot, const uint8_t* g) const {
        if (G == 16) cp_async16_rel(ring0 + (uint32_t)slot * kSlotStride, g); else cp_async8_rel(ring0 + (uint32_t)slot * kSlotStride, g);
    }
    // Everything that does not depend on the tile LUTs: x weights and the first D-1 iterations of the ring.  Called BEFORE
    // the dependency wait and the table build, so the pixel loads are in flight while the CTA waits for / packs the LUTs.
    __device__ __forceinline__ void start(const uint8_t* sp, uint8_t* dp_, size_t rstep_, int nrows_, const float* xwp, int tid, int group,
                                          uint32_t yw_off_, uint32_t yw_step_) {
        dp = dp_; rstep = rstep_; nrows = nrows_; ring0 = ring_base(tid, group); yw_off = yw_off_; yw_step = yw_step_;
        if (nrows > 0) {   // xwp is 32-byte aligned (the group's first column is a multiple of 8) and inside the table
#pragma unroll
            for (int k = 0; k < G; k += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xwp + k));
                xw[k / 2] = pack_f2(v.x, v.y);
                xw[k / 2 + 1] = pack_f2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int k = 0; k < G / 2; ++k) xw[k] = 0ull;
        }
#pragma unroll
        for (int j = 0; j < D - 1; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (j * R + r < nrows) issue(j * R + r, sp + (size_t)(j * R + r) * rstep);
            cp_async_commit();
        }
        spn = sp + (size_t)((D - 1) * R) * rstep;
    }
    __device__ __forceinline__ void run(GroupTickets& q, uint32_t lane8) {
#pragma unroll 1
        for (int i0 = 0; i0 < nrows; i0 += D * R) {
            if (i0 + D * R >= nrows) q.prefetch();
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const int i = i0 + j * R;
                if (i < nrows) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (i + (D - 1) * R + r < nrows) issue(((j + D - 1) % D) * R + r, spn + (size_t)r * rstep);
                    cp_async_commit();
                    cp_async_wait<D - 1>();
                    // a row past the end of the cell (odd row count, R = 2) is blended from stale ring bytes and not stored
                    uint32_t px[R][NW], out[R][NW];
                    uint64_t yw[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        yw[r] = (r == 0 || i + r < nrows) ? lds_b64_rel(yw_off + (uint32_t)r * yw_step) : 0ull;   // (the weight table ends with the cell)
                        if (G == 16) {
                            const int4 v = lds128_rel(ring0 + (uint32_t)(j * R + r) * kSlotStride);
                            px[r][0] = (uint32_t)v.x; px[r][1] = (uint32_t)v.y; px[r][NW - 2] = (uint32_t)v.z; px[r][NW - 1] = (uint32_t)v.w;
                        } else {
                            const uint2 v = lds64_rel(ring0 + (uint32_t)(j * R + r) * kSlotStride);
                            px[r][0] = v.x; px[r][NW - 1] = v.y;
                        }
                    }
                    clahe_blend_rows<NW, R>(px, lane8, xw, yw, out);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r == 0 || i + r < nrows) {
                            uint8_t* d = dp + (size_t)r * rstep;
                            if (G == 16) __stcs(reinterpret_cast<uint4*>(d), make_uint4(out[r][0], out[r][1], out[r][NW - 2], out[r][NW - 1]));
                            else __stcs(reinterpret_cast<uint2*>(d), make_uint2(out[r][0], out[r][NW - 1]));
                        }
                    }
                    spn += (size_t)R * rstep; dp += (size_t)R * rstep; yw_off += (uint32_t)R * yw_step;
                }
            }
        }
    }
};

template <int MIN_CTAS>
__global__ void __launch_bounds__(kBlockThreads, MIN_CTAS) clahe_kernel(const __grid_constant__ ClaheParams p) {
    uint8_t* const rows = reinterpret_cast<uint8_t*>(nv12eq_smem_rows);   // see the shared memory map above
    const int group = threadIdx.x / kCT, tid = threadIdx.x - group * kCT, lane = tid & 31, warp = tid >> 5;
    const int bar = 1 + group;
    uint8_t* const half = rows + group * kHalfBytes;                       // this group's 128 bytes of every table row
    float2* const s_yw = reinterpret_cast<float2*>(rows + kYwOff) + group * kMaxCellRows;
    uint32_t* const s_scratch = reinterpret_cast<uint32_t*>(rows + kScratchOff) + group * 2 * kCWarps;
    uint32_t* const s_ticket = reinterpret_cast<uint32_t*>(rows + kMiscOff) + group * kMiscWords;
    int* const s_flag = reinterpret_cast<int*>(s_ticket + 4);
    int* const s_last = reinterpret_cast<int*>(s_ticket + 5);

    if (smem_u32(rows) != kSmemBase) {   // the [reg + imm] accesses below would miss the tables: refuse loudly
        if (threadIdx.x == 0) atomicExch(p.status, 2u);
        return;
    }
    const int T = p.tile_items;
    const int I = p.cells_off ? 0 : p.nxc * p.nyc;
    const int U = p.cells_off ? 0 : p.uv_chunks;
    const int per_slot = T + I + U;
    const uint32_t total_items = (uint32_t)(p.n_frames + p.lag) * (uint32_t)per_slot;
    const uint32_t lane4 = (uint32_t)(group * kHalfBytes + lane * 4);                        // hist column of this lane
    const uint32_t lane8 = kByteTable ? (uint32_t)(group * kHalfBytes + lane * 4)                               // cell table replica of this lane
                                      : (uint32_t)(group * kHalfBytes + (lane & (kCellReps - 1)) * 8);
    const uint32_t yw_base = (uint32_t)(kYwOff + group * kMaxCellRows * 8);

    // mbarriers of the group's TMA stages: full[s] (armed by thread 0 with the stage's bytes), empty[s] (one arrival per warp)
    const uint32_t bar_full = kSmemBase + (uint32_t)(kMiscOff + (group * kMiscWords + 8) * 4);
    const uint32_t bar_empty = bar_full + kStages * 8;
    const uint32_t stage_base = (uint32_t)(kRingOff + group * kRingGroupBytes);     // offset from the start of the dynamic array
    if (kTmaTiles && tid == 0) {
#pragma unroll
        for (int s_ = 0; s_ < kStages; ++s_) {
            mbar_init(bar_full + s_ * 8, 1);
            mbar_init(bar_empty + s_ * 8, kCWarps);
        }
        fence_mbar_init();
    }
    uint32_t tk = 0;   // TMA stage uses of this group since the kernel started (same value in all of its threads)
    GroupTickets q{p.ticket, s_ticket, tid, bar, p, total_items, (uint32_t)per_slot, T, I, 0u, 0u, 0u, false};
    q.start();
    int publish = -1;   // thread 0: frame whose tile counter still has to be bumped for the tile item just finished
    // The LUT bytes of a tile item are stored by all threads before the item's closing barrier (q.advance); thread 0 then fences
    // and counts the tile at the top of the next iteration (the cooperative-groups grid-sync pattern: barrier, then one thread
    // fences and signals), while the other warps already work on the next item.
    // (Per-TILE flags, so that a cell only waits for its own four tiles, measured 2 % slower: four flag reads per look-ahead.)
    auto publish_tile = [&]() {
        if (tid == 0 && publish >= 0) {
            __threadfence();
            atomicAdd(p.tiles_done + publish, 1u);
            publish = -1;
        }
    };
    for (;;) {
        const uint32_t item = q.current();
        publish_tile();
        if (item >= total_items) break;
        // item = g * per_slot + r; the host supplies floor(2^32 / per_slot) + 1 when that multiplier divides exactly
        const int g = p.slot_magic ? (int)__umulhi(item, p.slot_magic) : (int)(item / (uint32_t)per_slot);
        const int r = (int)(item - (uint32_t)g * (uint32_t)per_slot);
        const int f = g - p.lag;
        const ItemTrace tr_{p.trace, tid};
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
                const uintptr_t align_or = (uintptr_t)y + (uintptr_t)x0;
                const bool vec16 = !(p.debug_skip & 1) && !p.padded && (p.tw & 15) == 0 && (p.stride & 15) == 0 && (align_or & 15) == 0 &&
                                   (p.tw >> 4) <= kCT;
                                // threads form a (rows_per_pass x vectors_per_row) grid over the tile
                const int vpr = vec16 ? (p.tw >> 4) : 1;
                const int rpp = small_div(kCT, vpr);
                const int tr = small_div(tid, vpr), tc = tid - tr * vpr;
                const int nrows = (vec16 && tr < rpp && tr < p.th) ? small_div(p.th - tr + rpp - 1, rpp) : 0;
                const uint8_t* ptr = y + (size_t)(y0 + tr) * p.stride + x0 + tc * 16;
                const size_t rstep = (size_t)rpp * p.stride;
                bool tma_ok = true;
                if (p.debug_skip & 1) {
                    hist256_zero(half, tid);
                    group_sync(bar);
                } else if (kTmaTiles && p.tma_tiles) {
                    // Rows staged by the TMA unit: stage k = rows [k * bh, k * bh + bh) of the tile as tma_nb boxes, kStages - 1
                    // stages in flight.  A stage is a run of 16-byte pieces (any piece order gives the same histogram): thread t
                    // takes piece pc of box bx in every stage.
                    const int nst = p.tma_nst;
                    const uint32_t tk0 = tk;
                    const uint32_t box_slot = (uint32_t)(p.tma_bw * p.tma_bh + 127) & ~127u;   // a box lands on a 128-byte boundary
                    auto issue_stage = [&](int k) {   // thread 0
                        const uint32_t s_ = (tk0 + (uint32_t)k) % kStages;
                        const bool tail = (k == nst - 1) && p.tma_bh_tail != p.tma_bh;
                        const int bh = tail ? p.tma_bh_tail : p.tma_bh;
                        mbar_arrive_expect_tx(bar_full + s_ * 8, (uint32_t)(p.tma_bw * bh * p.tma_nb));
                        for (int b = 0; b < p.tma_nb; ++b)
                            tma_load_3d(kSmemBase + stage_base + s_ * kStageBytes + (uint32_t)b * box_slot, &p.tile_map[tail ? 1 : 0], x0 + b * p.tma_bw,
                                        y0 + k * p.tma_bh, g, bar_full + s_ * 8, keep);
                    };
                    if (tid == 0) {
                        fence_proxy_async_smem();   // the ring bytes were last written by cp.async (generic proxy)
                        for (int k = 0; k < kStages - 1 && k < nst; ++k) issue_stage(k);
                    }
                    hist256_zero(half, tid);
                    group_sync(bar);
                    const int ppb_full = (p.tma_bw * p.tma_bh) >> 4, ppb_tail = (p.tma_bw * p.tma_bh_tail) >> 4;   // pieces per box
                    const int bx_f = small_div(tid, ppb_full), bx_t = small_div(tid, ppb_tail);
                    const uint32_t off_full = (uint32_t)bx_f * box_slot + (uint32_t)(tid - bx_f * ppb_full) * 16u;
                    const uint32_t off_tail = (uint32_t)bx_t * box_slot + (uint32_t)(tid - bx_t * ppb_tail) * 16u;
#pragma unroll 1
                    for (int k = 0; k < nst; ++k) {
                        const uint32_t s_ = tk % kStages, par = (tk / kStages) & 1u;
                        if (k + 2 >= nst) q.prefetch();
                        if (!mbar_wait(bar_full + s_ * 8, par)) { tma_ok = false; break; }
                        const bool last = k == nst - 1;
                        if ((last ? bx_t : bx_f) < p.tma_nb) hist256_vec(lds128_rel(stage_base + s_ * kStageBytes + (last ? off_tail : off_full)), lane4);
                        // The increments above could not issue before their piece arrived in registers, and this arrival issues after
                        // them: the stage may be overwritten once every warp has arrived.
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + s_ * 8);
                        if (tid == 0 && k + kStages - 1 < nst) {
                            // refill the stage of the previous iteration (all warps are done with it, or about to be)
                            if (k >= 1 && !mbar_wait(bar_empty + ((tk - 1) % kStages) * 8, ((tk - 1) / kStages) & 1u)) tma_ok = false;
                            else issue_stage(k + kStages - 1);
                        }
                        ++tk;
                    }
                } else if (vec16) {
                    TileRowsRing t;
                    t.start(ptr, rstep, nrows, tid, group, keep);   // the first loads leave before the table is zeroed
                    hist256_zero(half, tid);
                    group_sync(bar);
                    t.run(lane4, q);
                } else {
                    hist256_zero(half, tid);
                    group_sync(bar);
                    // general path: one warp per tile row, byte spans inside the image, reflected reads outside
                    for (int row = warp; row < p.th; row += kCWarps) {
                        const uint8_t* src_row = y + (size_t)reflect101(y0 + row, p.h) * p.stride;
                        const int xin = min(x0 + p.tw, p.w);  // end of the in-image part
                        if (x0 < xin) hist256_span_warp(src_row + x0, xin - x0, lane, lane4, keep);
                        for (int x = max(x0, p.w) + lane; x < x0 + p.tw; x += 32)
                            hist256_byte(src_row[reflect101(x, p.w)], lane4);
                    }
                }
                if (!tma_ok) {   // a stage never arrived: report and stop (the host resets the workspace)
                    atomicExch(p.status, 3u);
                    break;
                }
                q.prefetch();  // the row sums and the LUT build hide the ticket round trip
                group_sync(bar);
                if (!(p.debug_skip & 8))
                    clahe_tile_lut_block(hist256_row_sum(half, tid), p.clip_limit, p.lut_scale, p.luts + ((size_t)g * p.lut_tiles + r) * 256, s_scratch,
                                         tid, bar);
                publish = g;   // counted after the item's closing barrier, see publish_tile
            }
        } else if (f >= 0) {
            const uint8_t* src = p.in + (unsigned long long)f * p.pitch;
            uint8_t* dst = p.out + (unsigned long long)f * p.pitch;
            if (r < T + I) {
                // ------------------------- cell item: blend four tile LUTs -------------------------
                const int ci = r - T;
                const int cy = small_div(ci, p.nxc), cx = ci - cy * p.nxc;
                const int4 xc = cell_x(p, cx), yc = cell_y(p, cy);
                const int cw = xc.y - xc.x, ch = yc.y - yc.x;  // cell size in pixels (ch <= kMaxCellRows)
                // Geometry of the fast path: full groups of G = 16 (or 8) pixels, one 16- (8-) byte load / store per thread and
                // row.  Columns left over when the cell width is not a multiple of G (and everything when alignment does not
                // allow vector accesses) take the pixel-at-a-time path below.
                const uintptr_t align_or = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)p.stride | (uintptr_t)xc.x;
                const bool fast16 = kUseG16 && ((align_or & 15) == 0) && (cw & 15) == 0 && (cw >> 4) >= 1 && (cw >> 4) <= kCT;
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
                const uint32_t yw_off = yw_base + (uint32_t)tr * 8u, yw_step = (uint32_t)rpp * 8u;

                // dependency wait + table build; everything the row loop needs that does not depend on the LUTs has been
                // started by then (CellRows::start), so pixel loads overlap the wait and the LUT loads
                auto wait_and_build_table = [&]() -> bool {
                    // y weights of the cell's rows
                    for (int i = tid; i < ch; i += kCT) s_yw[i] = __ldg(p.yw + p.y_origin + yc.x + i);
                    if (!q.current_ready()) {   // look-ahead did not see the tiles complete: poll (rare in steady state)
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
                            *s_flag = ok;
                        }
                        group_sync(bar);
                        if (!*s_flag) return false;
                    }
                    tr_.mark(item, 1);
                    if (!(p.debug_skip & 16))
                    // pack the four LUTs: row v = 16 replicas of {bf16 L11 | L21 << 16, bf16 L12 | L22 << 16}
                    {
                        const uint8_t* L = p.luts + (size_t)f * p.lut_tiles * 256;
                        const int v = tid;
                        const uint32_t l11 = __ldcg(L + (size_t)(yc.z * p.tx + xc.z) * 256 + v);
                        const uint32_t l12 = __ldcg(L + (size_t)(yc.z * p.tx + xc.w) * 256 + v);
                        const uint32_t l21 = __ldcg(L + (size_t)(yc.w * p.tx + xc.z) * 256 + v);
                        const uint32_t l22 = __ldcg(L + (size_t)(yc.w * p.tx + xc.w) * 256 + v);
                        uint4 e;
                        if (kByteTable) {
                            e.x = l11 | (l21 << 8) | (l12 << 16) | (l22 << 24);
                            e.y = e.x;
                        } else {
                            e.x = (__float_as_uint((float)l11) >> 16) | (__float_as_uint((float)l21) & 0xffff0000u);
                            e.y = (__float_as_uint((float)l12) >> 16) | (__float_as_uint((float)l22) & 0xffff0000u);
                        }
                        e.z = e.x; e.w = e.y;
                        uint4* row = reinterpret_cast<uint4*>(half + v * kRowBytes);
#pragma unroll
                        for (int j = 0; j < kCellReps / 2; ++j) row[(j + v) & (kCellReps / 2 - 1)] = e;
                    }
                    group_sync(bar);
                    return true;
                };
                bool ok;
                if (fast16 && kUseG16) {
                    CellRows<16, kRows16, kRingBytesPerThread / (16 * kRows16)> cr;
                    cr.start(sp, dp, rstep, nrows, p.xw + xg, tid, group, yw_off, yw_step);
                    ok = wait_and_build_table();
                    if (ok) cr.run(q, lane8);
                } else {
                    CellRows<8, 2, kRingBytesPerThread / 16> cr;
                    cr.start(sp, dp, rstep, nrows, p.xw + xg, tid, group, yw_off, yw_step);
                    ok = wait_and_build_table();
                    if (ok) cr.run(q, lane8);
                }
                if (!ok) break;
                if (xslow < xc.y && !(p.debug_skip & 2)) {
                    // pixel-at-a-time path
                    const int sw = xc.y - xslow;
                    const long long npix = (long long)sw * ch;
                    for (long long i = tid; i < npix; i += kCT) {
                        const int ry = (int)(i / sw), rx = (int)(i - (long long)ry * sw);
                        const int x = xslow + rx, yy = yc.x + ry;
                        const float xa = __ldg(p.xw + x), xa1 = __fsub_rn(kXScale, xa);   // (1 - xa) * kXScale, exactly
                        const uint32_t v = src[(size_t)yy * p.stride + x];
                        const float res = clahe_blend_res(lds_entry_rel((v << kRowShift) + lane8), xa, xa1, lds_b64_rel(yw_base + (uint32_t)ry * 8u));
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
                    const int rows_uv = p.h / 2;
                    const int r0 = min(c * p.uv_rows_chunk, rows_uv), r1 = min(r0 + p.uv_rows_chunk, rows_uv);
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
    publish_tile();
    // the last group out returns the per-frame counters to zero for the next launch
    if (q.finish(s_last))
        for (int i = tid; i < p.n_frames; i += kCT) p.tiles_done[i] = 0;
}

}  // namespace nv12eq
