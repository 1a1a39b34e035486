// nv12eq_api.cu -- context, launch planning and the C-ABI of libnv12eq.so (see include/nv12eq.h).
//
// Host-side structure (replaces the reference's per-worker device context + blocking transfer protocol,
// OpenCLequalHist.cpp:63-81,106-192,349-365):
//   * one nv12eq_ctx per calling thread; it owns a "lane" per slot: a CUDA stream, device in/out buffers, pinned
//     staging buffers (only used when the caller's memory is pageable) and a private kernel workspace, all cached
//     by size.  H2D -> kernel -> D2H of one lane are stream-ordered; different lanes overlap, which gives the
//     double-buffered upload / compute / download pipeline the MPSoC zero-copy path is replaced with.
//   * device-resident entry points use a separate workspace and the caller's stream.
// There is no CPU implementation anywhere in this library.
#include <cuda_runtime.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define NV12EQ_HAVE_SSE2 1
#endif

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/nv12eq.h"
#include "clahe.cuh"
#include "clahe16.cuh"
#include "color.cuh"
#include "equalize.cuh"

using namespace nv12eq;

namespace {

thread_local std::string g_create_error;

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// Kernel workspace: counters that the kernels return to zero themselves (see equalize.cuh / clahe.cuh).
struct Workspace {
    DevBuf hist;      // [frames][256] u32
    DevBuf counters;  // [3][frames] u32: done/tiles_done, applied, (spare)
    DevBuf misc;      // ticket (u32 @0), status (u32 @4)
    DevBuf luts;      // clahe: [frames][tiles][256] u8
    DevBuf cells;     // clahe: int4 xcells[], ycells[]
    DevBuf luma;      // colour: [2][frames][h][w]
    DevBuf hist16;    // clahe16: [planes][tiles][65536] u32, kept zero between launches
    DevBuf luts16;    // clahe16: [planes][tiles][65536] u16
    DevBuf cells16;   // clahe16: [planes][cells][65536] four u16 per value
    DevBuf ormask16;  // clahe16: [planes] OR of the pixel values
    int frames_cap = 0;
    // cached CLAHE geometry
    int gw = 0, gh = 0, gtx = 0, gty = 0, nxc = 0, nyc = 0;
    std::vector<int4> h_cells;   // host copy of the cell tables (x cells, then y cells)
    DevBuf weights;              // CLAHE interpolation weights of the cached geometry: float xw[w], then float2 yw[frame height]
};

// Small pool of host threads for the chroma plane of host-buffer calls.  In passthrough mode the chroma bytes never
// cross PCIe: the GPU only sees the luma planes (8.3 of the 12.4 MB of a 4K frame, each way) while these threads do
// what the reference does on the CPU anyway, memcpy(out + y_size, in + y_size, uv_size) (nextimprovement.cpp:160) or
// memset(out + y_size, 128, uv_size) (OpenCVequalHist.cpp:162), concurrently with the DMA and the kernels.
struct HostTask {
    uint8_t* dst; const uint8_t* src;   // src == nullptr: fill with `value`
    int rows; size_t row_bytes, stride; int value;
    int* pending;                        // counter of the submitting lane, guarded by the pool mutex
};
class HostPool {
public:
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // Large copies use non-temporal stores: the destination is not read back by this thread, and skipping the
    // read-for-ownership of every destination line leaves more host memory bandwidth to the PCIe DMA running beside it.
    static void stream_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#ifdef NV12EQ_HAVE_SSE2
        if (n >= (256u << 10)) {
            const size_t head = (size_t)((16 - ((uintptr_t)dst & 15)) & 15);
            memcpy(dst, src, head);
            dst += head; src += head; n -= head;
            const size_t blocks = n / 64;
            for (size_t i = 0; i < blocks; ++i, dst += 64, src += 64) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
                const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
                const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
            }
            _mm_sfence();
            n -= blocks * 64;
        }
#endif
        memcpy(dst, src, n);
    }
    static void execute(const HostTask& t) {
        for (int r = 0; r < t.rows; ++r) {
            if (t.src) stream_copy(t.dst + (size_t)r * t.stride, t.src + (size_t)r * t.stride, t.row_bytes);
            else memset(t.dst + (size_t)r * t.stride, t.value, t.row_bytes);
        }
    }
    void submit(const HostTask& t) {
        { std::lock_guard<std::mutex> lk(mu_); ++*t.pending; q_.push_back(t); }
        cv_.notify_one();
    }
    void wait(int* pending) {
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return *pending == 0; });
    }
private:
    void run() {
        for (;;) {
            HostTask t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            execute(t);
            bool last;
            { std::lock_guard<std::mutex> lk(mu_); last = (--*t.pending == 0); }
            if (last) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<HostTask> q_;
    bool stop_ = false;
};

struct Lane {
    int host_pending = 0;                // chroma tasks of this lane still running in the host pool
    bool luma_only = false;              // pending job moved only the luma planes over PCIe
    int jw = 0, jh = 0, jstride = 0; size_t jpitch = 0;  // geometry of the pending job (for the pageable-output copy)
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf d_in, d_out;
    HostBuf h_in, h_out;
    Workspace ws;
    // pending job
    bool busy = false;
    uint8_t* user_out = nullptr;     // when non-null, wait() copies h_out -> user_out (pageable output)
    size_t out_bytes = 0;
    int frames = 0;
    uint32_t* h_status = nullptr;    // pinned 4 bytes
};

struct Plan {
    int w, h, stride, n;
    size_t pitch;
    bool flat;
};

}  // namespace

struct nv12eq_ctx {
    int device = 0;
    int max_w = 0, max_h = 0;
    int sm_count = 0;
    int tune_chunks = 0, tune_lag = 0, tune_ctas = 0, tune_schedule = 0;
    std::vector<Lane> lanes;
    cudaStream_t own_stream = nullptr;
    Workspace dev_ws;  // for *_device entry points
    cudaStream_t dev_ws_stream = nullptr;   // stream of the last *_device call: the workspace is ordered on it
    bool dev_ws_used = false;
    cudaEvent_t dev_ws_event = nullptr;
    std::string last_error;
    nv12eq_counters ctr{};
    bool attrs_set = false;
    HostPool* pool = nullptr;   // created on first use by a host-buffer NV12 call
    int pool_threads = -1;      // -1: decide from the host (NV12EQ_HOST_THREADS overrides), 0: chroma inline on the caller
};

namespace {

int fail(nv12eq_ctx* ctx, int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) {
        ctx->last_error = buf;
        ctx->ctr.errors++;
    } else {
        g_create_error = buf;
    }
    return status;
}

#define CK(ctx, call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            int st__ = (e__ == cudaErrorMemoryAllocation) ? NV12EQ_ERR_OUT_OF_MEMORY : NV12EQ_ERR_CUDA;     \
            return fail(ctx, st__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                   \
    } while (0)

int dev_reserve(nv12eq_ctx* ctx, DevBuf& b, size_t bytes, bool zero) {
    if (b.cap >= bytes && b.p) return NV12EQ_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t cap = std::max<size_t>(bytes, 256);
    CK(ctx, cudaMalloc(&b.p, cap));
    b.cap = cap;
    if (zero) {
        // cudaMemset runs on the legacy default stream, which non-blocking streams do not wait for
        CK(ctx, cudaMemset(b.p, 0, cap));
        CK(ctx, cudaStreamSynchronize(0));
    }
    return NV12EQ_OK;
}
int host_reserve(nv12eq_ctx* ctx, HostBuf& b, size_t bytes) {
    if (b.cap >= bytes && b.p) return NV12EQ_OK;
    if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
    CK(ctx, cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault));
    b.cap = bytes;
    return NV12EQ_OK;
}
void dev_release(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }
void host_release(HostBuf& b) { if (b.p) cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }

void ws_release(Workspace& w) {
    dev_release(w.hist); dev_release(w.counters); dev_release(w.misc); dev_release(w.luts); dev_release(w.cells); dev_release(w.weights);
    dev_release(w.luma); dev_release(w.hist16); dev_release(w.luts16); dev_release(w.cells16); dev_release(w.ormask16);
    w.frames_cap = 0; w.gw = w.gh = w.gtx = w.gty = 0;
}

// Counters are (re)allocated zeroed; the kernels keep them zero between launches.
int ws_reserve_frames(nv12eq_ctx* ctx, Workspace& w, int frames) {
    if (!w.misc.p) {
        int rc = dev_reserve(ctx, w.misc, 256, true);
        if (rc) return rc;
    }
    if (frames <= w.frames_cap) return NV12EQ_OK;
    int cap = std::max(frames, 16);
    dev_release(w.hist); dev_release(w.counters);
    int rc = dev_reserve(ctx, w.hist, (size_t)cap * 256 * sizeof(uint32_t), true);
    if (rc) return rc;
    rc = dev_reserve(ctx, w.counters, (size_t)cap * 3 * sizeof(uint32_t), true);
    if (rc) return rc;
    w.frames_cap = cap;
    return NV12EQ_OK;
}
// The kernels' sticky status word: non-zero when a dependency wait timed out (a logic error; cannot happen by construction).
// The CTAs that gave up leave histograms and counters non-zero, so the workspace is put back to its zero state before
// the next launch can trip over it.  Call with the stream idle.
void ws_reset(Workspace& w, cudaStream_t st) {
    if (w.misc.p) cudaMemsetAsync(w.misc.p, 0, w.misc.cap, st);
    if (w.hist.p) cudaMemsetAsync(w.hist.p, 0, w.hist.cap, st);
    if (w.counters.p) cudaMemsetAsync(w.counters.p, 0, w.counters.cap, st);
    cudaStreamSynchronize(st);
}
uint32_t* ws_counter(Workspace& w, int which) { return reinterpret_cast<uint32_t*>(w.counters.p) + (size_t)which * w.frames_cap; }
uint32_t* ws_ticket(Workspace& w) { return reinterpret_cast<uint32_t*>(w.misc.p); }
uint32_t* ws_status(Workspace& w) { return reinterpret_cast<uint32_t*>(w.misc.p) + 1; }

int ensure_attrs(nv12eq_ctx* ctx) {
    if (ctx->attrs_set) return NV12EQ_OK;
    CK(ctx, cudaFuncSetAttribute(equalize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneTableBytes));
    CK(ctx, cudaFuncSetAttribute(equalize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneTableBytes));
    CK(ctx, cudaFuncSetAttribute(equalize_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneTableBytes));
    CK(ctx, cudaFuncSetAttribute(equalize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneTableBytes));
    CK(ctx, cudaFuncSetAttribute(color_equalize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kColorEqSmemBytes));
    CK(ctx, cudaFuncSetAttribute(clahe16_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC16HistSmemBytes));
    CK(ctx, cudaFuncSetAttribute(clahe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClaheSmemBytes));
    CK(ctx, cudaFuncSetAttribute(clahe_kernel<kClaheCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClaheSmemBytes));
    CK(ctx, cudaFuncSetAttribute(clahe_kernel<kClaheCtas - 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClaheSmemBytes));
    ctx->attrs_set = true;
    return NV12EQ_OK;
}

int check_geometry(nv12eq_ctx* ctx, int w, int h, int stride, int n, size_t pitch, int uv_mode) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (w <= 0 || h <= 0 || stride < w || n < 0) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad geometry w=%d h=%d stride=%d n=%d", w, h, stride, n);
    if (uv_mode < 0 || uv_mode > 2) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad uv_mode %d", uv_mode);
    if ((long long)w * h >= (1ll << 31)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame has 2^31 pixels or more");
    if (w > ctx->max_w || h > ctx->max_h) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame %dx%d exceeds context maximum %dx%d", w, h, ctx->max_w, ctx->max_h);
    size_t need = (size_t)stride * (size_t)(h + h / 2);
    if (n > 1 && pitch < need) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "frame_pitch %zu < frame size %zu", pitch, need);
    return NV12EQ_OK;
}


// ---------------------------------------------------------------------------------------------------------
// equalizeHist launch planning
// ---------------------------------------------------------------------------------------------------------
int launch_equalize(nv12eq_ctx* ctx, Workspace& ws, const uint8_t* d_in, uint8_t* d_out, int n, size_t pitch, int w,
                    int h, int stride, int uv_mode, cudaStream_t st, int phases_override = 0, uint32_t* ext_hist = nullptr,
                    long long total_px = 0) {
    if (n == 0) return NV12EQ_OK;
    int rc = ensure_attrs(ctx);
    if (rc) return rc;
    rc = ws_reserve_frames(ctx, ws, n);
    if (rc) return rc;

    EqParams p{};
    p.in = d_in; p.out = d_out; p.pitch = pitch; p.n_frames = n;
    p.w = w; p.h = h; p.stride = stride; p.flat = (stride == w);
    p.uv_mode = uv_mode;
    p.y_bytes = (unsigned long long)w * h;
    p.uv_bytes = (unsigned long long)w * (h / 2);
    p.total_px = total_px ? total_px : (long long)w * h;

    const int per_sm = ctx->tune_ctas > 0 ? std::min(ctx->tune_ctas, 4) : kEqCtas;
    // chunks per frame: ~128 KB of luma per item, but never fewer items than ~2 waves of CTAs
    const int ctas = ctx->sm_count * per_sm;
    long long C = (long long)((p.y_bytes + 131071) / 131072);
    if (ctx->tune_chunks > 0) C = ctx->tune_chunks;
    else if ((long long)n * C < 2ll * ctas) C = std::min<long long>((2ll * ctas + n - 1) / n, (long long)((p.y_bytes + 16383) / 16384));
    C = std::max<long long>(1, std::min<long long>(C, 1 << 16));
    if (!p.flat) C = std::min<long long>(C, h);
    p.chunks = (int)C;
    auto round4k = [](unsigned long long v) { return (v + 4095ull) & ~4095ull; };
    p.y_chunk = std::max<unsigned long long>(4096, round4k((p.y_bytes + C - 1) / C));
    p.uv_chunk = std::max<unsigned long long>(4096, round4k((p.uv_bytes + C - 1) / C));
    p.y_rows_chunk = (int)((h + C - 1) / C);
    p.uv_rows_chunk = (int)((h / 2 + C - 1) / C);
    // Frame lag between a frame's histogram items and its apply items.  All resident CTAs finish about one item per
    // item time, so the histogram items of a frame are done once ~grid more tickets have been drawn: lag ~ grid /
    // items_per_slot.  Measured on B200 with tickets drawn late (tools/sweep.py): best lag 2 at 4K (296 CTAs / 128 items),
    // 8 at 1080p (296 / 32) -- i.e. floor(grid / items_per_slot).  The Y planes of `lag` frames have to stay in L2 for
    // the second read: cap the footprint at ~48 MB.
    {
        const long long grid_ctas = (long long)ctx->sm_count * per_sm;
        long long lag = std::max<long long>(1, grid_ctas / (2 * C));
        const long long cap = std::max<long long>(1, (48ll << 20) / (long long)std::max<unsigned long long>(1, p.y_bytes));
        lag = std::min(lag, cap);
        if (ctx->tune_lag > 0) lag = ctx->tune_lag;
        if (ctx->tune_lag < 0) lag = 0;
        p.lag = (int)std::min<long long>(lag, std::max(n - 1, 0));
    }
    p.hist = ext_hist ? ext_hist : reinterpret_cast<uint32_t*>(ws.hist.p);
    p.applied = ws_counter(ws, 1);
    p.ticket = ws_ticket(ws);
    p.status = ws_status(ws);

    const size_t smem = kLaneTableBytes;
    auto go = [&](int phases) -> int {
        p.phases = phases;
        const bool both = (phases & PH_HIST) && (phases & PH_APPLY);
        long long items = (long long)(n + (both ? p.lag : 0)) * 2 * C;
        if (items >= (1ll << 32)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "too many work items (%lld): split the batch", items);
        int grid = (int)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * per_sm, items));
        if (per_sm <= 1) equalize_kernel<1><<<grid, kThreads, smem, st>>>(p);
        else if (per_sm == 2) equalize_kernel<2><<<grid, kThreads, smem, st>>>(p);
        else if (per_sm == 3) equalize_kernel<3><<<grid, kThreads, smem, st>>>(p);
        else equalize_kernel<4><<<grid, kThreads, smem, st>>>(p);
        ctx->ctr.kernel_launches++;
        CK(ctx, cudaGetLastError());
        return NV12EQ_OK;
    };
    if (phases_override) return go(phases_override);
    if (ctx->tune_schedule == 2) {
        rc = go(PH_HIST);
        if (rc) return rc;
        return go(PH_APPLY);
    }
    return go(PH_HIST | PH_APPLY);
}

// ---------------------------------------------------------------------------------------------------------
// CLAHE launch planning
// ---------------------------------------------------------------------------------------------------------
struct ClaheGeom {
    int extW, extH, tw, th, clip_limit;
    float lut_scale, inv_tw, inv_th;
};

ClaheGeom clahe_geometry(int w, int h, double clip, int tx, int ty) {
    ClaheGeom g{};
    g.extW = w; g.extH = h;
    if (w % tx != 0 || h % ty != 0) {  // OpenCV pads BOTH dimensions, even the one that divides
        g.extW = w + (tx - (w % tx));
        g.extH = h + (ty - (h % ty));
    }
    g.tw = g.extW / tx; g.th = g.extH / ty;
    const int area = g.tw * g.th;
    g.clip_limit = 0;
    if (clip > 0.0) g.clip_limit = std::max(1, (int)(clip * area / 256.0));
    g.lut_scale = 255.0f / (float)area;
    g.inv_tw = 1.0f / (float)g.tw;
    g.inv_th = 1.0f / (float)g.th;
    return g;
}

// Interpolation cells along one axis: maximal runs of positions with the same floor(pos*inv - 0.5), cut into pieces
// of at most max_len so that one cell is one CTA-sized item.
void axis_cells(int n, float inv, int ntiles, int max_len, int align, std::vector<int4>& out) {
    out.clear();
    int start = 0;
    auto t1_of = [&](int pos) {
        volatile float f = (float)pos * inv;  // separately rounded multiply and subtract, as on the device
        volatile float g = f - 0.5f;
        return (int)floorf(g);
    };
    int cur = t1_of(0);
    auto flush = [&](int s, int e, int t1) {
        const int a = std::max(t1, 0), b = std::min(t1 + 1, ntiles - 1);
        int pos = s;
        while (pos < e) {
            int len = std::min(max_len, e - pos);
            if (pos + len < e && align > 1) {  // keep interior cuts aligned
                int cut = ((pos + len) / align) * align;
                if (cut > pos) len = cut - pos;
            }
            out.push_back(make_int4(pos, pos + len, a, b));
            pos += len;
        }
    };
    for (int pos = 1; pos < n; ++pos) {
        const int t = t1_of(pos);
        if (t != cur) { flush(start, pos, cur); start = pos; cur = t; }
    }
    flush(start, n, cur);
}

#ifndef NV12EQ_CLAHE_UV_CHUNK
#define NV12EQ_CLAHE_UV_CHUNK (128 << 10)
#endif
constexpr unsigned long long kUvChunk = NV12EQ_CLAHE_UV_CHUNK;   // chroma bytes per uv item (512 KB items measured 7 % slower at 1080p, equal at 4K)
// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library does not link libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return reinterpret_cast<TensorMapEncodeFn>(f);
    }();
    return fn;
}
// Rank-3 byte tensor (x, row, frame) over a batch of planes, box bw x bh x 1, no swizzle.
bool encode_plane_map(CUtensorMap* m, const void* base, int w, int h, int n, size_t stride, size_t pitch, int bw, int bh) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)pitch};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Row-band mode of launch_clahe (spatial split of one frame over GPUs): the plane is tile rows [first_tile_row, first_tile_row + ty)
// of a frame of full_h rows and full_ty tile rows (grid must divide the frame).  LutsOnly writes the band's ty * tx tables to
// `luts`; ApplyOnly interpolates the band from `luts` = (ty + 2) * tx tables: tile rows first_tile_row - 1 .. first_tile_row + ty
// (the rows outside the frame are never addressed: the y cells carry clamped tile rows).
struct ClaheBand {
    enum Mode { LutsOnly = 1, ApplyOnly = 2 } mode;
    int full_h, full_ty, first_tile_row;
    uint8_t* luts;
};
int launch_clahe(nv12eq_ctx* ctx, Workspace& ws, const uint8_t* d_in, uint8_t* d_out, int n, size_t pitch, int w, int h,
                 int stride, double clip, int tx, int ty, int uv_mode, cudaStream_t st, const ClaheBand* band = nullptr) {
    if (n == 0) return NV12EQ_OK;
    if (tx < 1 || ty < 1 || (long long)tx * ty > (1 << 20)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tx, ty);
    int rc = ensure_attrs(ctx);
    if (rc) return rc;
    rc = ws_reserve_frames(ctx, ws, n);
    if (rc) return rc;
    const ClaheGeom g = clahe_geometry(w, h, clip, tx, ty);
    const int T = tx * ty;
    if (!band) {
        rc = dev_reserve(ctx, ws.luts, (size_t)n * T * 256, false);
        if (rc) return rc;
    }
    if (band || ws.gw != w || ws.gh != h || ws.gtx != tx || ws.gty != ty || !ws.cells.p) {
        std::vector<int4> xc, yc;
        axis_cells(w, g.inv_tw, tx, 1024, 8, xc);
        if (band && band->mode == ClaheBand::ApplyOnly) {
            // the frame's y cells cut to the band's rows; tile rows re-based to the halo grid (row 0 = first_tile_row - 1)
            std::vector<int4> all;
            axis_cells(band->full_h, g.inv_th, band->full_ty, kMaxCellRows, 1, all);
            const int Y0 = band->first_tile_row * g.th, Y1 = Y0 + h, base = band->first_tile_row - 1;
            for (const int4& c : all) {
                const int a = std::max(c.x, Y0), b = std::min(c.y, Y1);
                if (a < b) yc.push_back(make_int4(a - Y0, b - Y0, c.z - base, c.w - base));
            }
        } else {
            axis_cells(h, g.inv_th, ty, kMaxCellRows, 1, yc);
        }
        rc = dev_reserve(ctx, ws.cells, (xc.size() + yc.size()) * sizeof(int4), false);
        if (rc) return rc;
        // stream-ordered after any kernel still reading the previous tables; the pageable source makes the call
        // return only after the bytes have been staged
        CK(ctx, cudaMemcpyAsync(ws.cells.p, xc.data(), xc.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(reinterpret_cast<int4*>(ws.cells.p) + xc.size(), yc.data(), yc.size() * sizeof(int4),
                                cudaMemcpyHostToDevice, st));
        // interpolation weights (same operation order as OpenCV / the oracle: every operation rounded separately, see the Makefile's
        // -ffp-contract=off): x weights of the plane, y weights of the whole frame (a band indexes them from its first row)
        {
            const int fh = (band && band->mode == ClaheBand::ApplyOnly) ? band->full_h : h;
            std::vector<float> wt((size_t)w + 2 * (size_t)fh);
            auto frac = [](int pos, float inv, float& a, float& a1) {
                volatile float f = (float)pos * inv;
                volatile float g2 = f - 0.5f;
                const float t1 = floorf(g2);
                volatile float av = g2 - t1;
                a = av;
                volatile float bv = 1.0f - av;
                a1 = bv;
            };
            for (int x = 0; x < w; ++x) { float a, a1; frac(x, g.inv_tw, a, a1); wt[x] = a * kXScale; }
            for (int y = 0; y < fh; ++y) { float a, a1; frac(y, g.inv_th, a, a1); wt[(size_t)w + 2 * y] = a1 * kYScale; wt[(size_t)w + 2 * y + 1] = a * kYScale; }
            const size_t xbytes = (((size_t)w * sizeof(float)) + 255) & ~(size_t)255;   // keeps the float2 table aligned
            rc = dev_reserve(ctx, ws.weights, xbytes + 2 * (size_t)fh * sizeof(float), false);
            if (rc) return rc;
            CK(ctx, cudaMemcpyAsync(ws.weights.p, wt.data(), (size_t)w * sizeof(float), cudaMemcpyHostToDevice, st));
            CK(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ws.weights.p) + xbytes, wt.data() + w, 2 * (size_t)fh * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        ws.gw = band ? -1 : w; ws.gh = h; ws.gtx = tx; ws.gty = ty;   // band tables are never reused
        ws.nxc = (int)xc.size(); ws.nyc = (int)yc.size();
        ws.h_cells = xc;
        ws.h_cells.insert(ws.h_cells.end(), yc.begin(), yc.end());
    }

    ClaheParams p{};
    p.in = d_in; p.out = d_out; p.pitch = pitch; p.n_frames = n;
    p.w = w; p.h = h; p.stride = stride; p.flat = (stride == w); p.uv_mode = uv_mode;
    p.tx = tx; p.ty = ty; p.tw = g.tw; p.th = g.th;
    p.padded = (g.extW != w || g.extH != h);
    p.clip_limit = g.clip_limit; p.lut_scale = g.lut_scale; p.inv_tw = g.inv_tw; p.inv_th = g.inv_th;
    p.nxc = ws.nxc; p.nyc = ws.nyc;
    p.xcells = reinterpret_cast<const int4*>(ws.cells.p);
    p.ycells = p.xcells + ws.nxc;
    p.xw = reinterpret_cast<const float*>(ws.weights.p);
    p.yw = reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(ws.weights.p) + ((((size_t)w * sizeof(float)) + 255) & ~(size_t)255));
    p.cells_in_params = (ws.nxc <= kParamCells && ws.nyc <= kParamCells);
    if (p.cells_in_params) {
        std::copy(ws.h_cells.begin(), ws.h_cells.begin() + ws.nxc, p.xc_small);
        std::copy(ws.h_cells.begin() + ws.nxc, ws.h_cells.end(), p.yc_small);
    }
    const bool uv_work = (uv_mode == UV_GRAY128) || (uv_mode == UV_COPY && d_in != d_out);
    p.uv_bytes = (unsigned long long)w * (h / 2);
    int U = 0;
    if (uv_work && h / 2 > 0) {
        U = (int)std::max<unsigned long long>(1, (p.uv_bytes + kUvChunk - 1) / kUvChunk);
        if (!p.flat) U = std::min(U, h / 2);
        p.uv_chunk = std::max<unsigned long long>(4096, (((p.uv_bytes + U - 1) / U) + 4095ull) & ~4095ull);
        p.uv_rows_chunk = (h / 2 + U - 1) / U;
    }
    p.uv_chunks = U;
    p.tile_items = T; p.lut_tiles = T;
    p.luts = reinterpret_cast<uint8_t*>(ws.luts.p);
    if (band) {
        p.luts = band->luts;
        if (band->mode == ClaheBand::LutsOnly) p.cells_off = 1;
        else { p.tile_items = 0; p.lut_tiles = tx * (ty + 2); p.y_origin = band->first_tile_row * g.th; }
    }
    p.tiles_done = ws_counter(ws, 0);
    p.ticket = ws_ticket(ws);
    p.status = ws_status(ws);
    if (const char* dbg = getenv("NV12EQ_DEBUG_SKIP")) p.debug_skip = atoi(dbg);

    // Tile rows through the TMA unit when the layout allows it (16-byte aligned planes and tile columns, tiles inside the image);
    // NV12EQ_CLAHE_TMA=0 keeps the per-thread cp.async ring (A/B tool).
    {
        const char* e = getenv("NV12EQ_CLAHE_TMA");
        const bool want = !(e && e[0] == '0');
        const int tw = g.tw, th = g.th;
        const int nb = (tw + 255) / 256;
        // L2 prefetch boxes: a tile is nb x nrb boxes (any 16-byte aligned layout; boxes may overshoot a border cell into its
        // neighbour, which another item is about to read anyway)
        const char* epf = getenv("NV12EQ_CLAHE_PF");
        if (kTmaPrefetch && !(epf && epf[0] == '0') && !p.padded && ((uintptr_t)d_in & 15) == 0 && stride % 16 == 0 && pitch % 16 == 0) {
            const int nbx = (tw + 255) / 256, nby = (th + 255) / 256;
            const int bwp = std::min(w, (((tw + nbx - 1) / nbx) + 15) & ~15), bhp = (th + nby - 1) / nby;
            if (bwp % 16 == 0 && bwp <= 256 && encode_plane_map(&p.pf_map, d_in, w, h, n, (size_t)stride, pitch, bwp, bhp)) {
                p.pf_on = 1; p.pf_bw = bwp; p.pf_bh = bhp;
            }
        }
        if (kTmaTiles && want && !p.padded && ((uintptr_t)d_in & 15) == 0 && stride % 16 == 0 && pitch % 16 == 0 && tw % 16 == 0 && tw <= kStageBytes &&
            tw % nb == 0 && (tw / nb) % 16 == 0) {
            const int bw = tw / nb;
            // rows per stage: the boxes of a stage (each on a 128-byte boundary) fit the stage, at most one 16-byte piece per thread
            int bh = std::max(1, std::min(std::min(256, kStageBytes / tw), th));
            while (bh > 1 && (((bw * bh + 127) & ~127) * (nb - 1) + bw * bh > kStageBytes || nb * ((bw * bh) >> 4) > kCT)) --bh;
            const int nst = (th + bh - 1) / bh;
            const int tail = th - (nst - 1) * bh;
            bool ok = encode_plane_map(&p.tile_map[0], d_in, w, h, n, (size_t)stride, pitch, bw, bh);
            if (ok && tail != bh) ok = encode_plane_map(&p.tile_map[1], d_in, w, h, n, (size_t)stride, pitch, bw, tail);
            if (ok) { p.tma_tiles = 1; p.tma_bw = bw; p.tma_nb = nb; p.tma_bh = bh; p.tma_bh_tail = tail; p.tma_nst = nst; }
        }
    }
    const long long per_slot = (long long)p.tile_items + (p.cells_off ? 0 : (long long)p.nxc * p.nyc + U);
    // CTAs per SM: four 64-register CTAs, or three with 80 registers (no rematerialisation in the blend loop).  Measured on
    // 4K frames with cool-downs between runs (tools/sweep.py --ctas 4,3,4,3 --cooldown 4): 8x8 grid (130 K-pixel tiles) 7.22 vs
    // 7.35 us per frame, 6x6 (230 K) 7.05 vs 7.04, 4x4 (518 K) 7.56 vs 7.48 -- the extra CTA wins until the tiles are so
    // large that the per-item phases no longer matter.
    const int auto_ctas = kClaheCtas;
    const int per_sm = ctx->tune_ctas > 0 ? std::min(ctx->tune_ctas, kClaheCtas) : auto_ctas;
    {
        // same reasoning as for equalizeHist; tile items run ~1.5x longer than the average item
        const long long grid_ctas = (long long)ctx->sm_count * per_sm * kGroups;   // work groups drawing tickets
        long long lag = (3 * grid_ctas / 2 + per_slot - 1) / per_slot + 1;
        const long long cap = std::max<long long>(1, (48ll << 20) / std::max<long long>(1, (long long)w * h));
        lag = std::min(lag, cap);
        if (ctx->tune_lag > 0) lag = ctx->tune_lag;
        if (ctx->tune_lag < 0) lag = 0;
        p.lag = (int)std::min<long long>(lag, std::max(n - 1, 0));
    }
    const long long items = (long long)(n + p.lag) * per_slot;
    p.slot_magic = (per_slot > 1 && items * per_slot < (1ll << 32)) ? (uint32_t)((1ull << 32) / (unsigned long long)per_slot) + 1u : 0u;
    if (items >= (1ll << 32)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "too many work items");
    // developer tool: NV12EQ_TRACE=<file> dumps per-item timestamps of this launch (synchronous, slow)
    DevBuf trace_buf;
    const char* trace_path = NV12EQ_ITEM_TRACE ? getenv("NV12EQ_TRACE") : nullptr;   // needs a build with -DNV12EQ_ITEM_TRACE=1
    if (trace_path) {
        rc = dev_reserve(ctx, trace_buf, (size_t)items * 4 * sizeof(unsigned long long), true);
        if (rc) return rc;
        p.trace = reinterpret_cast<unsigned long long*>(trace_buf.p);
    }
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * per_sm, (items + kGroups - 1) / kGroups));
    if (per_sm <= 1) clahe_kernel<1><<<grid, kBlockThreads, kClaheSmemBytes, st>>>(p);
    else if (per_sm < kClaheCtas) clahe_kernel<kClaheCtas - 1><<<grid, kBlockThreads, kClaheSmemBytes, st>>>(p);
    else clahe_kernel<kClaheCtas><<<grid, kBlockThreads, kClaheSmemBytes, st>>>(p);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    if (trace_path) {
        std::vector<unsigned long long> h((size_t)items * 4);
        CK(ctx, cudaStreamSynchronize(st));
        CK(ctx, cudaMemcpy(h.data(), trace_buf.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE* fp = fopen(trace_path, "wb")) {
            const long long hdr[8] = {items, per_slot, T, (long long)p.nxc * p.nyc, U, p.lag, grid, n};
            fwrite(hdr, sizeof hdr, 1, fp);
            fwrite(h.data(), sizeof(unsigned long long), h.size(), fp);
            fclose(fp);
        }
        dev_release(trace_buf);
    }
    return NV12EQ_OK;
}

// ---------------------------------------------------------------------------------------------------------
// colour path
// ---------------------------------------------------------------------------------------------------------
// Fused two-pass colour equalization (color.cuh::color_equalize_kernel) for flat, 16-byte aligned frames.
int launch_color_fused(nv12eq_ctx* ctx, Workspace& ws, const uint8_t* d_in, uint8_t* d_out, int n, size_t pitch, int w, int h, int mode,
                       cudaStream_t st) {
    int rc = ensure_attrs(ctx);
    if (rc) return rc;
    rc = ws_reserve_frames(ctx, ws, n);
    if (rc) return rc;
    ColorEqParams p{};
    p.in = d_in; p.out = d_out; p.pitch = pitch; p.n_frames = n;
    p.npx = (unsigned long long)w * h;
    p.rounds = (p.npx + 511) / 512;
    const int ctas = ctx->sm_count * 2;
    long long C = (long long)((p.rounds + 255) / 256);  // ~128K pixels (384 KB of BGR) per item
    if ((long long)n * C < 2ll * ctas) C = std::min<long long>((2ll * ctas + n - 1) / n, (long long)((p.rounds + 15) / 16));
    if (ctx->tune_chunks > 0) C = ctx->tune_chunks;
    C = std::max<long long>(1, std::min<long long>(C, std::min<long long>(1 << 16, (long long)p.rounds)));
    p.chunks = (int)C;
    p.rounds_chunk = (p.rounds + C - 1) / C;
    long long lag = (10ll * ctas + 16 * 2 * C - 1) / (10 * 2 * C);
    lag = std::min<long long>(lag, 8);
    if (ctx->tune_lag > 0) lag = ctx->tune_lag;
    if (ctx->tune_lag < 0) lag = 0;   // apply items of a frame directly behind its histogram items (the frame is still in L2; CTAs wait at the switch)
    p.lag = (int)std::min<long long>(lag, std::max(n - 1, 0));
    if (mode == COLOR_YUV) { p.kB = 8061; p.kR = 14369; p.iB = 33292; p.iG1 = -6472; p.iG2 = -9519; p.iR = 18678; }
    else { p.kB = 9241; p.kR = 11682; p.iB = 29049; p.iG1 = -5636; p.iG2 = -11698; p.iR = 22987; }
    p.hist = reinterpret_cast<uint32_t*>(ws.hist.p);
    p.applied = ws_counter(ws, 1);
    p.ticket = ws_ticket(ws);
    p.status = ws_status(ws);
    const long long items = (long long)(n + p.lag) * 2 * C;
    if (items >= (1ll << 32)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "too many work items (%lld): split the batch", items);
    const int grid = (int)std::max<long long>(1, std::min<long long>(ctas, items));
    color_equalize_kernel<2><<<grid, kThreads, kColorEqSmemBytes, st>>>(p);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}

int launch_color(nv12eq_ctx* ctx, Workspace& ws, const uint8_t* d_in, uint8_t* d_out, int n, size_t pitch, int w, int h,
                 int stride, int mode, bool use_clahe, double clip, int tx, int ty, cudaStream_t st) {
    if (n == 0) return NV12EQ_OK;
    const bool fused_ok = !use_clahe && stride == 3 * w && ((long long)w * h) % 16 == 0 && (pitch % 16 == 0 || n == 1) &&
                          ((((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0) && !getenv("NV12EQ_COLOR_THREE_PASS");
    if (fused_ok) return launch_color_fused(ctx, ws, d_in, d_out, n, pitch, w, h, mode, st);
    const size_t plane = (size_t)w * h;
    int rc = dev_reserve(ctx, ws.luma, 2 * plane * n + 64, false);
    if (rc) return rc;
    uint8_t* y1 = reinterpret_cast<uint8_t*>(ws.luma.p);
    uint8_t* y2 = y1 + ((plane * n + 15) & ~(size_t)15);
    ColorParams cp{};
    cp.bgr_in = d_in; cp.bgr_out = d_out; cp.bgr_pitch = pitch; cp.n_frames = n;
    cp.w = w; cp.h = h; cp.stride = stride; cp.y_plane = y1; cp.y2_plane = y2; cp.mode = mode;
    const long long quads = ((long long)plane + 3) / 4;
    const int gx = (int)std::max<long long>(1, std::min<long long>((quads + kColorThreads - 1) / kColorThreads, (long long)ctx->sm_count * 8));
    dim3 grid(gx, n);
    bgr_to_luma_kernel<<<grid, kColorThreads, 0, st>>>(cp);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    // the luma planes are Y-only "frames" of pitch w*h: stride == w, chroma skipped
    if (use_clahe) rc = launch_clahe(ctx, ws, y1, y2, n, plane, w, h, w, clip, tx, ty, UV_SKIP, st);
    else rc = launch_equalize(ctx, ws, y1, y2, n, plane, w, h, w, UV_SKIP, st);
    if (rc) return rc;
    bgr_recombine_kernel<<<grid, kColorThreads, 0, st>>>(cp);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host lanes
// ---------------------------------------------------------------------------------------------------------
bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

enum class Op { Equalize, Clahe, ColorEq, ColorClahe };
struct Job {
    Op op;
    const uint8_t* in; uint8_t* out;
    int n; size_t pitch; size_t frame_bytes;
    int w, h, stride, uv_mode;
    double clip; int tx, ty; int color_mode;
};

// Luma planes of the n frames of a job between two buffers of the job's layout (either side may be the device).
int copy_luma_async(nv12eq_ctx* ctx, uint8_t* dst, const uint8_t* src, const Job& j, cudaMemcpyKind kind, cudaStream_t st) {
    if (j.stride == j.w) {  // flat: one luma plane is one contiguous span, frames are `pitch` apart
        const size_t plane = (size_t)j.w * j.h;
        if (j.n == 1) { CK(ctx, cudaMemcpyAsync(dst, src, plane, kind, st)); }
        else { CK(ctx, cudaMemcpy2DAsync(dst, j.pitch, src, j.pitch, plane, (size_t)j.n, kind, st)); }
    } else {                // strided: only the `w` payload bytes of every row move; padding is never touched
        for (int k = 0; k < j.n; ++k)
            CK(ctx, cudaMemcpy2DAsync(dst + (size_t)k * j.pitch, j.stride, src + (size_t)k * j.pitch, j.stride, (size_t)j.w, (size_t)j.h, kind, st));
    }
    return NV12EQ_OK;
}
HostPool* host_pool(nv12eq_ctx* ctx);
// Host copies smaller than this stay on the calling thread: waking the pool costs more than it saves (NV12EQ_POOL_MIN_MB overrides)
size_t pool_min_bytes() {
    static const size_t v = [] {
        const char* e = getenv("NV12EQ_POOL_MIN_MB");
        return (size_t)(e ? std::max(0, atoi(e)) : 4) << 20;
    }();
    return v;
}
// Luma planes between the caller's pageable frames and pinned staging.  Large jobs are cut into row blocks for the context's
// host pool (the calling thread takes the first block): one thread copies a 4K plane in ~0.8 ms, several PCIe transfers' worth.
void copy_luma_host(nv12eq_ctx* ctx, uint8_t* dst, const uint8_t* src, int n, size_t pitch, int w, int h, int stride) {
    HostPool* pool = ((size_t)n * w * h >= pool_min_bytes()) ? host_pool(ctx) : nullptr;
    const int parts = pool ? std::max(1, std::min((ctx->pool_threads + 1 + n - 1) / n, std::min(8, h))) : 1;   // row blocks per frame
    const bool flat = stride == w;
    int pending = 0;
    for (int k = n - 1; k >= 0; --k)
        for (int b = parts - 1; b >= 0; --b) {
            const int r0 = (int)((long long)h * b / parts), r1 = (int)((long long)h * (b + 1) / parts);
            const size_t off = (size_t)k * pitch + (size_t)r0 * stride;
            HostTask t{dst + off, src + off, flat ? 1 : r1 - r0, flat ? (size_t)(r1 - r0) * stride : (size_t)w, (size_t)stride, 0, &pending};
            if (pool && (k > 0 || b > 0)) pool->submit(t);
            else HostPool::execute(t);
        }
    if (pool) pool->wait(&pending);
}
size_t luma_payload(const Job& j) { return (size_t)j.n * (size_t)j.w * (size_t)j.h; }

HostPool* host_pool(nv12eq_ctx* ctx) {
    if (ctx->pool_threads < 0) {
        int n = (int)std::thread::hardware_concurrency() / 4;
        if (const char* e = getenv("NV12EQ_HOST_THREADS")) n = atoi(e);
        ctx->pool_threads = std::max(0, std::min(n, 8));
    }
    if (ctx->pool_threads > 0 && !ctx->pool) ctx->pool = new (std::nothrow) HostPool(ctx->pool_threads);
    return ctx->pool;
}

// Chroma of a host NV12 job, on the host (see HostPool).  Runs concurrently with the GPU work queued just before.
void host_chroma(nv12eq_ctx* ctx, Lane& L, const Job& j) {
    const int rows = j.h / 2;
    const bool copy = (j.uv_mode == UV_COPY && j.in != j.out);
    if (rows == 0 || !(copy || j.uv_mode == UV_GRAY128)) return;
    const size_t off = (size_t)j.stride * j.h;
    const bool flat = (j.stride == j.w);
    HostPool* pool = ((size_t)j.n * rows * j.w >= (1u << 20)) ? host_pool(ctx) : nullptr;  // small jobs: inline
    for (int k = 0; k < j.n; ++k) {
        HostTask t{};
        t.dst = j.out + (size_t)k * j.pitch + off;
        t.src = copy ? j.in + (size_t)k * j.pitch + off : nullptr;
        t.value = 128;
        t.rows = flat ? 1 : rows;
        t.row_bytes = flat ? (size_t)j.w * rows : (size_t)j.w;
        t.stride = (size_t)j.stride;
        t.pending = &L.host_pending;
        if (pool) pool->submit(t); else HostPool::execute(t);
    }
}

// Host NV12 frames: only the luma planes cross PCIe, the kernels run with UV_SKIP, the chroma is handled on the host.
int lane_submit_nv12(nv12eq_ctx* ctx, Lane& L, const Job& j) {
    const size_t span = (size_t)(j.n - 1) * j.pitch + j.frame_bytes;  // bytes covered in the caller's buffers
    int rc;
    if ((rc = dev_reserve(ctx, L.d_in, span, false))) return rc;
    const bool in_place = (j.in == j.out);
    if (!in_place && (rc = dev_reserve(ctx, L.d_out, span, false))) return rc;
    uint8_t* d_in = reinterpret_cast<uint8_t*>(L.d_in.p);
    uint8_t* d_out = in_place ? d_in : reinterpret_cast<uint8_t*>(L.d_out.p);

    // everything that can fail without touching the device happens before the first copy is queued: a caller that sees an
    // error may free or reuse its buffers at once
    if (j.op == Op::Clahe && (j.tx < 1 || j.ty < 1 || (long long)j.tx * j.ty > (1 << 20)))
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", j.tx, j.ty);
    const bool in_pinned = is_pinned(j.in), out_pinned = is_pinned(j.out);
    if (!in_pinned && (rc = host_reserve(ctx, L.h_in, span))) return rc;
    if (!out_pinned && (rc = host_reserve(ctx, L.h_out, span))) return rc;
    if ((rc = ws_reserve_frames(ctx, L.ws, j.n))) return rc;
    const uint8_t* src = j.in;
    if (!in_pinned) {  // pageable caller memory: stage the luma through pinned memory so the DMA is asynchronous
        copy_luma_host(ctx, reinterpret_cast<uint8_t*>(L.h_in.p), j.in, j.n, j.pitch, j.w, j.h, j.stride);
        src = reinterpret_cast<const uint8_t*>(L.h_in.p);
    }
    // from here on copies from / to the caller's buffers may be in flight: a failure drains the lane before it is reported
    auto drained = [&](int status) { cudaStreamSynchronize(L.stream); return status; };
    if ((rc = copy_luma_async(ctx, d_in, src, j, cudaMemcpyHostToDevice, L.stream))) return drained(rc);
    ctx->ctr.bytes_in += luma_payload(j);
    if (j.op == Op::Equalize) rc = launch_equalize(ctx, L.ws, d_in, d_out, j.n, j.pitch, j.w, j.h, j.stride, UV_SKIP, L.stream);
    else rc = launch_clahe(ctx, L.ws, d_in, d_out, j.n, j.pitch, j.w, j.h, j.stride, j.clip, j.tx, j.ty, UV_SKIP, L.stream);
    if (rc) return drained(rc);
    uint8_t* dst = j.out;
    L.user_out = nullptr;
    if (!out_pinned) {
        dst = reinterpret_cast<uint8_t*>(L.h_out.p);
        L.user_out = j.out;
    }
    if ((rc = copy_luma_async(ctx, dst, d_out, j, cudaMemcpyDeviceToHost, L.stream))) return drained(rc);
    if (cudaMemcpyAsync(L.h_status, ws_status(L.ws), sizeof(uint32_t), cudaMemcpyDeviceToHost, L.stream) != cudaSuccess ||
        cudaEventRecord(L.done, L.stream) != cudaSuccess)
        return drained(fail(ctx, NV12EQ_ERR_CUDA, "queuing the completion of a lane failed: %s", cudaGetErrorString(cudaGetLastError())));
    ctx->ctr.bytes_out += luma_payload(j);
    host_chroma(ctx, L, j);  // overlaps with the GPU work queued above
    L.luma_only = true;
    L.jw = j.w; L.jh = j.h; L.jstride = j.stride; L.jpitch = j.pitch;
    L.out_bytes = span;
    L.frames = j.n;
    L.busy = true;
    return NV12EQ_OK;
}

int lane_submit(nv12eq_ctx* ctx, Lane& L, const Job& j) {
    if (L.busy) return fail(ctx, NV12EQ_ERR_BAD_SLOT, "slot is busy; call nv12eq_wait first");
    if (j.n == 0) return NV12EQ_OK;
    if (j.op == Op::Equalize || j.op == Op::Clahe) return lane_submit_nv12(ctx, L, j);
    // colour path: whole BGR frames
    const size_t span = (size_t)(j.n - 1) * j.pitch + j.frame_bytes;  // bytes covered in the caller's buffers
    int rc;
    if ((rc = dev_reserve(ctx, L.d_in, span, false))) return rc;
    const bool in_place = (j.in == j.out);
    if (!in_place && (rc = dev_reserve(ctx, L.d_out, span, false))) return rc;
    uint8_t* d_in = reinterpret_cast<uint8_t*>(L.d_in.p);
    uint8_t* d_out = in_place ? d_in : reinterpret_cast<uint8_t*>(L.d_out.p);

    if (j.op == Op::ColorClahe && (j.tx < 1 || j.ty < 1 || (long long)j.tx * j.ty > (1 << 20)))
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", j.tx, j.ty);
    const bool in_pinned = is_pinned(j.in), out_pinned = is_pinned(j.out);
    if (!in_pinned && (rc = host_reserve(ctx, L.h_in, span))) return rc;
    if (!out_pinned && (rc = host_reserve(ctx, L.h_out, span))) return rc;
    if ((rc = ws_reserve_frames(ctx, L.ws, j.n))) return rc;
    const uint8_t* src = j.in;
    if (!in_pinned) {  // pageable caller memory: stage through pinned memory so the DMA is asynchronous
        memcpy(L.h_in.p, j.in, span);
        src = reinterpret_cast<const uint8_t*>(L.h_in.p);
    }
    auto drained = [&](int status) { cudaStreamSynchronize(L.stream); return status; };
    auto cuda_fail = [&](const char* what) { return drained(fail(ctx, NV12EQ_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(cudaGetLastError()))); };
    if (cudaMemcpyAsync(d_in, src, span, cudaMemcpyHostToDevice, L.stream) != cudaSuccess) return cuda_fail("upload");
    ctx->ctr.bytes_in += span;
    // With strided rows the padding bytes of the caller's output must survive: the device image starts from them.
    const bool preserve_bgr = !in_place && (j.stride != 3 * j.w);
    if (preserve_bgr) {
        const uint8_t* osrc = j.out;
        if (!out_pinned) {
            memcpy(L.h_out.p, j.out, span);
            osrc = reinterpret_cast<const uint8_t*>(L.h_out.p);
        }
        if (cudaMemcpyAsync(d_out, osrc, span, cudaMemcpyHostToDevice, L.stream) != cudaSuccess) return cuda_fail("upload of the output padding");
        ctx->ctr.bytes_in += span;
    }
    rc = launch_color(ctx, L.ws, d_in, d_out, j.n, j.pitch, j.w, j.h, j.stride, j.color_mode, j.op == Op::ColorClahe, j.clip, j.tx,
                      j.ty, L.stream);
    if (rc) return drained(rc);
    uint8_t* dst = j.out;
    L.user_out = nullptr;
    if (!out_pinned) {
        dst = reinterpret_cast<uint8_t*>(L.h_out.p);
        L.user_out = j.out;
    }
    if (cudaMemcpyAsync(dst, d_out, span, cudaMemcpyDeviceToHost, L.stream) != cudaSuccess ||
        cudaMemcpyAsync(L.h_status, ws_status(L.ws), sizeof(uint32_t), cudaMemcpyDeviceToHost, L.stream) != cudaSuccess ||
        cudaEventRecord(L.done, L.stream) != cudaSuccess)
        return cuda_fail("download");
    ctx->ctr.bytes_out += span;
    L.luma_only = false;
    L.out_bytes = span;
    L.frames = j.n;
    L.busy = true;
    return NV12EQ_OK;
}

// After the lane's stream has been synchronised and its status word copied to L.h_status.
int lane_status(nv12eq_ctx* ctx, Lane& L) {
    if (!L.h_status || *L.h_status == 0) return NV12EQ_OK;
    *L.h_status = 0;
    ws_reset(L.ws, L.stream);
    return fail(ctx, NV12EQ_ERR_CUDA, "kernel dependency wait timed out; the workspace has been reset");
}

int lane_wait(nv12eq_ctx* ctx, Lane& L) {
    if (!L.busy) return NV12EQ_OK;
    L.busy = false;
    if (ctx->pool) ctx->pool->wait(&L.host_pending);  // chroma tasks read / write the caller's buffers: always drain them
    CK(ctx, cudaEventSynchronize(L.done));
    if (*L.h_status != 0) {
        // a kernel gave up waiting (should be impossible); put the workspace back to a known state
        *L.h_status = 0;
        ws_reset(L.ws, L.stream);
        return fail(ctx, NV12EQ_ERR_CUDA, "kernel dependency wait timed out");
    }
    if (L.user_out) {
        if (L.luma_only) copy_luma_host(ctx, L.user_out, reinterpret_cast<const uint8_t*>(L.h_out.p), L.frames, L.jpitch, L.jw, L.jh, L.jstride);
        else memcpy(L.user_out, L.h_out.p, L.out_bytes);
    }
    L.user_out = nullptr;
    ctx->ctr.frames += (uint64_t)L.frames;
    return NV12EQ_OK;
}

int check_host_job(nv12eq_ctx* ctx, const Job& j) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (!j.in || !j.out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null frame pointer");
    if (j.in != j.out) {
        const size_t span = j.n > 0 ? (size_t)(j.n - 1) * j.pitch + j.frame_bytes : 0;
        const uint8_t* a = j.in; const uint8_t* b = j.out;
        if ((a < b && a + span > b) || (b < a && b + span > a)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "input and output overlap partially");
    }
    return NV12EQ_OK;
}

// Synchronous batch through two (or more) lanes: group k uses lane k % L, so the upload of one group overlaps the
// kernels and the download of the previous one.
int run_batch(nv12eq_ctx* ctx, Job j) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = check_host_job(ctx, j);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    const int nl = (int)ctx->lanes.size();
    size_t target = 64ull << 20;   // bytes per group: large enough to amortise the per-group calls, small enough to pipeline
    if (const char* e = getenv("NV12EQ_GROUP_MB")) target = (size_t)std::max(1, atoi(e)) << 20;
    int group = (int)std::max<size_t>(1, std::min<size_t>(64, target / std::max<size_t>(1, j.pitch)));
    if (j.n <= group) group = std::max(1, (j.n + std::min(nl, j.n) - 1) / std::max(1, std::min(nl, j.n)));
    int k = 0, first_err = NV12EQ_OK;
    for (int f0 = 0; f0 < j.n; f0 += group, ++k) {
        Lane& L = ctx->lanes[k % nl];
        rc = lane_wait(ctx, L);
        if (rc && !first_err) first_err = rc;
        Job sub = j;
        sub.n = std::min(group, j.n - f0);
        sub.in = j.in + (size_t)f0 * j.pitch;
        sub.out = j.out + (size_t)f0 * j.pitch;
        rc = lane_submit(ctx, L, sub);
        if (rc) { first_err = first_err ? first_err : rc; break; }
    }
    for (auto& L : ctx->lanes) {
        rc = lane_wait(ctx, L);
        if (rc && !first_err) first_err = rc;
    }
    ctx->ctr.busy_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    return first_err;
}

Job make_nv12_job(Op op, const uint8_t* in, uint8_t* out, int n, size_t pitch, int w, int h, int stride, int uv_mode,
                  double clip, int tx, int ty) {
    Job j{};
    j.op = op; j.in = in; j.out = out; j.n = n;
    j.frame_bytes = (size_t)stride * (size_t)(h + h / 2);
    j.pitch = (n > 1 || pitch) ? pitch : j.frame_bytes;
    if (j.pitch == 0) j.pitch = j.frame_bytes;
    j.w = w; j.h = h; j.stride = stride; j.uv_mode = uv_mode; j.clip = clip; j.tx = tx; j.ty = ty;
    return j;
}

}  // namespace

// =========================================================================================================
// C-ABI
// =========================================================================================================
extern "C" {

int nv12eq_version(void) { return NV12EQ_VERSION_MAJOR * 100 + NV12EQ_VERSION_MINOR; }

const char* nv12eq_status_string(int s) {
    switch (s) {
        case NV12EQ_OK: return "ok";
        case NV12EQ_ERR_INVALID_ARGUMENT: return "invalid argument";
        case NV12EQ_ERR_SHORT_BUFFER: return "buffer too small for the frame";
        case NV12EQ_ERR_CUDA: return "CUDA error";
        case NV12EQ_ERR_NO_DEVICE: return "no usable CUDA device";
        case NV12EQ_ERR_OUT_OF_MEMORY: return "out of memory";
        case NV12EQ_ERR_BAD_SLOT: return "bad or busy slot";
        case NV12EQ_ERR_TOO_LARGE: return "frame too large";
        case NV12EQ_ERR_DROPPED: return "frame dropped by the back-pressure policy";
        case NV12EQ_ERR_EMPTY: return "no frame ready";
        default: return "unknown status";
    }
}

const char* nv12eq_last_error_string(const nv12eq_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int nv12eq_create(int device, int max_width, int max_height, int slots, nv12eq_ctx** out_ctx) {
    if (!out_ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    if (max_width <= 0 || max_height <= 0 || slots < 1 || slots > 64) return fail(nullptr, NV12EQ_ERR_INVALID_ARGUMENT, "bad create arguments");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, NV12EQ_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(nullptr, NV12EQ_ERR_NO_DEVICE, "device %d out of range (0..%d)", device, count - 1);
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (major != 10) return fail(nullptr, NV12EQ_ERR_NO_DEVICE, "device %d is compute capability %d.x; libnv12eq is built for sm_100a only", device, major);
    nv12eq_ctx* ctx = new (std::nothrow) nv12eq_ctx();
    if (!ctx) return NV12EQ_ERR_OUT_OF_MEMORY;
    ctx->device = device; ctx->max_w = max_width; ctx->max_h = max_height; ctx->sm_count = sms;
    DeviceGuard guard(device);
    ctx->lanes.resize(slots);
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& L : ctx->lanes) {
        ok = ok && cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&L.h_status), sizeof(uint32_t), cudaHostAllocDefault) == cudaSuccess;
        if (ok) *L.h_status = 0;
    }
    if (!ok) {
        fail(nullptr, NV12EQ_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        nv12eq_destroy(ctx);
        return NV12EQ_ERR_CUDA;
    }
    *out_ctx = ctx;
    return NV12EQ_OK;
}

void nv12eq_destroy(nv12eq_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    for (auto& L : ctx->lanes) {
        if (L.stream) cudaStreamSynchronize(L.stream);
        dev_release(L.d_in); dev_release(L.d_out); host_release(L.h_in); host_release(L.h_out);
        ws_release(L.ws);
        if (L.h_status) cudaFreeHost(L.h_status);
        if (L.done) cudaEventDestroy(L.done);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->own_stream); cudaStreamDestroy(ctx->own_stream); }
    if (ctx->dev_ws_event) cudaEventDestroy(ctx->dev_ws_event);
    ws_release(ctx->dev_ws);
    delete ctx->pool;
    delete ctx;
}

int nv12eq_get_counters(const nv12eq_ctx* ctx, nv12eq_counters* out) {
    if (!ctx || !out) return NV12EQ_ERR_INVALID_ARGUMENT;
    *out = ctx->ctr;
    return NV12EQ_OK;
}

int nv12eq_set_tuning(nv12eq_ctx* ctx, int chunks_per_frame, int lag_frames, int ctas_per_sm, int schedule) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (chunks_per_frame < 0 || ctas_per_sm < 0 || ctas_per_sm > 8 || schedule < 0 || schedule > 2)
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tuning values");
    ctx->tune_chunks = chunks_per_frame; ctx->tune_lag = lag_frames; ctx->tune_ctas = ctas_per_sm; ctx->tune_schedule = schedule;
    return NV12EQ_OK;
}

int nv12eq_host_alloc(size_t bytes, void** out_ptr) {
    if (!out_ptr || bytes == 0) return NV12EQ_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostAlloc(out_ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *out_ptr = nullptr; return e == cudaErrorMemoryAllocation ? NV12EQ_ERR_OUT_OF_MEMORY : NV12EQ_ERR_CUDA; }
    return NV12EQ_OK;
}
int nv12eq_host_free(void* ptr) {
    if (!ptr) return NV12EQ_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? NV12EQ_OK : NV12EQ_ERR_CUDA;
}

// ---- host frame / batch forms ---------------------------------------------------------------------------
int nv12eq_equalize_hist(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width,
                         int height, int stride, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, 1, 0, uv_mode);
    if (rc) return rc;
    const size_t need = (size_t)stride * (size_t)(height + height / 2);
    if (in_size < need || out_size < need) return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "buffer %zu/%zu bytes < frame %zu bytes", in_size, out_size, need);
    return run_batch(ctx, make_nv12_job(Op::Equalize, in, out, 1, need, width, height, stride, uv_mode, 0, 0, 0));
}

int nv12eq_clahe(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width, int height,
                 int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, 1, 0, uv_mode);
    if (rc) return rc;
    if (tiles_x < 1 || tiles_y < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tiles_x, tiles_y);
    const size_t need = (size_t)stride * (size_t)(height + height / 2);
    if (in_size < need || out_size < need) return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "buffer %zu/%zu bytes < frame %zu bytes", in_size, out_size, need);
    return run_batch(ctx, make_nv12_job(Op::Clahe, in, out, 1, need, width, height, stride, uv_mode, clip_limit, tiles_x, tiles_y));
}

int nv12eq_equalize_hist_batch(nv12eq_ctx* ctx, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch, int width,
                               int height, int stride, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    return run_batch(ctx, make_nv12_job(Op::Equalize, in, out, n_frames, frame_pitch, width, height, stride, uv_mode, 0, 0, 0));
}

int nv12eq_clahe_batch(nv12eq_ctx* ctx, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch, int width,
                       int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    if (tiles_x < 1 || tiles_y < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tiles_x, tiles_y);
    return run_batch(ctx, make_nv12_job(Op::Clahe, in, out, n_frames, frame_pitch, width, height, stride, uv_mode, clip_limit, tiles_x, tiles_y));
}

static int submit_common(nv12eq_ctx* ctx, int slot, const Job& j) {
    if (slot < 0 || slot >= (int)ctx->lanes.size()) return fail(ctx, NV12EQ_ERR_BAD_SLOT, "slot %d out of range", slot);
    int rc = check_host_job(ctx, j);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    return lane_submit(ctx, ctx->lanes[slot], j);
}

int nv12eq_submit_equalize_hist(nv12eq_ctx* ctx, int slot, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch,
                                int width, int height, int stride, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    return submit_common(ctx, slot, make_nv12_job(Op::Equalize, in, out, n_frames, frame_pitch, width, height, stride, uv_mode, 0, 0, 0));
}

int nv12eq_submit_clahe(nv12eq_ctx* ctx, int slot, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch, int width,
                        int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    if (tiles_x < 1 || tiles_y < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tiles_x, tiles_y);
    return submit_common(ctx, slot, make_nv12_job(Op::Clahe, in, out, n_frames, frame_pitch, width, height, stride, uv_mode, clip_limit, tiles_x, tiles_y));
}

int nv12eq_wait(nv12eq_ctx* ctx, int slot) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (slot < 0 || slot >= (int)ctx->lanes.size()) return fail(ctx, NV12EQ_ERR_BAD_SLOT, "slot %d out of range", slot);
    DeviceGuard guard(ctx->device);
    return lane_wait(ctx, ctx->lanes[slot]);
}

int nv12eq_query(nv12eq_ctx* ctx, int slot) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (slot < 0 || slot >= (int)ctx->lanes.size()) return fail(ctx, NV12EQ_ERR_BAD_SLOT, "slot %d out of range", slot);
    Lane& L = ctx->lanes[slot];
    if (!L.busy) return NV12EQ_OK;
    DeviceGuard guard(ctx->device);
    cudaError_t e = cudaEventQuery(L.done);
    if (e == cudaSuccess) return NV12EQ_OK;
    if (e == cudaErrorNotReady) { cudaGetLastError(); return NV12EQ_ERR_BAD_SLOT; }
    return fail(ctx, NV12EQ_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
}

// ---- device forms ---------------------------------------------------------------------------------------
// All *_device calls of a context share one workspace (ticket counter, histograms, tables).  Calls on the same stream are
// ordered by the stream; when the caller switches streams the new stream is made to wait for everything the previous one
// has been given so far, so two streams never run kernels on the workspace at the same time.
static cudaStream_t pick_stream(nv12eq_ctx* ctx, void* s) {
    const cudaStream_t st = s ? reinterpret_cast<cudaStream_t>(s) : ctx->own_stream;
    if (ctx->dev_ws_used && st != ctx->dev_ws_stream) {
        DeviceGuard guard(ctx->device);
        bool ordered = false;
        if (!ctx->dev_ws_event) cudaEventCreateWithFlags(&ctx->dev_ws_event, cudaEventDisableTiming);
        if (ctx->dev_ws_event && cudaEventRecord(ctx->dev_ws_event, ctx->dev_ws_stream) == cudaSuccess)
            ordered = cudaStreamWaitEvent(st, ctx->dev_ws_event, 0) == cudaSuccess;
        if (!ordered) {   // e.g. the previous stream has been destroyed by the caller: its work is complete or will never run
            cudaGetLastError();
            cudaDeviceSynchronize();
        }
    }
    ctx->dev_ws_stream = st;
    ctx->dev_ws_used = true;
    return st;
}

int nv12eq_equalize_hist_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch, int width,
                                int height, int stride, int uv_mode, void* cuda_stream) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    if (!d_in || !d_out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null device pointer");
    DeviceGuard guard(ctx->device);
    rc = launch_equalize(ctx, ctx->dev_ws, d_in, d_out, n_frames, frame_pitch, width, height, stride, uv_mode, pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}

int nv12eq_clahe_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch, int width,
                        int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode, void* cuda_stream) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, uv_mode);
    if (rc) return rc;
    if (!d_in || !d_out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null device pointer");
    DeviceGuard guard(ctx->device);
    rc = launch_clahe(ctx, ctx->dev_ws, d_in, d_out, n_frames, frame_pitch, width, height, stride, clip_limit, tiles_x, tiles_y, uv_mode,
                      pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}

int nv12eq_sync(nv12eq_ctx* ctx) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    CK(ctx, cudaStreamSynchronize(ctx->own_stream));
    // the *_device forms never synchronise, so this is where their kernels' status word is read: wait for the stream of the
    // last device call, fetch the word, and on a (never expected) dependency time-out put the workspace back to zero
    if (ctx->dev_ws_used && ctx->dev_ws.misc.p) {
        if (ctx->dev_ws_stream != ctx->own_stream && cudaStreamSynchronize(ctx->dev_ws_stream) != cudaSuccess) {
            cudaGetLastError();   // the caller destroyed that stream: its work is complete or will never run
            CK(ctx, cudaDeviceSynchronize());
        }
        uint32_t st = 0;
        CK(ctx, cudaMemcpy(&st, ws_status(ctx->dev_ws), sizeof st, cudaMemcpyDeviceToHost));
        if (st != 0) {
            ws_reset(ctx->dev_ws, ctx->own_stream);
            return fail(ctx, NV12EQ_ERR_CUDA, "kernel dependency wait timed out in a *_device call; the workspace has been reset");
        }
    }
    return NV12EQ_OK;
}

int nv12eq_hist_device(nv12eq_ctx* ctx, const uint8_t* d_y, int n_planes, size_t plane_pitch, int width, int height, int stride,
                       uint32_t* d_hist, void* cuda_stream) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (!d_y || !d_hist || width <= 0 || height <= 0 || stride < width || n_planes < 0) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    if (n_planes > 1 && plane_pitch < (size_t)stride * height) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "plane_pitch too small");
    DeviceGuard guard(ctx->device);
    return launch_equalize(ctx, ctx->dev_ws, d_y, nullptr, n_planes, plane_pitch, width, height, stride, UV_SKIP, pick_stream(ctx, cuda_stream),
                           PH_HIST | PH_EXTERNAL_HIST, d_hist);
}

int nv12eq_equalize_apply_device(nv12eq_ctx* ctx, const uint8_t* d_y_in, uint8_t* d_y_out, int n_planes, size_t plane_pitch, int width,
                                 int height, int stride, const uint32_t* d_hist, int64_t total_pixels, void* cuda_stream) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (!d_y_in || !d_y_out || !d_hist || width <= 0 || height <= 0 || stride < width || n_planes < 0 || total_pixels <= 0)
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    if (n_planes > 1 && plane_pitch < (size_t)stride * height) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "plane_pitch too small");
    DeviceGuard guard(ctx->device);
    return launch_equalize(ctx, ctx->dev_ws, d_y_in, d_y_out, n_planes, plane_pitch, width, height, stride, UV_SKIP,
                           pick_stream(ctx, cuda_stream), PH_APPLY | PH_EXTERNAL_HIST, const_cast<uint32_t*>(d_hist), (long long)total_pixels);
}

// ---- spatial split of one frame, CLAHE: band stages (the tile LUT halo exchange between them is the caller's NCCL send/recv) ----
static int check_band(nv12eq_ctx* ctx, int width, int full_height, int stride, int tiles_x, int tiles_y, int first_tile_row, int band_tiles_y) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (width <= 0 || full_height <= 0 || stride < width || tiles_x < 1 || tiles_y < 1 || first_tile_row < 0 || band_tiles_y < 1 ||
        first_tile_row + band_tiles_y > tiles_y)
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad band arguments");
    if (width % tiles_x != 0 || full_height % tiles_y != 0)
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "the spatial split needs a tile grid that divides the frame (%dx%d tiles on %dx%d)", tiles_x, tiles_y,
                    width, full_height);
    if (width > ctx->max_w || full_height > ctx->max_h) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame exceeds the context maximum");
    return NV12EQ_OK;
}
int nv12eq_clahe_band_luts_device(nv12eq_ctx* ctx, const uint8_t* d_y_band, int width, int full_height, int stride, double clip_limit, int tiles_x,
                                  int tiles_y, int first_tile_row, int band_tiles_y, uint8_t* d_luts, void* cuda_stream) {
    int rc = check_band(ctx, width, full_height, stride, tiles_x, tiles_y, first_tile_row, band_tiles_y);
    if (rc) return rc;
    if (!d_y_band || !d_luts || clip_limit < 0) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    const int rows = full_height / tiles_y * band_tiles_y;
    ClaheBand band{ClaheBand::LutsOnly, full_height, tiles_y, first_tile_row, d_luts};
    return launch_clahe(ctx, ctx->dev_ws, d_y_band, nullptr, 1, 0, width, rows, stride, clip_limit, tiles_x, band_tiles_y, UV_SKIP,
                        pick_stream(ctx, cuda_stream), &band);
}
int nv12eq_clahe_band_apply_device(nv12eq_ctx* ctx, const uint8_t* d_y_band, uint8_t* d_out_band, int width, int full_height, int stride, int tiles_x,
                                   int tiles_y, int first_tile_row, int band_tiles_y, const uint8_t* d_luts_halo, void* cuda_stream) {
    int rc = check_band(ctx, width, full_height, stride, tiles_x, tiles_y, first_tile_row, band_tiles_y);
    if (rc) return rc;
    if (!d_y_band || !d_out_band || !d_luts_halo) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    const int rows = full_height / tiles_y * band_tiles_y;
    ClaheBand band{ClaheBand::ApplyOnly, full_height, tiles_y, first_tile_row, const_cast<uint8_t*>(d_luts_halo)};
    return launch_clahe(ctx, ctx->dev_ws, d_y_band, d_out_band, 1, 0, width, rows, stride, 0.0, tiles_x, band_tiles_y, UV_SKIP,
                        pick_stream(ctx, cuda_stream), &band);
}

// ---- colour path ----------------------------------------------------------------------------------------
static int check_color(nv12eq_ctx* ctx, int w, int h, int stride, int mode) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (w <= 0 || h <= 0 || stride < 3 * w) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad BGR geometry w=%d h=%d stride=%d", w, h, stride);
    if (mode != NV12EQ_COLOR_YUV && mode != NV12EQ_COLOR_YCRCB) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad color_mode %d", mode);
    if ((long long)w * h >= (1ll << 31)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame has 2^31 pixels or more");
    if (w > ctx->max_w || h > ctx->max_h) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame %dx%d exceeds context maximum %dx%d", w, h, ctx->max_w, ctx->max_h);
    return NV12EQ_OK;
}

static Job make_color_job(Op op, const uint8_t* in, uint8_t* out, int w, int h, int stride, int mode, double clip, int tx, int ty) {
    Job j{};
    j.op = op; j.in = in; j.out = out; j.n = 1;
    // tight span: the caller's buffer may end right after the last pixel of the last row
    j.frame_bytes = (size_t)stride * (h - 1) + 3 * (size_t)w; j.pitch = j.frame_bytes;
    j.w = w; j.h = h; j.stride = stride; j.uv_mode = UV_SKIP; j.color_mode = mode; j.clip = clip; j.tx = tx; j.ty = ty;
    return j;
}

int nv12eq_color_equalize(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int width, int height, int stride, int color_mode) {
    int rc = check_color(ctx, width, height, stride, color_mode);
    if (rc) return rc;
    return run_batch(ctx, make_color_job(Op::ColorEq, bgr_in, bgr_out, width, height, stride, color_mode, 0, 0, 0));
}

int nv12eq_color_equalize_batch(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int n_frames, size_t frame_pitch, int width,
                                int height, int stride, int color_mode) {
    int rc = check_color(ctx, width, height, stride, color_mode);
    if (rc) return rc;
    if (n_frames < 0 || (n_frames > 1 && frame_pitch < (size_t)stride * height)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad batch arguments");
    Job j = make_color_job(Op::ColorEq, bgr_in, bgr_out, width, height, stride, color_mode, 0, 0, 0);
    j.n = n_frames;
    if (n_frames > 1) j.pitch = frame_pitch;
    return run_batch(ctx, j);
}

int nv12eq_color_clahe(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int width, int height, int stride, int color_mode,
                       double clip_limit, int tiles_x, int tiles_y) {
    int rc = check_color(ctx, width, height, stride, color_mode);
    if (rc) return rc;
    if (tiles_x < 1 || tiles_y < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tiles_x, tiles_y);
    return run_batch(ctx, make_color_job(Op::ColorClahe, bgr_in, bgr_out, width, height, stride, color_mode, clip_limit, tiles_x, tiles_y));
}

int nv12eq_color_equalize_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch, int width,
                                 int height, int stride, int color_mode, void* cuda_stream) {
    int rc = check_color(ctx, width, height, stride, color_mode);
    if (rc) return rc;
    if (!d_in || !d_out || n_frames < 0 || (n_frames > 1 && frame_pitch < (size_t)stride * height)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    rc = launch_color(ctx, ctx->dev_ws, d_in, d_out, n_frames, frame_pitch, width, height, stride, color_mode, false, 0, 0, 0,
                      pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}

int nv12eq_color_clahe_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch, int width,
                              int height, int stride, int color_mode, double clip_limit, int tiles_x, int tiles_y, void* cuda_stream) {
    int rc = check_color(ctx, width, height, stride, color_mode);
    if (rc) return rc;
    if (!d_in || !d_out || n_frames < 0 || (n_frames > 1 && frame_pitch < (size_t)stride * height)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    if (tiles_x < 1 || tiles_y < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tiles_x, tiles_y);
    DeviceGuard guard(ctx->device);
    rc = launch_color(ctx, ctx->dev_ws, d_in, d_out, n_frames, frame_pitch, width, height, stride, color_mode, true, clip_limit, tiles_x,
                      tiles_y, pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}


// ---- frames with per-plane offsets / strides -----------------------------------------------------------
// The DMA engine does the re-layout: the luma rows are gathered into a flat device plane (so the kernels always take
// their contiguous fast path), processed with UV_SKIP, and scattered back with the output's stride; the chroma rows are
// copied / filled on the host like in the packed entry points.
static int meta_frame(nv12eq_ctx* ctx, Op op, const uint8_t* in, size_t in_size, const nv12eq_layout* il, uint8_t* out, size_t out_size,
                      const nv12eq_layout* ol, int w, int h, double clip, int tx, int ty, int uv_mode) {
    int rc = check_geometry(ctx, w, h, w, 1, 0, uv_mode);
    if (rc) return rc;
    if (!in || !out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null frame pointer");
    if (op == Op::Clahe && (tx < 1 || ty < 1)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tx, ty);
    const nv12eq_layout packed{{0, (size_t)w * h}, {w, w}};
    const nv12eq_layout& I = il ? *il : packed;
    const nv12eq_layout& O = ol ? *ol : packed;
    const int uvh = h / 2;
    for (const nv12eq_layout* L : {&I, &O})
        if (L->stride[0] < w || L->stride[1] < w) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "plane stride smaller than the width");
    auto plane_end = [&](const nv12eq_layout& L, int pl, int rows) { return rows > 0 ? L.offset[pl] + (size_t)L.stride[pl] * (rows - 1) + w : (size_t)0; };
    const bool in_uv_needed = (uv_mode == UV_COPY);
    if (plane_end(I, 0, h) > in_size || (in_uv_needed && plane_end(I, 1, uvh) > in_size) || plane_end(O, 0, h) > out_size ||
        (uv_mode != UV_SKIP && plane_end(O, 1, uvh) > out_size))
        return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "a plane described by the layout does not fit in its buffer");
    const auto t0 = std::chrono::steady_clock::now();
    DeviceGuard guard(ctx->device);
    Lane& L = ctx->lanes[0];
    if ((rc = lane_wait(ctx, L))) return rc;
    const size_t plane = (size_t)w * h;
    if ((rc = dev_reserve(ctx, L.d_in, plane, false))) return rc;
    if ((rc = dev_reserve(ctx, L.d_out, plane, false))) return rc;
    uint8_t* d_in = reinterpret_cast<uint8_t*>(L.d_in.p);
    uint8_t* d_out = reinterpret_cast<uint8_t*>(L.d_out.p);
    CK(ctx, cudaMemcpy2DAsync(d_in, (size_t)w, in + I.offset[0], (size_t)I.stride[0], (size_t)w, (size_t)h, cudaMemcpyHostToDevice, L.stream));
    rc = (op == Op::Clahe) ? launch_clahe(ctx, L.ws, d_in, d_out, 1, plane, w, h, w, clip, tx, ty, UV_SKIP, L.stream)
                           : launch_equalize(ctx, L.ws, d_in, d_out, 1, plane, w, h, w, UV_SKIP, L.stream);
    if (rc) return rc;
    CK(ctx, cudaMemcpy2DAsync(out + O.offset[0], (size_t)O.stride[0], d_out, (size_t)w, (size_t)w, (size_t)h, cudaMemcpyDeviceToHost, L.stream));
    // chroma on the host while the GPU works (rows may alias when in == out with identical layouts: then nothing to do)
    if (uvh > 0 && uv_mode != UV_SKIP) {
        const uint8_t* src = in + I.offset[1];
        uint8_t* dst = out + O.offset[1];
        const bool same = (uv_mode == UV_COPY && src == dst && I.stride[1] == O.stride[1]);
        if (!same)
            for (int r = 0; r < uvh; ++r) {
                if (uv_mode == UV_COPY) memmove(dst + (size_t)r * O.stride[1], src + (size_t)r * I.stride[1], (size_t)w);
                else memset(dst + (size_t)r * O.stride[1], 128, (size_t)w);
            }
    }
    CK(ctx, cudaMemcpyAsync(L.h_status, ws_status(L.ws), sizeof(uint32_t), cudaMemcpyDeviceToHost, L.stream));
    CK(ctx, cudaStreamSynchronize(L.stream));
    if ((rc = lane_status(ctx, L))) return rc;
    ctx->ctr.bytes_in += plane; ctx->ctr.bytes_out += plane; ctx->ctr.frames++;
    ctx->ctr.busy_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    return NV12EQ_OK;
}

int nv12eq_equalize_hist_meta(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, const nv12eq_layout* in_layout, uint8_t* out,
                              size_t out_size, const nv12eq_layout* out_layout, int width, int height, int uv_mode) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    return meta_frame(ctx, Op::Equalize, in, in_size, in_layout, out, out_size, out_layout, width, height, 0.0, 0, 0, uv_mode);
}

int nv12eq_clahe_meta(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, const nv12eq_layout* in_layout, uint8_t* out, size_t out_size,
                      const nv12eq_layout* out_layout, int width, int height, double clip_limit, int tiles_x, int tiles_y, int uv_mode) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    return meta_frame(ctx, Op::Clahe, in, in_size, in_layout, out, out_size, out_layout, width, height, clip_limit, tiles_x, tiles_y, uv_mode);
}

// ---- 16-bit CLAHE ---------------------------------------------------------------------------------------
static int launch_clahe16(nv12eq_ctx* ctx, Workspace& ws, const uint16_t* d_in, uint16_t* d_out, int n, size_t pitch, int w, int h,
                          int stride, double clip, int tx, int ty, cudaStream_t st) {
    if (n == 0) return NV12EQ_OK;
    if (tx < 1 || ty < 1 || (long long)tx * ty > 4096) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad 16-bit tile grid %dx%d (at most 4096 tiles)", tx, ty);
    const int T = tx * ty;
    int extW = w, extH = h;
    if (w % tx != 0 || h % ty != 0) { extW = w + (tx - (w % tx)); extH = h + (ty - (h % ty)); }
    Clahe16Params p{};
    p.w = w; p.h = h; p.stride = stride; p.pitch = pitch;
    p.tx = tx; p.ty = ty; p.tw = extW / tx; p.th = extH / ty;
    const int area = p.tw * p.th;
    p.clip_limit = clip > 0.0 ? std::max(1, (int)(clip * area / 65536.0)) : 0;
    p.lut_scale = 65535.0f / (float)area;
    p.inv_tw = 1.0f / (float)p.tw; p.inv_th = 1.0f / (float)p.th;
    if (p.tw > kC16StripPixels) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "16-bit tile rows of %d pixels are not supported", p.tw);
    // Planes per pass.  Histograms (256 KB per tile) and LUTs (128 KB per tile) of a pass should stay L2-resident between the
    // histogram and the LUT kernel: up to 256 tiles per pass, which also fills the GPU with LUT clusters.  The cell tables
    // (512 KB per cell) are gathered from at random by the blend and must not spill to HBM: about 64 tiles' worth per pass.
    const int cells = (tx + 1) * (ty + 1);
    const int group = std::max(1, std::min(n, 256 / T));
    const int cgroup = std::max(1, std::min(group, 64 / T));
    int rc = ensure_attrs(ctx);
    if (rc) return rc;
    rc = dev_reserve(ctx, ws.hist16, (size_t)group * T * kBins16 * sizeof(uint32_t), true);
    if (rc) return rc;
    if ((rc = dev_reserve(ctx, ws.luts16, (size_t)group * T * kBins16 * sizeof(uint16_t), false))) return rc;
    if ((rc = dev_reserve(ctx, ws.cells16, (size_t)cgroup * cells * kBins16 * sizeof(uint2), false))) return rc;
    if ((rc = dev_reserve(ctx, ws.ormask16, (size_t)group * sizeof(uint32_t), false))) return rc;
    p.hist = reinterpret_cast<uint32_t*>(ws.hist16.p);
    p.cells = reinterpret_cast<uint2*>(ws.cells16.p);
    uint32_t* const ormask = reinterpret_cast<uint32_t*>(ws.ormask16.p);
    uint16_t* const luts = reinterpret_cast<uint16_t*>(ws.luts16.p);
    for (int g0 = 0; g0 < n; g0 += group) {
        const int ng = std::min(group, n - g0);
        p.in = d_in + (size_t)g0 * pitch; p.out = d_out + (size_t)g0 * pitch; p.n_planes = ng; p.luts = luts; p.ormask = ormask;
        CK(ctx, cudaMemsetAsync(ormask, 0, (size_t)ng * sizeof(uint32_t), st));
        // strips: short enough for 16-bit counters; one CTA runs per SM, so as many strips as fit into one wave (every
        // further strip costs another zeroing and flush of the 128 KB counter table)
        const int max_rows = std::max(1, kC16StripPixels / p.tw);
        const long long want = std::max<long long>(1, (long long)ctx->sm_count / ((long long)T * ng));
        p.rows_strip = (int)std::max<long long>(1, std::min<long long>(max_rows, (p.th + want - 1) / want));
        p.strips = (p.th + p.rows_strip - 1) / p.rows_strip;
        clahe16_hist_kernel<<<dim3(p.strips, T, ng), kC16HistThreads, kC16HistSmemBytes, st>>>(p);
        clahe16_lut_kernel<<<dim3(kC16Parts, T, ng), kC16LutThreads, 0, st>>>(p);
        ctx->ctr.kernel_launches += 2;
        for (int c0 = 0; c0 < ng; c0 += cgroup) {
            const int nc = std::min(cgroup, ng - c0);
            Clahe16Params q = p;
            q.in = p.in + (size_t)c0 * pitch; q.out = p.out + (size_t)c0 * pitch; q.n_planes = nc;
            q.luts = luts + (size_t)c0 * T * kBins16; q.ormask = ormask + c0;
            clahe16_cell_table_kernel<<<dim3(8, cells, nc), kC16Threads, 0, st>>>(q);
            clahe16_interp_kernel<<<dim3((w + kC16Threads - 1) / kC16Threads, (h + kC16RowsPerCta - 1) / kC16RowsPerCta, nc), kC16Threads, 0, st>>>(q);
            ctx->ctr.kernel_launches += 2;
        }
        CK(ctx, cudaGetLastError());
    }
    return NV12EQ_OK;
}

static int check_plane16(nv12eq_ctx* ctx, int w, int h, int stride, int tx, int ty) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (w <= 0 || h <= 0 || stride < w) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad 16-bit geometry w=%d h=%d stride=%d", w, h, stride);
    if (h > 65535) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "at most 65535 rows");
    if (w > ctx->max_w || h > ctx->max_h) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame %dx%d exceeds context maximum %dx%d", w, h, ctx->max_w, ctx->max_h);
    if (tx < 1 || ty < 1) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", tx, ty);
    return NV12EQ_OK;
}

int nv12eq_clahe16_device(nv12eq_ctx* ctx, const uint16_t* d_in, uint16_t* d_out, int n_planes, size_t plane_pitch, int width, int height,
                          int stride, double clip_limit, int tiles_x, int tiles_y, void* cuda_stream) {
    int rc = check_plane16(ctx, width, height, stride, tiles_x, tiles_y);
    if (rc) return rc;
    if (!d_in || !d_out || n_planes < 0 || n_planes > 65535 || (n_planes > 1 && plane_pitch < (size_t)stride * height))
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    rc = launch_clahe16(ctx, ctx->dev_ws, d_in, d_out, n_planes, plane_pitch, width, height, stride, clip_limit, tiles_x, tiles_y,
                        pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_planes;
    return rc;
}

// one 16-bit plane between host buffers with row strides in bytes; the device plane is packed
static int plane16_host(nv12eq_ctx* ctx, const uint8_t* in, size_t in_stride, uint8_t* out, size_t out_stride, int w, int h, double clip,
                        int tx, int ty) {
    DeviceGuard guard(ctx->device);
    Lane& L = ctx->lanes[0];
    int rc = lane_wait(ctx, L);
    if (rc) return rc;
    const size_t plane = (size_t)w * h * sizeof(uint16_t);
    if ((rc = dev_reserve(ctx, L.d_in, plane, false))) return rc;
    if ((rc = dev_reserve(ctx, L.d_out, plane, false))) return rc;
    CK(ctx, cudaMemcpy2DAsync(L.d_in.p, (size_t)w * 2, in, in_stride, (size_t)w * 2, (size_t)h, cudaMemcpyHostToDevice, L.stream));
    rc = launch_clahe16(ctx, L.ws, reinterpret_cast<const uint16_t*>(L.d_in.p), reinterpret_cast<uint16_t*>(L.d_out.p), 1, (size_t)w * h, w, h,
                        w, clip, tx, ty, L.stream);
    if (rc) return rc;
    CK(ctx, cudaMemcpy2DAsync(out, out_stride, L.d_out.p, (size_t)w * 2, (size_t)w * 2, (size_t)h, cudaMemcpyDeviceToHost, L.stream));
    ctx->ctr.bytes_in += plane; ctx->ctr.bytes_out += plane;
    return NV12EQ_OK;
}

int nv12eq_clahe16(nv12eq_ctx* ctx, const uint16_t* in, uint16_t* out, int width, int height, int stride, double clip_limit, int tiles_x,
                   int tiles_y) {
    int rc = check_plane16(ctx, width, height, stride, tiles_x, tiles_y);
    if (rc) return rc;
    if (!in || !out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null plane pointer");
    rc = plane16_host(ctx, reinterpret_cast<const uint8_t*>(in), (size_t)stride * 2, reinterpret_cast<uint8_t*>(out), (size_t)stride * 2, width,
                      height, clip_limit, tiles_x, tiles_y);
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->lanes[0].stream));
    ctx->ctr.frames++;
    return NV12EQ_OK;
}

int nv12eq_p010_clahe(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width, int height, int stride,
                      double clip_limit, int tiles_x, int tiles_y, int uv_mode) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (width <= 0 || height <= 0 || stride < 2 * width || (stride & 1)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad P010 geometry w=%d h=%d stride=%d", width, height, stride);
    if (uv_mode < 0 || uv_mode > 2) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad uv_mode %d", uv_mode);
    int rc = check_plane16(ctx, width, height, stride / 2, tiles_x, tiles_y);
    if (rc) return rc;
    if (!in || !out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null frame pointer");
    const size_t need = (size_t)stride * (size_t)(height + height / 2);
    if (in_size < need || out_size < need) return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "buffer %zu/%zu bytes < P010 frame %zu bytes", in_size, out_size, need);
    rc = plane16_host(ctx, in, (size_t)stride, out, (size_t)stride, width, height, clip_limit, tiles_x, tiles_y);
    if (rc) return rc;
    // chroma on the host while the GPU works (neutral chroma of P010 = 512 << 6)
    const size_t off = (size_t)stride * height;
    for (int r = 0; r < height / 2 && uv_mode != UV_SKIP; ++r) {
        uint8_t* d = out + off + (size_t)r * stride;
        if (uv_mode == UV_COPY) { if (in != out) memcpy(d, in + off + (size_t)r * stride, (size_t)width * 2); }
        else { uint16_t* d16 = reinterpret_cast<uint16_t*>(d); for (int c = 0; c < width; ++c) d16[c] = 0x8000; }
    }
    CK(ctx, cudaStreamSynchronize(ctx->lanes[0].stream));
    ctx->ctr.frames++;
    return NV12EQ_OK;
}

// ---- BGR -> I420 adapter --------------------------------------------------------------------------------
static int launch_i420(nv12eq_ctx* ctx, const uint8_t* d_bgr, uint8_t* d_out, int n, size_t bgr_pitch, size_t out_pitch, int w, int h,
                       int stride, cudaStream_t st) {
    if (n == 0) return NV12EQ_OK;
    if (n > 65535) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "at most 65535 frames per call");
    I420Params p{};
    p.bgr = d_bgr; p.out = d_out; p.bgr_pitch = bgr_pitch; p.out_pitch = out_pitch; p.w = w; p.h = h; p.stride = stride;
    const long long blocks = (long long)((w + 3) / 4) * (h / 2);
    const int gx = (int)std::max<long long>(1, std::min<long long>((blocks + kColorThreads - 1) / kColorThreads, (long long)ctx->sm_count * 8));
    bgr_to_i420_kernel<<<dim3(gx, n), kColorThreads, 0, st>>>(p);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}
static int check_i420(nv12eq_ctx* ctx, int w, int h, int stride) {
    int rc = check_color(ctx, w, h, stride, NV12EQ_COLOR_YUV);
    if (rc) return rc;
    if ((w & 1) || (h & 1)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "BGR->I420 needs even width and height, got %dx%d", w, h);
    return NV12EQ_OK;
}

int nv12eq_bgr_to_i420_device(nv12eq_ctx* ctx, const uint8_t* d_bgr, uint8_t* d_out, int n_frames, size_t bgr_pitch, size_t out_pitch,
                              int width, int height, int stride, void* cuda_stream) {
    int rc = check_i420(ctx, width, height, stride);
    if (rc) return rc;
    const size_t out_frame = (size_t)width * height * 3 / 2;
    if (!d_bgr || !d_out || n_frames < 0 || (n_frames > 1 && (bgr_pitch < (size_t)stride * height || out_pitch < out_frame)))
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    rc = launch_i420(ctx, d_bgr, d_out, n_frames, bgr_pitch, out_pitch, width, height, stride, pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}

int nv12eq_bgr_to_i420(nv12eq_ctx* ctx, const uint8_t* bgr, int width, int height, int stride, uint8_t* out, size_t out_size) {
    int rc = check_i420(ctx, width, height, stride);
    if (rc) return rc;
    if (!bgr || !out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null frame pointer");
    const size_t out_frame = (size_t)width * height * 3 / 2;
    if (out_size < out_frame) return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "output %zu bytes < I420 frame %zu bytes", out_size, out_frame);
    DeviceGuard guard(ctx->device);
    Lane& L = ctx->lanes[0];
    if ((rc = lane_wait(ctx, L))) return rc;
    const size_t in_span = (size_t)stride * (height - 1) + 3 * (size_t)width;
    if ((rc = dev_reserve(ctx, L.d_in, in_span, false))) return rc;
    if ((rc = dev_reserve(ctx, L.d_out, out_frame, false))) return rc;
    CK(ctx, cudaMemcpyAsync(L.d_in.p, bgr, in_span, cudaMemcpyHostToDevice, L.stream));
    rc = launch_i420(ctx, reinterpret_cast<const uint8_t*>(L.d_in.p), reinterpret_cast<uint8_t*>(L.d_out.p), 1, in_span, out_frame, width,
                     height, stride, L.stream);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(out, L.d_out.p, out_frame, cudaMemcpyDeviceToHost, L.stream));
    CK(ctx, cudaStreamSynchronize(L.stream));
    ctx->ctr.bytes_in += in_span; ctx->ctr.bytes_out += out_frame; ctx->ctr.frames++;
    return NV12EQ_OK;
}

// ---- NV12 <-> BGR adapters ------------------------------------------------------------------------------
static size_t nv12_frame_bytes(int stride, int h) { return (size_t)stride * (size_t)(h + h / 2); }
static int check_nv12_bgr(nv12eq_ctx* ctx, int w, int h, int stride, int bgr_stride) {
    if (!ctx) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (w <= 0 || h <= 0 || stride < w || bgr_stride < 3 * w)
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad geometry w=%d h=%d stride=%d bgr_stride=%d", w, h, stride, bgr_stride);
    if ((w & 1) || (h & 1)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "NV12 <-> BGR needs even width and height, got %dx%d", w, h);
    if ((long long)w * h >= (1ll << 31)) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame has 2^31 pixels or more");
    if (w > ctx->max_w || h > ctx->max_h) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "frame %dx%d exceeds context maximum %dx%d", w, h, ctx->max_w, ctx->max_h);
    return NV12EQ_OK;
}
static int launch_nv12_bgr(nv12eq_ctx* ctx, bool to_bgr, const uint8_t* d_in, uint8_t* d_out, int n, size_t in_pitch, size_t out_pitch, int w,
                           int h, int stride, int bgr_stride, cudaStream_t st) {
    if (n == 0) return NV12EQ_OK;
    if (n > 65535) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "at most 65535 frames per call");
    Nv12BgrParams p{};
    p.in = d_in; p.out = d_out; p.in_pitch = in_pitch; p.out_pitch = out_pitch; p.w = w; p.h = h; p.nv12_stride = stride; p.bgr_stride = bgr_stride;
    const long long blocks = (long long)((w + 3) / 4) * (h / 2);
    const int gx = (int)std::max<long long>(1, std::min<long long>((blocks + kColorThreads - 1) / kColorThreads, (long long)ctx->sm_count * 8));
    if (to_bgr) nv12_to_bgr_kernel<<<dim3(gx, n), kColorThreads, 0, st>>>(p);
    else bgr_to_nv12_kernel<<<dim3(gx, n), kColorThreads, 0, st>>>(p);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}
static int nv12_bgr_device(nv12eq_ctx* ctx, bool to_bgr, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t nv12_pitch, size_t bgr_pitch,
                           int width, int height, int stride, int bgr_stride, void* cuda_stream) {
    int rc = check_nv12_bgr(ctx, width, height, stride, bgr_stride);
    if (rc) return rc;
    if (!d_in || !d_out || n_frames < 0 ||
        (n_frames > 1 && (nv12_pitch < nv12_frame_bytes(stride, height) || bgr_pitch < (size_t)bgr_stride * height)))
        return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    DeviceGuard guard(ctx->device);
    rc = launch_nv12_bgr(ctx, to_bgr, d_in, d_out, n_frames, to_bgr ? nv12_pitch : bgr_pitch, to_bgr ? bgr_pitch : nv12_pitch, width, height, stride,
                         bgr_stride, pick_stream(ctx, cuda_stream));
    if (!rc) ctx->ctr.frames += (uint64_t)n_frames;
    return rc;
}
static int nv12_bgr_host(nv12eq_ctx* ctx, bool to_bgr, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width, int height,
                         int stride, int bgr_stride) {
    int rc = check_nv12_bgr(ctx, width, height, stride, bgr_stride);
    if (rc) return rc;
    if (!in || !out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null frame pointer");
    const size_t nv12_need = nv12_frame_bytes(stride, height);
    const size_t bgr_need = (size_t)bgr_stride * (height - 1) + 3 * (size_t)width;   // the buffer may end with the last pixel
    const size_t in_need = to_bgr ? nv12_need : bgr_need, out_need = to_bgr ? bgr_need : nv12_need;
    if (in_size < in_need || out_size < out_need)
        return fail(ctx, NV12EQ_ERR_SHORT_BUFFER, "buffers %zu / %zu bytes < %zu / %zu bytes", in_size, out_size, in_need, out_need);
    DeviceGuard guard(ctx->device);
    Lane& L = ctx->lanes[0];
    if ((rc = lane_wait(ctx, L))) return rc;
    if ((rc = dev_reserve(ctx, L.d_in, in_need, false))) return rc;
    if ((rc = dev_reserve(ctx, L.d_out, out_need, false))) return rc;
    CK(ctx, cudaMemcpyAsync(L.d_in.p, in, in_need, cudaMemcpyHostToDevice, L.stream));
    // rows of a strided output keep whatever the device buffer held between them: upload the caller's bytes first when there are gaps
    const bool gaps = to_bgr ? (bgr_stride != 3 * width) : (stride != width);
    if (gaps) CK(ctx, cudaMemcpyAsync(L.d_out.p, out, out_need, cudaMemcpyHostToDevice, L.stream));
    rc = launch_nv12_bgr(ctx, to_bgr, reinterpret_cast<const uint8_t*>(L.d_in.p), reinterpret_cast<uint8_t*>(L.d_out.p), 1, in_need, out_need, width,
                         height, stride, bgr_stride, L.stream);
    if (rc) { cudaStreamSynchronize(L.stream); return rc; }
    CK(ctx, cudaMemcpyAsync(out, L.d_out.p, out_need, cudaMemcpyDeviceToHost, L.stream));
    CK(ctx, cudaStreamSynchronize(L.stream));
    ctx->ctr.bytes_in += in_need; ctx->ctr.bytes_out += out_need; ctx->ctr.frames++;
    return NV12EQ_OK;
}
int nv12eq_nv12_to_bgr(nv12eq_ctx* ctx, const uint8_t* nv12, size_t nv12_size, int width, int height, int stride, uint8_t* bgr, size_t bgr_size,
                       int bgr_stride) {
    return nv12_bgr_host(ctx, true, nv12, nv12_size, bgr, bgr_size, width, height, stride, bgr_stride);
}
int nv12eq_bgr_to_nv12(nv12eq_ctx* ctx, const uint8_t* bgr, size_t bgr_size, int width, int height, int bgr_stride, uint8_t* nv12, size_t nv12_size,
                       int stride) {
    return nv12_bgr_host(ctx, false, bgr, bgr_size, nv12, nv12_size, width, height, stride, bgr_stride);
}
int nv12eq_nv12_to_bgr_device(nv12eq_ctx* ctx, const uint8_t* d_nv12, uint8_t* d_bgr, int n_frames, size_t nv12_pitch, size_t bgr_pitch, int width,
                              int height, int stride, int bgr_stride, void* cuda_stream) {
    return nv12_bgr_device(ctx, true, d_nv12, d_bgr, n_frames, nv12_pitch, bgr_pitch, width, height, stride, bgr_stride, cuda_stream);
}
int nv12eq_bgr_to_nv12_device(nv12eq_ctx* ctx, const uint8_t* d_bgr, uint8_t* d_nv12, int n_frames, size_t bgr_pitch, size_t nv12_pitch, int width,
                              int height, int bgr_stride, int stride, void* cuda_stream) {
    return nv12_bgr_device(ctx, false, d_bgr, d_nv12, n_frames, nv12_pitch, bgr_pitch, width, height, stride, bgr_stride, cuda_stream);
}

// ---- ordered, back-pressured frame stream ---------------------------------------------------------------
// FIFO ring of `depth` lanes.  push() takes the tail lane, pop() the head lane, so delivery order == push order by
// construction; sequence numbers make drops visible to the consumer.
}  // extern "C"

struct nv12eq_stream {
    nv12eq_ctx* ctx = nullptr;
    nv12eq_stream_config cfg{};
    size_t frame_bytes = 0;
    std::vector<Lane> lanes;
    std::vector<uint64_t> seq;                                   // sequence number held by each lane
    std::vector<std::chrono::steady_clock::time_point> t_push;   // push time of each lane's frame
    size_t head = 0, count = 0;                                  // ring state (guarded by mu)
    uint64_t next_seq = 0;
    nv12eq_stream_stats st{};
    std::mutex mu;
    std::condition_variable cv;
    std::string last_error;
};

namespace {
int stream_fail(nv12eq_stream* s, int status, const char* msg) {
    if (s) { s->last_error = msg; if (s->ctx) s->ctx->last_error = msg; }
    return status;
}
Job stream_job(const nv12eq_stream* s, const Lane& L) {
    const nv12eq_stream_config& c = s->cfg;
    return make_nv12_job(c.op == NV12EQ_OP_CLAHE ? Op::Clahe : Op::Equalize, reinterpret_cast<const uint8_t*>(L.h_in.p),
                         reinterpret_cast<uint8_t*>(L.h_out.p), 1, s->frame_bytes, c.width, c.height, c.stride, c.uv_mode, c.clip_limit,
                         c.tiles_x, c.tiles_y);
}
// Launch one frame on a lane whose pinned input already holds the luma plane: only the luma crosses PCIe.
int stream_launch(nv12eq_stream* s, Lane& L) {
    nv12eq_ctx* ctx = s->ctx;
    const Job j = stream_job(s, L);
    uint8_t* d_in = reinterpret_cast<uint8_t*>(L.d_in.p);
    uint8_t* d_out = reinterpret_cast<uint8_t*>(L.d_out.p);
    int rc = copy_luma_async(ctx, d_in, j.in, j, cudaMemcpyHostToDevice, L.stream);
    if (rc) return rc;
    rc = (j.op == Op::Clahe) ? launch_clahe(ctx, L.ws, d_in, d_out, 1, j.pitch, j.w, j.h, j.stride, j.clip, j.tx, j.ty, UV_SKIP, L.stream)
                             : launch_equalize(ctx, L.ws, d_in, d_out, 1, j.pitch, j.w, j.h, j.stride, UV_SKIP, L.stream);
    if (rc) return rc;
    if ((rc = copy_luma_async(ctx, j.out, d_out, j, cudaMemcpyDeviceToHost, L.stream))) return rc;
    CK(ctx, cudaMemcpyAsync(L.h_status, ws_status(L.ws), sizeof(uint32_t), cudaMemcpyDeviceToHost, L.stream));
    CK(ctx, cudaEventRecord(L.done, L.stream));
    return NV12EQ_OK;
}
// rows of `w` payload bytes between two frames of the stream's layout: r0 <= row < r1 (luma rows 0..h, chroma rows h..h+h/2)
// Large frames are cut into row blocks for the context's host pool (the calling thread takes one block itself): a single
// thread copies a 4K frame in ~1.2 ms, which was most of the push -> pop latency of a 4K stream.
void stream_copy_rows(const nv12eq_stream* s, uint8_t* dst, const uint8_t* src, int r0, int r1) {
    const size_t stride = (size_t)s->cfg.stride, w = (size_t)s->cfg.width;
    const int rows = r1 - r0;
    if (rows <= 0) return;
    // (a stream has a producer and a consumer thread copying at the same time: the pool pays from a quarter of the size at which it
    //  pays for a single caller -- 1080p: p50 0.42 vs 0.60 ms pooled, while a lone 1080p host call is 0.34 ms inline vs 0.53 ms pooled)
    HostPool* pool = ((size_t)rows * w >= pool_min_bytes() / 4) ? host_pool(s->ctx) : nullptr;
    const int parts = pool ? std::min(s->ctx->pool_threads + 1, std::min(8, rows)) : 1;
    int pending = 0;
    const bool flat = stride == w;
    for (int k = parts - 1; k >= 0; --k) {   // block 0 last: it is the caller's own
        const int a = r0 + (int)((long long)rows * k / parts), b = r0 + (int)((long long)rows * (k + 1) / parts);
        HostTask t{dst + (size_t)a * stride, src + (size_t)a * stride, flat ? 1 : b - a, flat ? (size_t)(b - a) * stride : w, stride, 0, &pending};
        if (k > 0) pool->submit(t);
        else HostPool::execute(t);
    }
    if (pool && parts > 1) pool->wait(&pending);
}
}  // namespace

extern "C" {

int nv12eq_stream_open(nv12eq_ctx* ctx, const nv12eq_stream_config* cfg, nv12eq_stream** out_stream) {
    if (!ctx || !cfg || !out_stream) return NV12EQ_ERR_INVALID_ARGUMENT;
    *out_stream = nullptr;
    int rc = check_geometry(ctx, cfg->width, cfg->height, cfg->stride, 1, 0, cfg->uv_mode);
    if (rc) return rc;
    if (cfg->op != NV12EQ_OP_EQUALIZE && cfg->op != NV12EQ_OP_CLAHE) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad stream op %d", cfg->op);
    if (cfg->op == NV12EQ_OP_CLAHE && (cfg->tiles_x < 1 || cfg->tiles_y < 1)) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad tile grid %dx%d", cfg->tiles_x, cfg->tiles_y);
    if (cfg->depth < 1 || cfg->depth > 64) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "stream depth %d out of range 1..64", cfg->depth);
    if (cfg->full_policy < 0 || cfg->full_policy > 2) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad full_policy %d", cfg->full_policy);
    nv12eq_stream* s = new (std::nothrow) nv12eq_stream();
    if (!s) return NV12EQ_ERR_OUT_OF_MEMORY;
    host_pool(ctx);   // created here, before the producer and the consumer thread can race for it (stream_copy_rows)
    s->ctx = ctx; s->cfg = *cfg;
    s->frame_bytes = (size_t)cfg->stride * (size_t)(cfg->height + cfg->height / 2);
    s->lanes.resize(cfg->depth); s->seq.assign(cfg->depth, 0); s->t_push.resize(cfg->depth);
    DeviceGuard guard(ctx->device);
    // With UV_SKIP or strided frames some output bytes are never written by the kernels: start them at zero.
    for (auto& L : s->lanes) {
        bool ok = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&L.h_status), sizeof(uint32_t), cudaHostAllocDefault) == cudaSuccess;
        if (ok) *L.h_status = 0;
        if (ok) rc = dev_reserve(ctx, L.d_in, s->frame_bytes, false);
        if (ok && !rc) rc = dev_reserve(ctx, L.d_out, s->frame_bytes, true);
        if (ok && !rc) rc = host_reserve(ctx, L.h_in, s->frame_bytes);
        if (ok && !rc) rc = host_reserve(ctx, L.h_out, s->frame_bytes);
        if (ok && !rc) {
            // the staged output frame: payload bytes the stream never writes start at zero; neutral chroma is written once
            memset(L.h_out.p, 0, s->frame_bytes);
            if (cfg->uv_mode == UV_GRAY128) {
                uint8_t* o = reinterpret_cast<uint8_t*>(L.h_out.p);
                for (int r = cfg->height; r < cfg->height + cfg->height / 2; ++r) memset(o + (size_t)r * cfg->stride, 128, (size_t)cfg->width);
            }
        }
        if (!ok || rc) {
            if (!rc) rc = fail(ctx, NV12EQ_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            nv12eq_stream_close(s);
            return rc;
        }
    }
    *out_stream = s;
    return NV12EQ_OK;
}

int nv12eq_stream_push(nv12eq_stream* s, const uint8_t* in, size_t in_size, uint64_t* out_seq) {
    if (!s || !in) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (in_size < s->frame_bytes) return stream_fail(s, NV12EQ_ERR_SHORT_BUFFER, "stream push: buffer smaller than the frame");
    DeviceGuard guard(s->ctx->device);
    const size_t depth = s->lanes.size();
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->count == depth) {
        switch (s->cfg.full_policy) {
            case NV12EQ_FULL_DROP_NEWEST:
                s->st.dropped_backpressure++;
                if (out_seq) *out_seq = s->next_seq;
                s->next_seq++;  // the dropped frame keeps its number: the consumer sees a gap
                return NV12EQ_ERR_DROPPED;
            case NV12EQ_FULL_DROP_OLDEST: {
                Lane& old = s->lanes[s->head];
                lk.unlock();
                cudaEventSynchronize(old.done);  // its GPU work has to leave the lane before the lane is reused
                lk.lock();
                if (s->count == depth) {         // the consumer may have popped it meanwhile
                    s->head = (s->head + 1) % depth;
                    s->count--;
                    s->st.dropped_backpressure++;
                }
                break;
            }
            default:
                s->cv.wait(lk, [&] { return s->count < depth; });
        }
    }
    const size_t tail = (s->head + s->count) % depth;
    Lane& L = s->lanes[tail];
    lk.unlock();
    // the tail lane is free (not in [head, head+count)) and only the producer touches free lanes.  Luma goes to the
    // pinned upload buffer; passthrough chroma goes straight to the staged output frame and never crosses PCIe.
    const int H = s->cfg.height;
    stream_copy_rows(s, reinterpret_cast<uint8_t*>(L.h_in.p), in, 0, H);
    if (s->cfg.uv_mode == UV_COPY) stream_copy_rows(s, reinterpret_cast<uint8_t*>(L.h_out.p), in, H, H + H / 2);
    const auto now = std::chrono::steady_clock::now();
    int rc = stream_launch(s, L);
    if (rc) return rc;
    lk.lock();
    const uint64_t q = s->next_seq++;
    s->seq[tail] = q;
    s->t_push[tail] = now;
    s->count++;
    s->st.pushed++;
    s->st.in_flight = s->count;
    s->st.max_in_flight = std::max<uint64_t>(s->st.max_in_flight, s->count);
    if (out_seq) *out_seq = q;
    lk.unlock();
    s->cv.notify_all();
    return NV12EQ_OK;
}

int nv12eq_stream_pop(nv12eq_stream* s, uint8_t* out, size_t out_size, uint64_t* out_seq, int block) {
    if (!s || !out) return NV12EQ_ERR_INVALID_ARGUMENT;
    if (out_size < s->frame_bytes) return stream_fail(s, NV12EQ_ERR_SHORT_BUFFER, "stream pop: buffer smaller than the frame");
    DeviceGuard guard(s->ctx->device);
    const size_t depth = s->lanes.size();
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->count == 0) {
        if (!block) return NV12EQ_ERR_EMPTY;
        s->cv.wait(lk, [&] { return s->count > 0; });
    }
    const size_t h = s->head;
    Lane& L = s->lanes[h];
    const uint64_t q = s->seq[h];
    lk.unlock();
    if (block) {
        cudaError_t e = cudaEventSynchronize(L.done);
        if (e != cudaSuccess) return stream_fail(s, NV12EQ_ERR_CUDA, cudaGetErrorString(e));
    } else {
        cudaError_t e = cudaEventQuery(L.done);
        if (e == cudaErrorNotReady) { cudaGetLastError(); return NV12EQ_ERR_EMPTY; }
        if (e != cudaSuccess) return stream_fail(s, NV12EQ_ERR_CUDA, cudaGetErrorString(e));
    }
    lk.lock();
    if (s->head != h || s->count == 0 || s->seq[h] != q) {
        // DROP_OLDEST discarded this frame while we were waiting for it: try again with the new head
        lk.unlock();
        return nv12eq_stream_pop(s, out, out_size, out_seq, block);
    }
    if (*L.h_status != 0) {
        // a kernel of this lane gave up on a dependency wait (never expected): the frame is not valid.  Reset the lane's
        // workspace and deliver the failure in the frame's place -- the consumer drops it like any failed frame.
        *L.h_status = 0;
        ws_reset(L.ws, L.stream);
        s->head = (s->head + 1) % depth;
        s->count--;
        s->st.in_flight = s->count;
        if (out_seq) *out_seq = q;
        lk.unlock();
        s->cv.notify_all();
        return stream_fail(s, NV12EQ_ERR_CUDA, "kernel dependency wait timed out; the frame is dropped and the lane has been reset");
    }
    // copy out under the lock: a DROP_OLDEST producer must not recycle the lane while it is being read
    {
        const int H = s->cfg.height;
        stream_copy_rows(s, out, reinterpret_cast<const uint8_t*>(L.h_out.p), 0, s->cfg.uv_mode == UV_SKIP ? H : H + H / 2);
    }
    const uint64_t us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - s->t_push[h]).count();
    s->head = (s->head + 1) % depth;
    s->count--;
    s->st.delivered++;
    s->st.in_flight = s->count;
    s->st.latency_us_sum += us;
    s->st.latency_us_max = std::max(s->st.latency_us_max, us);
    if (out_seq) *out_seq = q;
    lk.unlock();
    s->cv.notify_all();
    return NV12EQ_OK;
}

int nv12eq_stream_get_stats(nv12eq_stream* s, nv12eq_stream_stats* out) {
    if (!s || !out) return NV12EQ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->st;
    return NV12EQ_OK;
}

void nv12eq_stream_close(nv12eq_stream* s) {
    if (!s) return;
    DeviceGuard guard(s->ctx->device);
    for (auto& L : s->lanes) {
        if (L.stream) cudaStreamSynchronize(L.stream);
        dev_release(L.d_in); dev_release(L.d_out); host_release(L.h_in); host_release(L.h_out);
        ws_release(L.ws);
        if (L.h_status) cudaFreeHost(L.h_status);
        if (L.done) cudaEventDestroy(L.done);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    delete s;
}

// ---- synthetic inputs -----------------------------------------------------------------------------------
int nv12eq_synth_nv12_device(nv12eq_ctx* ctx, uint8_t* d_out, int n_frames, size_t frame_pitch, int width, int height, int stride,
                             uint32_t seed, uint32_t first_frame, void* cuda_stream) {
    int rc = check_geometry(ctx, width, height, stride, n_frames, frame_pitch, 0);
    if (rc) return rc;
    if (!d_out) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "null device pointer");
    if (n_frames == 0) return NV12EQ_OK;
    if (n_frames > 65535) return fail(ctx, NV12EQ_ERR_TOO_LARGE, "at most 65535 frames per call");
    DeviceGuard guard(ctx->device);
    dim3 grid(ctx->sm_count * 4, n_frames);
    synth_nv12_kernel<<<grid, kSynthThreads, 0, pick_stream(ctx, cuda_stream)>>>(d_out, frame_pitch, width, height, stride, seed, first_frame);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}

int nv12eq_synth_bgr_device(nv12eq_ctx* ctx, uint8_t* d_out, int n_frames, size_t frame_pitch, int width, int height, int stride,
                            uint32_t first_frame, void* cuda_stream) {
    int rc = check_color(ctx, width, height, stride, 0);
    if (rc) return rc;
    if (!d_out || n_frames < 0 || n_frames > 65535) return fail(ctx, NV12EQ_ERR_INVALID_ARGUMENT, "bad arguments");
    if (n_frames == 0) return NV12EQ_OK;
    DeviceGuard guard(ctx->device);
    dim3 grid(ctx->sm_count * 4, n_frames);
    synth_bgr_kernel<<<grid, kSynthThreads, 0, pick_stream(ctx, cuda_stream)>>>(d_out, frame_pitch, width, height, stride, first_frame);
    ctx->ctr.kernel_launches++;
    CK(ctx, cudaGetLastError());
    return NV12EQ_OK;
}

}  // extern "C"
