/*
 * nv12eq.h -- C-ABI of libnv12eq.so: B200-native (sm_100a CUDA) histogram equalization and CLAHE on the
 * Y plane of NV12 frames, UV passed through.  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point replaces one piece of the reference's per-frame hot path (citations are into the
 * reference tree kimkimhun3/OpenCV-OpenCL):
 *
 *   nv12eq_create / nv12eq_destroy      <- shared + per-worker device context set-up,
 *                                          OpenCLequalHist.cpp:106-161 (SharedOpenCLContext/WorkerOpenCLContext :63-81),
 *                                          buffers cached by size like allocate_worker_opencl_buffers :175-192
 *   nv12eq_equalize_hist                <- the frame body of the worker: UV memcpy + cv::equalizeHist on views,
 *                                          nextimprovement.cpp:128-170 (also OpenCVequalHist.cpp:127-163,
 *                                          ColoropenCVCwqualHist.cpp:146-165) and the device launch protocol it
 *                                          replaces, OpenCLequalHist.cpp:349-365 + NV12 rebuild :388-390
 *   nv12eq_clahe                        <- cv::createCLAHE(clip, Size(t,t)) + clahe->apply on the Y view + NV12
 *                                          rebuild, clahevideo.cpp:178-201,:497 (CLAHECompare.cpp:144-154)
 *   nv12eq_*_batch / submit / wait      <- N worker threads popping one queue, OpenCVequalHist.cpp:71-98,397-402
 *   nv12eq_*_device                     <- the kernel boundary equalizeHist_accel(in, ref, out, rows, cols),
 *                                          accel.cpp:36-61 (device buffers in, device buffers out)
 *   nv12eq_color_equalize / _clahe      <- cvtColor(BGR2YUV) -> split -> equalizeHist/CLAHE(Y) -> merge ->
 *                                          cvtColor(YUV2BGR), singlecolor.cpp:39-66, clahe1frame.cpp:83-102
 *   nv12eq_*_meta                       <- the same frame body for buffers whose planes carry GstVideoMeta offsets/strides
 *                                          (the reference reads GstVideoInfo at nextimprovement.cpp:128-129 but assumes
 *                                          packed planes)
 *   nv12eq_clahe16* / nv12eq_p010_clahe <- cv::CLAHE::apply on CV_16UC1 (65536 bins), the 16-bit path of clahevideo.cpp:195's
 *                                          operator for P010 decoders
 *   nv12eq_nv12_to_bgr / bgr_to_nv12    <- cvtColor(COLOR_YUV2BGR_NV12) for display / the I420 arithmetic with interleaved chroma
 *   nv12eq_bgr_to_i420 / _device        <- cvtColor(bgr, COLOR_BGR2YUV_I420) in front of the Y-plane operator,
 *                                          1frameMeasure.cpp:32-35
 *   nv12eq_get_counters                 <- the Counters struct + status tick, OpenCLequalHist.cpp:45-61,439-508
 *   nv12eq_stream_*                     <- the frame queue between capture and encoder: GAsyncQueue + worker threads
 *                                          (OpenCVequalHist.cpp:71-98,102-196,397-402) with the leaky queues around them
 *                                          (leaky=downstream max-size-buffers=8 / drop=true, :296-297,:312), plus the
 *                                          in-order output the reference's unpublished IMP/improvement binaries added
 *                                          ("frame-output-ordering", "Max reorder", "Dropped: late/backpressure")
 *
 * Conventions:
 *   - every function returns an nv12eq_status (0 = ok); nothing throws, aborts or exits.  A failed frame is
 *     the caller's to drop, as the reference does (processing_errors / opencl_errors, OpenCLequalHist.cpp:339-344).
 *   - the caller owns both frame buffers for the duration of the call; the library owns only device memory and
 *     pinned staging memory.  Output must not partially overlap input; in == out (in place) is allowed.
 *   - one context per calling thread (thread-compatible, like one WorkerOpenCLContext per worker); contexts on
 *     the same device share the device's primary CUDA context.
 *   - NV12 layout: `height` rows of Y then `height/2` rows of interleaved UV, every row `stride` bytes apart,
 *     `stride >= width`; a frame occupies stride*(height + height/2) bytes.  Bytes between width and stride are
 *     never read or written.  The reference assumes stride == width (OpenCVequalHist.cpp:140).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     NV12EQ_ERR_NO_DEVICE / NV12EQ_ERR_CUDA.
 */
#ifndef NV12EQ_H_
#define NV12EQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define NV12EQ_VERSION_MAJOR 0
#define NV12EQ_VERSION_MINOR 2   /* 2: NV12 <-> BGR adapters, CLAHE band stages (spatial split) */

typedef struct nv12eq_ctx nv12eq_ctx;

typedef enum nv12eq_status {
    NV12EQ_OK = 0,
    NV12EQ_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive size, stride < width, bad enum, bad tile grid */
    NV12EQ_ERR_SHORT_BUFFER = 2,     /* buffer smaller than stride*(h + h/2): reference rejects these too
                                        (OpenCVequalHist.cpp:132-137) */
    NV12EQ_ERR_CUDA = 3,             /* a CUDA call failed; see nv12eq_last_error_string */
    NV12EQ_ERR_NO_DEVICE = 4,        /* no CUDA device / wrong architecture (needs sm_100) */
    NV12EQ_ERR_OUT_OF_MEMORY = 5,
    NV12EQ_ERR_BAD_SLOT = 6,         /* slot index out of range, or slot busy / not submitted */
    NV12EQ_ERR_TOO_LARGE = 7,        /* frame exceeds the max_width/max_height given to nv12eq_create, or 2^31 pixels */
    NV12EQ_ERR_DROPPED = 8,          /* stream: the frame was dropped by the back-pressure policy (counted, not an error
                                        for the pipeline: the reference drops too, OpenCVequalHist.cpp:296,312) */
    NV12EQ_ERR_EMPTY = 9             /* stream: nothing to pop (or the oldest frame is not finished and block == 0) */
} nv12eq_status;

/* What to put in the output chroma plane. */
typedef enum nv12eq_uv_mode {
    NV12EQ_UV_COPY = 0,    /* passthrough (nextimprovement.cpp:160, ColoropenCVCwqualHist.cpp:165) -- the default */
    NV12EQ_UV_GRAY128 = 1, /* neutral grey (OpenCVequalHist.cpp:162, clahevideo.cpp:201) */
    NV12EQ_UV_SKIP = 2     /* leave the output chroma bytes untouched (Y-only callers) */
} nv12eq_uv_mode;

typedef enum nv12eq_color_mode {
    NV12EQ_COLOR_YUV = 0,  /* COLOR_BGR2YUV / COLOR_YUV2BGR, what the reference uses (singlecolor.cpp:39,66) */
    NV12EQ_COLOR_YCRCB = 1 /* COLOR_BGR2YCrCb / COLOR_YCrCb2BGR (BASELINE.json config 5 wording) */
} nv12eq_color_mode;

/* Cumulative per-context counters (the reference's Counters struct, OpenCLequalHist.cpp:45-61). */
typedef struct nv12eq_counters {
    uint64_t frames;        /* frames processed successfully */
    uint64_t bytes_in;      /* host->device bytes moved by host-buffer entry points */
    uint64_t bytes_out;     /* device->host bytes */
    uint64_t errors;        /* calls that returned non-zero */
    uint64_t kernel_launches; /* CUDA kernels launched by this context */
    uint64_t busy_us;       /* wall time spent inside synchronous entry points */
} nv12eq_counters;

/* ---- library / context -------------------------------------------------------------------------------- */
int nv12eq_version(void);                       /* major*100 + minor */
const char* nv12eq_status_string(int status);
/* Last error text of this context (or of the calling thread's last failed nv12eq_create when ctx == NULL). */
const char* nv12eq_last_error_string(const nv12eq_ctx* ctx);

/* device: CUDA ordinal.  max_width/max_height: largest frame this context will see (sizes staging memory).
 * slots: number of host batches that may be in flight through nv12eq_submit_* (>= 1; 2 = double buffering).
 * Device and pinned memory are allocated lazily and cached by size. */
int nv12eq_create(int device, int max_width, int max_height, int slots, nv12eq_ctx** out_ctx);
void nv12eq_destroy(nv12eq_ctx* ctx);
int nv12eq_get_counters(const nv12eq_ctx* ctx, nv12eq_counters* out);
/* Tuning knobs (0 keeps the built-in default): chunks per frame, frame lag of the fused kernel, CTAs per SM,
 * fused (1) vs two-kernel (2) schedule. */
int nv12eq_set_tuning(nv12eq_ctx* ctx, int chunks_per_frame, int lag_frames, int ctas_per_sm, int schedule);

/* Pinned host memory helpers: buffers from here skip the staging copy inside the host entry points. */
int nv12eq_host_alloc(size_t bytes, void** out_ptr);
int nv12eq_host_free(void* ptr);

/* ---- host frame in / frame out (synchronous) ---------------------------------------------------------- */
int nv12eq_equalize_hist(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size,
                         int width, int height, int stride, int uv_mode);
/* clip_limit <= 0 disables clipping (as in OpenCV).  tiles_x/tiles_y >= 1 (reference: --tile, default 8). */
int nv12eq_clahe(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width,
                 int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode);

/* ---- host frames described by per-plane offsets and strides (GstVideoMeta: offset[2], stride[2]) ------------------------
 * the reference assumes stride == width and chroma at width*height (OpenCVequalHist.cpp:140);
 * real decoders and cameras pad rows and planes.  Plane 0 is Y (height rows), plane 1 is interleaved UV (height/2 rows);
 * the input and the output buffer may use different layouts.  Only the `width` payload bytes of every row are read or
 * written.  A NULL layout means the packed default {0, stride*height} / {width, width}. */
typedef struct nv12eq_layout {
    size_t offset[2];   /* byte offset of plane 0 (Y) and plane 1 (UV) from the buffer start */
    int stride[2];      /* bytes between rows of each plane, >= width */
} nv12eq_layout;
int nv12eq_equalize_hist_meta(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, const nv12eq_layout* in_layout, uint8_t* out,
                              size_t out_size, const nv12eq_layout* out_layout, int width, int height, int uv_mode);
int nv12eq_clahe_meta(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, const nv12eq_layout* in_layout, uint8_t* out,
                      size_t out_size, const nv12eq_layout* out_layout, int width, int height, double clip_limit, int tiles_x,
                      int tiles_y, int uv_mode);

/* ---- host batches: n_frames frames, frame k at in + k*frame_pitch (frame_pitch >= stride*(h+h/2)) ------ */
/* Synchronous; internally pipelined (pinned double buffering, H2D / kernels / D2H on separate streams). */
int nv12eq_equalize_hist_batch(nv12eq_ctx* ctx, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch,
                               int width, int height, int stride, int uv_mode);
int nv12eq_clahe_batch(nv12eq_ctx* ctx, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch,
                       int width, int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode);
/* Asynchronous: returns once the work is queued on slot `slot`; buffers stay owned by the caller but must not be
 * touched until nv12eq_wait(ctx, slot) returns.  nv12eq_query returns NV12EQ_OK when done, NV12EQ_ERR_BAD_SLOT
 * while still running. */
int nv12eq_submit_equalize_hist(nv12eq_ctx* ctx, int slot, const uint8_t* in, uint8_t* out, int n_frames,
                                size_t frame_pitch, int width, int height, int stride, int uv_mode);
int nv12eq_submit_clahe(nv12eq_ctx* ctx, int slot, const uint8_t* in, uint8_t* out, int n_frames, size_t frame_pitch,
                        int width, int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode);
int nv12eq_wait(nv12eq_ctx* ctx, int slot);
int nv12eq_query(nv12eq_ctx* ctx, int slot);

/* ---- device-resident forms (d_* are device pointers; asynchronous on `cuda_stream`) -------------------- */
/* cuda_stream is a cudaStream_t passed as void*.  NULL selects the context's own (non-blocking) stream, which
 * nv12eq_sync waits for; to run on the default stream pass cudaStreamLegacy or cudaStreamPerThread explicitly.
 * The device forms of one context share a workspace: calls are ordered on their stream, and when consecutive calls use
 * different streams the later stream is made to wait for the earlier one (use one context per stream for overlap). */
int nv12eq_equalize_hist_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch,
                                int width, int height, int stride, int uv_mode, void* cuda_stream);
int nv12eq_clahe_device(nv12eq_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_frames, size_t frame_pitch,
                        int width, int height, int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode,
                        void* cuda_stream);
int nv12eq_sync(nv12eq_ctx* ctx);

/* Stage-level device forms for the spatially split single-frame mode: a rank histograms
 * its band of rows, the 256-bin histograms are summed across ranks by the caller (ncclAllReduce), and every rank
 * applies the LUT of the summed histogram to its band.
 *   nv12eq_hist_device:   d_hist[n_planes][256] (uint32) += histogram of plane k (height rows of `width` bytes)
 *   nv12eq_equalize_apply_device: out plane = LUT(d_hist[k], total_pixels)[in plane]; total_pixels is the pixel
 *                                 count of the WHOLE frame the summed histogram describes. */
int nv12eq_hist_device(nv12eq_ctx* ctx, const uint8_t* d_y, int n_planes, size_t plane_pitch, int width, int height,
                       int stride, uint32_t* d_hist, void* cuda_stream);
int nv12eq_equalize_apply_device(nv12eq_ctx* ctx, const uint8_t* d_y_in, uint8_t* d_y_out, int n_planes,
                                 size_t plane_pitch, int width, int height, int stride, const uint32_t* d_hist,
                                 int64_t total_pixels, void* cuda_stream);

/* Spatial split of ONE frame for CLAHE (optional): rank r holds tile rows [first_tile_row, first_tile_row +
 * band_tiles_y) of the Y plane (the tile grid must divide the frame).  band_luts writes the band's band_tiles_y * tiles_x tile LUTs
 * (256 bytes each, row-major) to d_luts.  The caller then exchanges one tile row of LUTs with each neighbour (NCCL send/recv of
 * tiles_x * 256 bytes: sharding.SpatialSplitClahe) into a grid of (band_tiles_y + 2) * tiles_x tables -- row 0 = tile row
 * first_tile_row - 1, row band_tiles_y + 1 = tile row first_tile_row + band_tiles_y; rows outside the frame are never read -- and
 * band_apply interpolates the band with the weights of the whole frame.  d_y_band / d_out_band point at the band's first row. */
int nv12eq_clahe_band_luts_device(nv12eq_ctx* ctx, const uint8_t* d_y_band, int width, int full_height, int stride, double clip_limit,
                                  int tiles_x, int tiles_y, int first_tile_row, int band_tiles_y, uint8_t* d_luts, void* cuda_stream);
int nv12eq_clahe_band_apply_device(nv12eq_ctx* ctx, const uint8_t* d_y_band, uint8_t* d_out_band, int width, int full_height, int stride,
                                   int tiles_x, int tiles_y, int first_tile_row, int band_tiles_y, const uint8_t* d_luts_halo,
                                   void* cuda_stream);

/* ---- colour path: packed 8-bit BGR in, BGR out (stride in bytes, >= 3*width) --------------------------- */
int nv12eq_color_equalize(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int width, int height, int stride,
                          int color_mode);
int nv12eq_color_clahe(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int width, int height, int stride,
                       int color_mode, double clip_limit, int tiles_x, int tiles_y);
/* Host batch: n_frames packed BGR frames, frame k at in + k*frame_pitch (frame_pitch >= stride*height); pipelined over the
 * context's slots like the NV12 batch forms. */
int nv12eq_color_equalize_batch(nv12eq_ctx* ctx, const uint8_t* bgr_in, uint8_t* bgr_out, int n_frames, size_t frame_pitch, int width,
                                int height, int stride, int color_mode);
int nv12eq_color_equalize_device(nv12eq_ctx* ctx, const uint8_t* d_bgr_in, uint8_t* d_bgr_out, int n_frames,
                                 size_t frame_pitch, int width, int height, int stride, int color_mode,
                                 void* cuda_stream);
int nv12eq_color_clahe_device(nv12eq_ctx* ctx, const uint8_t* d_bgr_in, uint8_t* d_bgr_out, int n_frames,
                              size_t frame_pitch, int width, int height, int stride, int color_mode, double clip_limit,
                              int tiles_x, int tiles_y, void* cuda_stream);

/* ---- BGR -> I420 adapter (1frameMeasure.cpp:32): packed 8-bit BGR in, planar Y[h][w] U[h/2][w/2] V[h/2][w/2] out ---- */
/* width and height must be even (OpenCV rejects odd sizes for this conversion).  out holds w*h*3/2 bytes per frame. */
int nv12eq_bgr_to_i420(nv12eq_ctx* ctx, const uint8_t* bgr, int width, int height, int stride, uint8_t* out, size_t out_size);
int nv12eq_bgr_to_i420_device(nv12eq_ctx* ctx, const uint8_t* d_bgr, uint8_t* d_out, int n_frames, size_t bgr_pitch,
                              size_t out_pitch, int width, int height, int stride, void* cuda_stream);

/* ---- NV12 <-> BGR adapters --------------------------------------------------------------
 * nv12_to_bgr = cvtColor(nv12, COLOR_YUV2BGR_NV12): the display-side inverse of the NV12 path (the reference's still-image tools
 * convert back with cvtColor, singlecolor.cpp:66 / clahe1frame.cpp:102; its pipelines hand NV12 to the encoder).
 * bgr_to_nv12 = the arithmetic of COLOR_BGR2YUV_I420 (1frameMeasure.cpp:32) with the chroma planes interleaved (U first): the
 * frame the NV12 operators above take.  width and height must be even.  `stride` is the NV12 row stride (chroma rows at
 * stride * height), `bgr_stride` the BGR row stride, both in bytes; pitches are bytes between frames. */
int nv12eq_nv12_to_bgr(nv12eq_ctx* ctx, const uint8_t* nv12, size_t nv12_size, int width, int height, int stride, uint8_t* bgr,
                       size_t bgr_size, int bgr_stride);
int nv12eq_bgr_to_nv12(nv12eq_ctx* ctx, const uint8_t* bgr, size_t bgr_size, int width, int height, int bgr_stride, uint8_t* nv12,
                       size_t nv12_size, int stride);
int nv12eq_nv12_to_bgr_device(nv12eq_ctx* ctx, const uint8_t* d_nv12, uint8_t* d_bgr, int n_frames, size_t nv12_pitch,
                              size_t bgr_pitch, int width, int height, int stride, int bgr_stride, void* cuda_stream);
int nv12eq_bgr_to_nv12_device(nv12eq_ctx* ctx, const uint8_t* d_bgr, uint8_t* d_nv12, int n_frames, size_t bgr_pitch,
                              size_t nv12_pitch, int width, int height, int bgr_stride, int stride, void* cuda_stream);

/* ---- 16-bit planes: CLAHE on CV_16UC1 (P010 / 16-bit, histSize 65536) -------------------
 * cv::CLAHE::apply accepts CV_16UC1 with 65536-bin tile histograms; the reference only calls it on 8-bit Y planes
 * (clahevideo.cpp:195), so this is a widening row.  Strides and pitches of the 16-bit forms are in ELEMENTS (uint16). */
int nv12eq_clahe16_device(nv12eq_ctx* ctx, const uint16_t* d_in, uint16_t* d_out, int n_planes, size_t plane_pitch, int width,
                          int height, int stride, double clip_limit, int tiles_x, int tiles_y, void* cuda_stream);
int nv12eq_clahe16(nv12eq_ctx* ctx, const uint16_t* in, uint16_t* out, int width, int height, int stride, double clip_limit,
                   int tiles_x, int tiles_y);
/* P010 frame (10-bit samples in the high bits of 16-bit words): `height` rows of Y then `height/2` rows of interleaved UV,
 * `stride` BYTES apart (>= 2*width).  Y goes through the 16-bit CLAHE; chroma per uv_mode (GRAY128 writes 0x8000). */
int nv12eq_p010_clahe(nv12eq_ctx* ctx, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int width, int height,
                      int stride, double clip_limit, int tiles_x, int tiles_y, int uv_mode);

/* ---- ordered, back-pressured frame stream -------------------------------- */
/* A stream is the reference's worker queue as one object: frames are pushed in capture order, run on `depth` slots
 * (own CUDA stream, pinned staging and device buffers each, so upload / kernel / download of consecutive frames
 * overlap), and are popped strictly in push order with their sequence number.  One producer thread may push while
 * one consumer thread pops; the context must not be used for other calls meanwhile. */
typedef struct nv12eq_stream nv12eq_stream;
typedef enum nv12eq_op { NV12EQ_OP_EQUALIZE = 0, NV12EQ_OP_CLAHE = 1 } nv12eq_op;
typedef enum nv12eq_full_policy {
    NV12EQ_FULL_BLOCK = 0,       /* push waits until the consumer pops (GAsyncQueue without a bound never drops) */
    NV12EQ_FULL_DROP_NEWEST = 1, /* push returns NV12EQ_ERR_DROPPED and the frame is not processed */
    NV12EQ_FULL_DROP_OLDEST = 2  /* the oldest undelivered frame is discarded: GStreamer leaky=downstream / drop=true,
                                    what the reference configures (OpenCVequalHist.cpp:296-297) */
} nv12eq_full_policy;
typedef struct nv12eq_stream_config {
    int op;                 /* nv12eq_op */
    int width, height, stride;
    int uv_mode;            /* nv12eq_uv_mode */
    double clip_limit;      /* CLAHE only */
    int tiles_x, tiles_y;   /* CLAHE only */
    int depth;              /* frames in flight, 1..64 (reference: max-size-buffers=8) */
    int full_policy;        /* nv12eq_full_policy */
} nv12eq_stream_config;
typedef struct nv12eq_stream_stats {
    uint64_t pushed;               /* frames accepted */
    uint64_t delivered;            /* frames popped */
    uint64_t dropped_backpressure; /* frames dropped by the full policy (newest or oldest) */
    uint64_t in_flight;            /* accepted, not yet popped or dropped */
    uint64_t max_in_flight;
    uint64_t latency_us_sum;       /* push -> pop wall time of delivered frames */
    uint64_t latency_us_max;
} nv12eq_stream_stats;
int nv12eq_stream_open(nv12eq_ctx* ctx, const nv12eq_stream_config* cfg, nv12eq_stream** out_stream);
/* Copies the frame (in_size >= stride*(h + h/2)) and queues it.  *out_seq (optional) = its sequence number, 0, 1, ... */
int nv12eq_stream_push(nv12eq_stream* s, const uint8_t* in, size_t in_size, uint64_t* out_seq);
/* Oldest undelivered frame -> out; *out_seq = its sequence number (gaps = dropped frames).  block != 0 waits for a
 * frame to be pushed and finished; block == 0 returns NV12EQ_ERR_EMPTY instead of waiting. */
int nv12eq_stream_pop(nv12eq_stream* s, uint8_t* out, size_t out_size, uint64_t* out_seq, int block);
int nv12eq_stream_get_stats(nv12eq_stream* s, nv12eq_stream_stats* out);
void nv12eq_stream_close(nv12eq_stream* s);

/* ---- synthetic inputs on the device (SURVEY.md Appendix B generator; bench/test utility) --------------- */
/* Frame k of the batch is synth_nv12(width, height, seed, first_frame + k). */
int nv12eq_synth_nv12_device(nv12eq_ctx* ctx, uint8_t* d_out, int n_frames, size_t frame_pitch, int width, int height,
                             int stride, uint32_t seed, uint32_t first_frame, void* cuda_stream);
/* BGR frame k = planes synthesised with seeds 3026/4026/5026 and frame first_frame + k. */
int nv12eq_synth_bgr_device(nv12eq_ctx* ctx, uint8_t* d_out, int n_frames, size_t frame_pitch, int width, int height,
                            int stride, uint32_t first_frame, void* cuda_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NV12EQ_H_ */
