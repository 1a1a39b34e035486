/*
 * nv12eq_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path is the CUDA
 * library in opencv-opencl_b200/csrc and it has no CPU fallback.
 *
 * What it restates
 * ----------------
 * The reference (kimkimhun3/OpenCV-OpenCL) does no arithmetic of its own on this path: each
 * program wraps the first `height` rows of an NV12 buffer as an 8-bit single channel image and
 * calls OpenCV, then copies (or greys) the chroma plane:
 *   - NV12 adapter            nextimprovement.cpp:128-165, OpenCVequalHist.cpp:127-142
 *   - cv::equalizeHist        nextimprovement.cpp:168, OpenCVequalHist.cpp:145, singlecolor.cpp:55
 *   - cv::CLAHE::apply        clahevideo.cpp:184-195, CLAHECompare.cpp:144-150, clahe1frame.cpp:88-93
 *   - UV passthrough / 128    nextimprovement.cpp:160 / OpenCVequalHist.cpp:162, clahevideo.cpp:201
 *   - colour path             singlecolor.cpp:39-66, clahe1frame.cpp:83-102
 * OpenCV itself is a third-party dependency that is NOT vendored under /root/reference (the
 * shipped binaries link libopencv_imgproc.so.4.4; compile.sh:10 uses `pkg-config opencv4`, no
 * version pin).  The algorithm restated here is OpenCV's published one (imgproc histogram.cpp
 * equalizeHist, imgproc clahe.cpp, imgproc color_yuv 8-bit fixed point), as written down in
 * SURVEY.md Appendix A.  Parity is PINNED, not assumed: tests/golden/make_golden.py runs the very
 * functions the reference calls (cv2 4.13.0, same C++ code behind Python bindings) on seeded
 * inputs and commits digests + small raw fixtures; tests/test_oracle.py checks this file against
 * them (and against live cv2 whenever cv2 is importable).
 *
 * Floating point: every f32 operation below must be a separately rounded IEEE binary32 operation
 * (OpenCV's x86 baseline build has no FMA).  Build with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

enum { UV_COPY = 0, UV_GRAY128 = 1, UV_SKIP = 2 };
enum { COLOR_YUV = 0, COLOR_YCRCB = 1 };

/* cvRound on x86 = cvtss2si under the default rounding mode = round half to even. */
static inline int rne(float v) { return (int)lrintf(v); }
static inline uint8_t sat8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* ------------------------------------------------------------------------------------------
 * Appendix B: deterministic synthetic NV12 frames (all arithmetic mod 2^32).
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t fmix32(uint32_t k) {
    k ^= k >> 16; k *= 0x85EBCA6Bu; k ^= k >> 13; k *= 0xC2B2AE35u; k ^= k >> 16;
    return k;
}

ORACLE_API void oracle_synth_y(uint8_t* y, int stride, int W, int H, uint32_t seed, uint32_t frame) {
    int bw = W / 16 > 1 ? W / 16 : 1;
    int bh = H / 9 > 1 ? H / 9 : 1;
    for (int r = 0; r < H; ++r) {
        for (int c = 0; c < W; ++c) {
            uint32_t idx = (uint32_t)r * (uint32_t)W + (uint32_t)c;
            uint32_t k = fmix32(idx * 0x9E3779B1u + seed * 0x85EBCA77u + frame * 0xC2B2AE3Du);
            int base = 48 + (c * 128) / W + (r * 48) / H;
            int noise = (int)(k & 63u) - 32;
            if (((c / bw) + (r / bh)) % 5 == 0) { base = 200; noise = (int)(k & 3u); }
            y[(size_t)r * stride + c] = sat8(base + noise);
        }
    }
}

ORACLE_API void oracle_synth_uv(uint8_t* uv, int stride, int W, int H, uint32_t seed, uint32_t frame) {
    int rows = H / 2;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < W; ++c) {
            uint32_t j = (uint32_t)r * (uint32_t)W + (uint32_t)c;
            uint32_t k = fmix32(j * 0x9E3779B1u + seed + 0x01234567u + frame * 0xC2B2AE3Du);
            uv[(size_t)r * stride + c] = (uint8_t)(128 + (int)(k & 31u) - 16);
        }
}

/* NV12 frame = H rows of Y then H/2 rows of interleaved UV, all with the same row stride. */
ORACLE_API void oracle_synth_nv12(uint8_t* nv12, int stride, int W, int H, uint32_t seed, uint32_t frame) {
    oracle_synth_y(nv12, stride, W, H, seed, frame);
    oracle_synth_uv(nv12 + (size_t)stride * H, stride, W, H, seed, frame);
}

/* ------------------------------------------------------------------------------------------
 * A.1 equalizeHist (8UC1).  Follows cv::equalizeHist as called at nextimprovement.cpp:168.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_hist256(const uint8_t* src, int stride, int W, int H, int32_t hist[256]) {
    memset(hist, 0, 256 * sizeof(int32_t));
    for (int r = 0; r < H; ++r) {
        const uint8_t* p = src + (size_t)r * stride;
        for (int c = 0; c < W; ++c) hist[p[c]]++;
    }
}

/* Returns 1 when the image is constant (lut filled with i0), else 0. */
ORACLE_API int oracle_equalize_lut(const int32_t hist[256], int64_t total, uint8_t lut[256]) {
    int i0 = 0;
    while (i0 < 256 && hist[i0] == 0) ++i0;
    if (i0 == 256) { memset(lut, 0, 256); return 1; }          /* empty image */
    if ((int64_t)hist[i0] == total) { memset(lut, i0, 256); return 1; }
    float scale = 255.0f / (float)(total - hist[i0]);
    int sum = 0;
    for (int i = 0; i <= i0; ++i) lut[i] = 0;
    for (int i = i0 + 1; i < 256; ++i) {
        sum += hist[i];
        lut[i] = sat8(rne((float)sum * scale));
    }
    return 0;
}

ORACLE_API void oracle_equalize_hist(const uint8_t* src, int sstride, uint8_t* dst, int dstride, int W, int H) {
    int32_t hist[256];
    uint8_t lut[256];
    if (W <= 0 || H <= 0) return;
    oracle_hist256(src, sstride, W, H, hist);
    oracle_equalize_lut(hist, (int64_t)W * H, lut);
    for (int r = 0; r < H; ++r) {
        const uint8_t* s = src + (size_t)r * sstride;
        uint8_t* d = dst + (size_t)r * dstride;
        for (int c = 0; c < W; ++c) d[c] = lut[s[c]];
    }
}

/* ------------------------------------------------------------------------------------------
 * A.2 CLAHE (8UC1).  Follows cv::createCLAHE(clip, Size(tx,ty))->apply as called at
 * clahevideo.cpp:184-195.
 * ---------------------------------------------------------------------------------------- */
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* Geometry shared with the tests: padded size, tile size, integer clip limit. */
ORACLE_API void oracle_clahe_geometry(int W, int H, double clip, int tx, int ty,
                                      int* extW, int* extH, int* tw, int* th, int* clipLimit) {
    int eW = W, eH = H;
    if (W % tx != 0 || H % ty != 0) {
        eW = W + (tx - (W % tx));
        eH = H + (ty - (H % ty));
    }
    *extW = eW; *extH = eH;
    *tw = eW / tx; *th = eH / ty;
    int area = (*tw) * (*th);
    int cl = 0;
    if (clip > 0.0) {
        cl = (int)(clip * area / 256.0);
        if (cl < 1) cl = 1;
    }
    *clipLimit = cl;
}

/* luts: tx*ty tables of 256 bytes, row-major over tiles. */
ORACLE_API void oracle_clahe_tile_luts(const uint8_t* src, int stride, int W, int H, double clip,
                                       int tx, int ty, uint8_t* luts) {
    int eW, eH, tw, th, clipLimit;
    oracle_clahe_geometry(W, H, clip, tx, ty, &eW, &eH, &tw, &th, &clipLimit);
    float lutScale = 255.0f / (float)(tw * th);
    for (int tyi = 0; tyi < ty; ++tyi) {
        for (int txi = 0; txi < tx; ++txi) {
            int h[256];
            memset(h, 0, sizeof h);
            for (int r = tyi * th; r < (tyi + 1) * th; ++r) {
                const uint8_t* row = src + (size_t)reflect101(r, H) * stride;
                for (int c = txi * tw; c < (txi + 1) * tw; ++c) h[row[reflect101(c, W)]]++;
            }
            if (clipLimit > 0) {
                int clipped = 0;
                for (int i = 0; i < 256; ++i)
                    if (h[i] > clipLimit) { clipped += h[i] - clipLimit; h[i] = clipLimit; }
                int redistBatch = clipped / 256;
                int residual = clipped - redistBatch * 256;
                for (int i = 0; i < 256; ++i) h[i] += redistBatch;
                if (residual != 0) {
                    int step = 256 / residual; if (step < 1) step = 1;
                    for (int i = 0; i < 256 && residual > 0; i += step, residual--) h[i]++;
                }
            }
            uint8_t* lut = luts + (size_t)(tyi * tx + txi) * 256;
            int sum = 0;
            for (int i = 0; i < 256; ++i) {
                sum += h[i];
                lut[i] = sat8(rne((float)sum * lutScale));
            }
        }
    }
}

ORACLE_API int oracle_clahe(const uint8_t* src, int sstride, uint8_t* dst, int dstride, int W, int H,
                            double clip, int tx, int ty) {
    if (W <= 0 || H <= 0 || tx < 1 || ty < 1) return -1;
    int eW, eH, tw, th, clipLimit;
    oracle_clahe_geometry(W, H, clip, tx, ty, &eW, &eH, &tw, &th, &clipLimit);
    uint8_t* luts = (uint8_t*)malloc((size_t)tx * ty * 256);
    int* ind1 = (int*)malloc(sizeof(int) * W * 2);
    float* xa = (float*)malloc(sizeof(float) * W * 2);
    if (!luts || !ind1 || !xa) { free(luts); free(ind1); free(xa); return -2; }
    int* ind2 = ind1 + W;
    float* xa1 = xa + W;
    oracle_clahe_tile_luts(src, sstride, W, H, clip, tx, ty, luts);

    float inv_tw = 1.0f / (float)tw;
    float inv_th = 1.0f / (float)th;
    for (int x = 0; x < W; ++x) {
        float txf = (float)x * inv_tw - 0.5f;
        int t1 = (int)floorf(txf);
        int t2 = t1 + 1;
        xa[x] = txf - (float)t1;
        xa1[x] = 1.0f - xa[x];
        if (t1 < 0) t1 = 0;
        if (t2 > tx - 1) t2 = tx - 1;
        ind1[x] = t1; ind2[x] = t2;
    }
    for (int y = 0; y < H; ++y) {
        float tyf = (float)y * inv_th - 0.5f;
        int t1 = (int)floorf(tyf);
        int t2 = t1 + 1;
        float ya = tyf - (float)t1;
        float ya1 = 1.0f - ya;
        if (t1 < 0) t1 = 0;
        if (t2 > ty - 1) t2 = ty - 1;
        const uint8_t* plane1 = luts + (size_t)t1 * tx * 256;
        const uint8_t* plane2 = luts + (size_t)t2 * tx * 256;
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < W; ++x) {
            int v = s[x];
            float top = (float)plane1[ind1[x] * 256 + v] * xa1[x] + (float)plane1[ind2[x] * 256 + v] * xa[x];
            float bot = (float)plane2[ind1[x] * 256 + v] * xa1[x] + (float)plane2[ind2[x] * 256 + v] * xa[x];
            float res = top * ya1 + bot * ya;
            d[x] = sat8(rne(res));
        }
    }
    free(luts); free(ind1); free(xa);
    return 0;
}


/* Interpolation stage of oracle_clahe for rows [y_first, y_first + rows) of a W x H frame whose tile grid divides it, from a LUT
 * grid with halo: `luts` holds tile rows first_tile_row - 1 .. (row-major, tx tables of 256 bytes per tile row); tile rows are
 * clamped to the frame exactly as above, so the halo rows outside the frame are never read.  src / dst point at row y_first.
 * This is the per-rank stage of the spatially split single-frame mode. */
ORACLE_API int oracle_clahe_interp_band(const uint8_t* src, int sstride, uint8_t* dst, int dstride, int W, int H, int tx, int ty,
                                        int y_first, int rows, const uint8_t* luts, int first_tile_row) {
    if (!src || !dst || !luts || W <= 0 || H <= 0 || tx < 1 || ty < 1 || W % tx || H % ty || y_first < 0 || y_first + rows > H) return -1;
    const int tw = W / tx, th = H / ty;
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    for (int y = y_first; y < y_first + rows; ++y) {
        float tyf = (float)y * inv_th - 0.5f;
        int t1 = (int)floorf(tyf);
        int t2 = t1 + 1;
        float ya = tyf - (float)t1;
        float ya1 = 1.0f - ya;
        if (t1 < 0) t1 = 0;
        if (t2 > ty - 1) t2 = ty - 1;
        const uint8_t* plane1 = luts + (size_t)(t1 - (first_tile_row - 1)) * tx * 256;
        const uint8_t* plane2 = luts + (size_t)(t2 - (first_tile_row - 1)) * tx * 256;
        const uint8_t* s = src + (size_t)(y - y_first) * sstride;
        uint8_t* d = dst + (size_t)(y - y_first) * dstride;
        for (int x = 0; x < W; ++x) {
            float txf = (float)x * inv_tw - 0.5f;
            int x1 = (int)floorf(txf);
            int x2 = x1 + 1;
            float xa = txf - (float)x1;
            float xa1 = 1.0f - xa;
            if (x1 < 0) x1 = 0;
            if (x2 > tx - 1) x2 = tx - 1;
            int v = s[x];
            float top = (float)plane1[x1 * 256 + v] * xa1 + (float)plane1[x2 * 256 + v] * xa;
            float bot = (float)plane2[x1 * 256 + v] * xa1 + (float)plane2[x2 * 256 + v] * xa;
            float res = top * ya1 + bot * ya;
            d[x] = sat8(rne(res));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CLAHE on CV_16UC1 (the P010 / 16-bit path OpenCV's CLAHE also accepts).
 * Same algorithm as A.2 with histSize = 65536, lutScale = 65535.f / tileArea, clipLimit = clip * tileArea / 65536
 * (>= 1 when clip > 0), 16-bit LUTs.  Verified bit-exact against cv2 4.13.0 (tests/golden/make_golden_ext.py).
 * Strides are in ELEMENTS (uint16), not bytes.  Returns 0, -1 bad arguments, -2 out of memory.
 * ---------------------------------------------------------------------------------------- */
static inline uint16_t sat16(int v) { return (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); }

ORACLE_API int oracle_clahe16(const uint16_t* src, int sstride, uint16_t* dst, int dstride, int W, int H,
                              double clip, int tx, int ty) {
    if (!src || !dst || W <= 0 || H <= 0 || tx < 1 || ty < 1) return -1;
    int eW = W, eH = H;
    if (W % tx != 0 || H % ty != 0) { eW = W + (tx - (W % tx)); eH = H + (ty - (H % ty)); }
    const int tw = eW / tx, th = eH / ty, area = tw * th;
    int clipLimit = 0;
    if (clip > 0.0) { clipLimit = (int)(clip * area / 65536.0); if (clipLimit < 1) clipLimit = 1; }
    const float lutScale = 65535.0f / (float)area;
    uint16_t* luts = (uint16_t*)malloc((size_t)tx * ty * 65536 * sizeof(uint16_t));
    int* h = (int*)malloc(sizeof(int) * 65536);
    int* ind1 = (int*)malloc(sizeof(int) * W * 2);
    float* xa = (float*)malloc(sizeof(float) * W * 2);
    if (!luts || !h || !ind1 || !xa) { free(luts); free(h); free(ind1); free(xa); return -2; }
    for (int tyi = 0; tyi < ty; ++tyi)
        for (int txi = 0; txi < tx; ++txi) {
            memset(h, 0, sizeof(int) * 65536);
            for (int r = tyi * th; r < (tyi + 1) * th; ++r) {
                const uint16_t* row = src + (size_t)reflect101(r, H) * sstride;
                for (int c = txi * tw; c < (txi + 1) * tw; ++c) h[row[reflect101(c, W)]]++;
            }
            if (clipLimit > 0) {
                long long clipped = 0;
                for (int i = 0; i < 65536; ++i)
                    if (h[i] > clipLimit) { clipped += h[i] - clipLimit; h[i] = clipLimit; }
                int redistBatch = (int)(clipped / 65536);
                int residual = (int)(clipped - (long long)redistBatch * 65536);
                for (int i = 0; i < 65536; ++i) h[i] += redistBatch;
                if (residual != 0) {
                    int step = 65536 / residual; if (step < 1) step = 1;
                    for (int i = 0; i < 65536 && residual > 0; i += step, residual--) h[i]++;
                }
            }
            uint16_t* lut = luts + (size_t)(tyi * tx + txi) * 65536;
            int sum = 0;
            for (int i = 0; i < 65536; ++i) { sum += h[i]; lut[i] = sat16(rne((float)sum * lutScale)); }
        }
    int* ind2 = ind1 + W;
    float* xa1 = xa + W;
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    for (int x = 0; x < W; ++x) {
        float txf = (float)x * inv_tw - 0.5f;
        int t1 = (int)floorf(txf), t2 = t1 + 1;
        xa[x] = txf - (float)t1;
        xa1[x] = 1.0f - xa[x];
        if (t1 < 0) t1 = 0;
        if (t2 > tx - 1) t2 = tx - 1;
        ind1[x] = t1; ind2[x] = t2;
    }
    for (int y = 0; y < H; ++y) {
        float tyf = (float)y * inv_th - 0.5f;
        int t1 = (int)floorf(tyf), t2 = t1 + 1;
        float ya = tyf - (float)t1, ya1 = 1.0f - ya;
        if (t1 < 0) t1 = 0;
        if (t2 > ty - 1) t2 = ty - 1;
        const uint16_t* plane1 = luts + (size_t)t1 * tx * 65536;
        const uint16_t* plane2 = luts + (size_t)t2 * tx * 65536;
        const uint16_t* s = src + (size_t)y * sstride;
        uint16_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < W; ++x) {
            int v = s[x];
            float top = (float)plane1[(size_t)ind1[x] * 65536 + v] * xa1[x] + (float)plane1[(size_t)ind2[x] * 65536 + v] * xa[x];
            float bot = (float)plane2[(size_t)ind1[x] * 65536 + v] * xa1[x] + (float)plane2[(size_t)ind2[x] * 65536 + v] * xa[x];
            float res = top * ya1 + bot * ya;
            d[x] = sat16(rne(res));
        }
    }
    free(luts); free(h); free(ind1); free(xa);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * NV12 frame-in / frame-out forms (the per-frame body of the reference worker,
 * nextimprovement.cpp:128-170 and clahevideo.cpp:158-201).
 * ---------------------------------------------------------------------------------------- */
static void uv_plane(const uint8_t* in, uint8_t* out, int W, int H, int stride, int uv_mode) {
    size_t off = (size_t)stride * H;
    int rows = H / 2;
    if (uv_mode == UV_SKIP) return;
    for (int r = 0; r < rows; ++r) {
        if (uv_mode == UV_COPY) {
            if (out != in) memcpy(out + off + (size_t)r * stride, in + off + (size_t)r * stride, (size_t)W);
        } else {
            memset(out + off + (size_t)r * stride, 128, (size_t)W);
        }
    }
}

ORACLE_API int oracle_nv12_equalize_hist(const uint8_t* in, uint8_t* out, int W, int H, int stride, int uv_mode) {
    if (!in || !out || W <= 0 || H <= 0 || stride < W) return -1;
    uv_plane(in, out, W, H, stride, uv_mode);
    oracle_equalize_hist(in, stride, out, stride, W, H);
    return 0;
}

ORACLE_API int oracle_nv12_clahe(const uint8_t* in, uint8_t* out, int W, int H, int stride,
                                 double clip, int tx, int ty, int uv_mode) {
    if (!in || !out || W <= 0 || H <= 0 || stride < W) return -1;
    uv_plane(in, out, W, H, stride, uv_mode);
    return oracle_clahe(in, stride, out, stride, W, H, clip, tx, ty);
}

/* Frame-parallel batch forms: one frame per OpenMP thread, mirroring the reference's
 * --workers N frame-level data parallelism (OpenCVequalHist.cpp:397-402). */
ORACLE_API int oracle_nv12_equalize_hist_batch(const uint8_t* in, uint8_t* out, int n, size_t pitch,
                                               int W, int H, int stride, int uv_mode, int threads) {
    int rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < n; ++i) {
        int r = oracle_nv12_equalize_hist(in + (size_t)i * pitch, out + (size_t)i * pitch, W, H, stride, uv_mode);
        if (r) rc = r;
    }
    return rc;
}

ORACLE_API int oracle_nv12_clahe_batch(const uint8_t* in, uint8_t* out, int n, size_t pitch, int W, int H,
                                       int stride, double clip, int tx, int ty, int uv_mode, int threads) {
    int rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < n; ++i) {
        int r = oracle_nv12_clahe(in + (size_t)i * pitch, out + (size_t)i * pitch, W, H, stride, clip, tx, ty, uv_mode);
        if (r) rc = r;
    }
    return rc;
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * A.3 8-bit colour conversions (Q14), and the colour path of singlecolor.cpp:39-66:
 * BGR -> YUV (or YCrCb), equalize channel 0, -> BGR.
 * ---------------------------------------------------------------------------------------- */
static inline int descale14(int v) { return (v + 8192) >> 14; }

static inline void bgr_to_ycc(int B, int G, int R, int mode, uint8_t* o0, uint8_t* o1, uint8_t* o2) {
    int Y = descale14(1868 * B + 9617 * G + 4899 * R);
    if (mode == COLOR_YUV) {
        int U = descale14((B - Y) * 8061 + (128 << 14));
        int V = descale14((R - Y) * 14369 + (128 << 14));
        *o0 = sat8(Y); *o1 = sat8(U); *o2 = sat8(V);
    } else {
        int Cr = descale14((R - Y) * 11682 + (128 << 14));
        int Cb = descale14((B - Y) * 9241 + (128 << 14));
        *o0 = sat8(Y); *o1 = sat8(Cr); *o2 = sat8(Cb);
    }
}

static inline void ycc_to_bgr(int c0, int c1, int c2, int mode, uint8_t* B, uint8_t* G, uint8_t* R) {
    if (mode == COLOR_YUV) {
        int U = c1 - 128, V = c2 - 128;
        *B = sat8(c0 + descale14(U * 33292));
        *G = sat8(c0 + descale14(U * -6472 + V * -9519));
        *R = sat8(c0 + descale14(V * 18678));
    } else {
        int Cr = c1 - 128, Cb = c2 - 128;
        *B = sat8(c0 + descale14(Cb * 29049));
        *G = sat8(c0 + descale14(Cb * -5636 + Cr * -11698));
        *R = sat8(c0 + descale14(Cr * 22987));
    }
}

ORACLE_API void oracle_bgr2ycc(const uint8_t* bgr, int sstride, uint8_t* ycc, int dstride, int W, int H, int mode) {
    for (int r = 0; r < H; ++r) {
        const uint8_t* s = bgr + (size_t)r * sstride;
        uint8_t* d = ycc + (size_t)r * dstride;
        for (int c = 0; c < W; ++c) bgr_to_ycc(s[3 * c], s[3 * c + 1], s[3 * c + 2], mode, d + 3 * c, d + 3 * c + 1, d + 3 * c + 2);
    }
}

ORACLE_API void oracle_ycc2bgr(const uint8_t* ycc, int sstride, uint8_t* bgr, int dstride, int W, int H, int mode) {
    for (int r = 0; r < H; ++r) {
        const uint8_t* s = ycc + (size_t)r * sstride;
        uint8_t* d = bgr + (size_t)r * dstride;
        for (int c = 0; c < W; ++c) ycc_to_bgr(s[3 * c], s[3 * c + 1], s[3 * c + 2], mode, d + 3 * c, d + 3 * c + 1, d + 3 * c + 2);
    }
}

/* use_clahe == 0: equalizeHist on channel 0 (singlecolor.cpp:55); else CLAHE (clahe1frame.cpp:93). */
ORACLE_API int oracle_color_equalize(const uint8_t* bgr_in, uint8_t* bgr_out, int W, int H, int stride, int mode,
                                     int use_clahe, double clip, int tx, int ty) {
    if (!bgr_in || !bgr_out || W <= 0 || H <= 0 || stride < 3 * W) return -1;
    size_t n = (size_t)W * H;
    uint8_t* ycc = (uint8_t*)malloc(n * 3);
    uint8_t* y = (uint8_t*)malloc(n * 2);
    if (!ycc || !y) { free(ycc); free(y); return -2; }
    uint8_t* y2 = y + n;
    oracle_bgr2ycc(bgr_in, stride, ycc, 3 * W, W, H, mode);
    for (size_t i = 0; i < n; ++i) y[i] = ycc[3 * i];
    int rc = 0;
    if (use_clahe) rc = oracle_clahe(y, W, y2, W, W, H, clip, tx, ty);
    else oracle_equalize_hist(y, W, y2, W, W, H);
    for (size_t i = 0; i < n; ++i) ycc[3 * i] = y2[i];
    oracle_ycc2bgr(ycc, 3 * W, bgr_out, stride, W, H, mode);
    free(ycc); free(y);
    return rc;
}

/* cv::cvtColor(bgr, COLOR_BGR2YUV_I420) as used by 1frameMeasure.cpp:32: Q20 limited-range BT.601,
 * chroma sampled at the top-left pixel of every 2x2 block, planar output Y[H][W], U[H/2][W/2], V[H/2][W/2].
 * W and H must be even (OpenCV rejects odd sizes).  Returns 0, or -1 on bad arguments. */
ORACLE_API int oracle_bgr2i420(const uint8_t* bgr, int stride, uint8_t* out, int W, int H) {
    if (!bgr || !out || W <= 0 || H <= 0 || (W & 1) || (H & 1) || stride < 3 * W) return -1;
    uint8_t* Y = out;
    uint8_t* U = out + (size_t)W * H;
    uint8_t* V = U + (size_t)(W / 2) * (H / 2);
    for (int r = 0; r < H; ++r) {
        const uint8_t* s = bgr + (size_t)r * stride;
        for (int c = 0; c < W; ++c) {
            const int B = s[3 * c], G = s[3 * c + 1], R = s[3 * c + 2];
            Y[(size_t)r * W + c] = (uint8_t)((269484 * R + 528482 * G + 102760 * B + (16 << 20) + (1 << 19)) >> 20);
            if (!(r & 1) && !(c & 1)) {
                const size_t k = (size_t)(r / 2) * (W / 2) + c / 2;
                U[k] = (uint8_t)((-155188 * R - 305135 * G + 460324 * B + (128 << 20) + (1 << 19)) >> 20);
                V[k] = (uint8_t)((460324 * R - 385875 * G - 74448 * B + (128 << 20) + (1 << 19)) >> 20);
            }
        }
    }
    return 0;
}

/* cv::cvtColor(nv12, COLOR_YUV2BGR_NV12): the display-side inverse of the NV12 path (the reference's
 * pipelines hand NV12 to the encoder, its still-image tools convert back with cvtColor, singlecolor.cpp:66 / clahe1frame.cpp:102).
 * OpenCV's YUV420sp2RGB: Q20 limited-range BT.601, one chroma pair per 2x2 block, saturating.  W and H even.
 * nv12: Y rows at `stride`, UV rows at nv12 + stride * H. */
ORACLE_API int oracle_nv12_to_bgr(const uint8_t* nv12, int stride, uint8_t* bgr, int bgr_stride, int W, int H) {
    if (!nv12 || !bgr || W <= 0 || H <= 0 || (W & 1) || (H & 1) || stride < W || bgr_stride < 3 * W) return -1;
    const uint8_t* uv = nv12 + (size_t)stride * H;
    for (int r = 0; r < H; ++r) {
        const uint8_t* yrow = nv12 + (size_t)r * stride;
        const uint8_t* uvrow = uv + (size_t)(r / 2) * stride;
        uint8_t* d = bgr + (size_t)r * bgr_stride;
        for (int c = 0; c < W; ++c) {
            const int u = (int)uvrow[c & ~1] - 128, v = (int)uvrow[(c & ~1) + 1] - 128;
            const int y = (yrow[c] > 16 ? (int)yrow[c] - 16 : 0) * 1220542;
            const int ruv = (1 << 19) + 1673527 * v;
            const int guv = (1 << 19) - 852492 * v - 409993 * u;
            const int buv = (1 << 19) + 2116026 * u;
            const int B = (y + buv) >> 20, G = (y + guv) >> 20, R = (y + ruv) >> 20;
            d[3 * c] = (uint8_t)(B < 0 ? 0 : B > 255 ? 255 : B);
            d[3 * c + 1] = (uint8_t)(G < 0 ? 0 : G > 255 ? 255 : G);
            d[3 * c + 2] = (uint8_t)(R < 0 ? 0 : R > 255 ? 255 : R);
        }
    }
    return 0;
}

/* BGR -> NV12: the arithmetic of COLOR_BGR2YUV_I420 above with the two chroma planes interleaved (U first), i.e. the frame the
 * NV12 operators take.  (OpenCV has no direct BGR -> NV12 code; the golden vectors are cv2's I420 output re-interleaved.) */
ORACLE_API int oracle_bgr_to_nv12(const uint8_t* bgr, int bgr_stride, uint8_t* nv12, int stride, int W, int H) {
    if (!bgr || !nv12 || W <= 0 || H <= 0 || (W & 1) || (H & 1) || stride < W || bgr_stride < 3 * W) return -1;
    uint8_t* uv = nv12 + (size_t)stride * H;
    for (int r = 0; r < H; ++r) {
        const uint8_t* s = bgr + (size_t)r * bgr_stride;
        for (int c = 0; c < W; ++c) {
            const int B = s[3 * c], G = s[3 * c + 1], R = s[3 * c + 2];
            nv12[(size_t)r * stride + c] = (uint8_t)((269484 * R + 528482 * G + 102760 * B + (16 << 20) + (1 << 19)) >> 20);
            if (!(r & 1) && !(c & 1)) {
                uint8_t* d = uv + (size_t)(r / 2) * stride + c;
                d[0] = (uint8_t)((-155188 * R - 305135 * G + 460324 * B + (128 << 20) + (1 << 19)) >> 20);
                d[1] = (uint8_t)((460324 * R - 385875 * G - 74448 * B + (128 << 20) + (1 << 19)) >> 20);
            }
        }
    }
    return 0;
}

/* Colour-path synthetic input (Appendix B): B,G,R planes = Y syntheses with seeds 3026/4026/5026. */
ORACLE_API void oracle_synth_bgr(uint8_t* bgr, int stride, int W, int H, uint32_t frame) {
    uint8_t* plane = (uint8_t*)malloc((size_t)W * H);
    if (!plane) return;
    const uint32_t seeds[3] = {3026u, 4026u, 5026u};
    for (int ch = 0; ch < 3; ++ch) {
        oracle_synth_y(plane, W, W, H, seeds[ch], frame);
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) bgr[(size_t)r * stride + 3 * c + ch] = plane[(size_t)r * W + c];
    }
    free(plane);
}
