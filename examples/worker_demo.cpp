// worker_demo.cpp -- the reference's per-frame worker body (nextimprovement.cpp:102-196 / clahevideo.cpp:105-283) with the
// OpenCV calls replaced by libnv12eq, in the reference's own language (C++) and call shape: N worker threads, each with
// its own context (OpenCLequalHist.cpp:289), pop NV12 frames from a queue, process them, and hand them to an "encoder"
// that wants them in capture order.  GStreamer is replaced by a synthetic frame source so that the program is
// self-contained; everything between `pop` and `push` is what a maintainer would paste into the reference.
//
//   make -C examples        (g++ -std=c++17 -O2 -Iinclude worker_demo.cpp -L../opencv-opencl_b200 -lnv12eq -lpthread)
//   ./examples/worker_demo [--op clahe|equalize] [--width 1920] [--height 1080] [--frames 240] [--workers 2]
//                          [--clipLimit 2.0] [--tile 8] [--stream]
//
// Without --stream: the reference's structure (unordered worker pool + reorder by sequence number).
// With    --stream: one nv12eq_stream per process does the queueing, overlap and ordering.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "nv12eq.h"

namespace {

struct Options {
    std::string op = "equalize";
    int width = 1920, height = 1080, frames = 240, workers = 2, tile = 8;
    double clip = 2.0;
    bool stream = false;
};

// Appendix B of SURVEY.md: the deterministic synthetic NV12 frame (luma only here; chroma is a flat gradient).
uint32_t fmix32(uint32_t k) { k ^= k >> 16; k *= 0x85EBCA6Bu; k ^= k >> 13; k *= 0xC2B2AE35u; k ^= k >> 16; return k; }
void synth_frame(std::vector<uint8_t>& f, int W, int H, uint32_t frame) {
    f.resize((size_t)W * (H + H / 2));
    const int bw = std::max(W / 16, 1), bh = std::max(H / 9, 1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint32_t k = fmix32((uint32_t)(y * W + x) * 0x9E3779B1u + 2026u * 0x85EBCA77u + frame * 0xC2B2AE3Du);
            int base = 48 + (x * 128) / W + (y * 48) / H, noise = (int)(k & 63u) - 32;
            if (((x / bw) + (y / bh)) % 5 == 0) { base = 200; noise = (int)(k & 3u); }
            f[(size_t)y * W + x] = (uint8_t)std::min(std::max(base + noise, 0), 255);
        }
    for (size_t j = 0; j < (size_t)W * (H / 2); ++j) f[(size_t)W * H + j] = (uint8_t)(112 + (j & 31));
}

struct Frame { uint64_t seq; std::vector<uint8_t> data; };

class Queue {  // GAsyncQueue stand-in
public:
    void push(Frame&& f) { { std::lock_guard<std::mutex> lk(mu_); q_.push_back(std::move(f)); } cv_.notify_one(); }
    bool pop(Frame& out) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return closed_ || !q_.empty(); });
        if (q_.empty()) return false;
        out = std::move(q_.front()); q_.pop_front();
        return true;
    }
    void close() { { std::lock_guard<std::mutex> lk(mu_); closed_ = true; } cv_.notify_all(); }
private:
    std::mutex mu_; std::condition_variable cv_; std::deque<Frame> q_; bool closed_ = false;
};

struct Counters { std::atomic<uint64_t> processed{0}, errors{0}, delivered{0}, out_of_order{0}; };

int process(nv12eq_ctx* ctx, const Options& o, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size) {
    // ---- this call replaces nextimprovement.cpp:159-168 (UV memcpy + cv::equalizeHist on the Y views) or
    // ---- clahevideo.cpp:178-201 (createCLAHE / apply / memcpy Y / memset UV)
    if (o.op == "clahe")
        return nv12eq_clahe(ctx, in, in_size, out, out_size, o.width, o.height, o.width, o.clip, o.tile, o.tile, NV12EQ_UV_GRAY128);
    return nv12eq_equalize_hist(ctx, in, in_size, out, out_size, o.width, o.height, o.width, NV12EQ_UV_COPY);
}

void run_worker_pool(const Options& o, Counters& c) {
    Queue work;
    std::mutex out_mu;
    std::map<uint64_t, std::vector<uint8_t>> reorder;   // the IMP binary's "frame-output-ordering" buffer
    uint64_t next_out = 0;
    auto deliver = [&](uint64_t seq, std::vector<uint8_t>&& f) {
        std::lock_guard<std::mutex> lk(out_mu);
        reorder.emplace(seq, std::move(f));
        while (!reorder.empty() && reorder.begin()->first == next_out) {   // gst_app_src_push_buffer would go here
            reorder.erase(reorder.begin());
            ++next_out;
            c.delivered++;
        }
    };
    std::vector<std::thread> workers;
    for (int w = 0; w < o.workers; ++w)
        workers.emplace_back([&] {
            nv12eq_ctx* ctx = nullptr;                     // one context per worker (OpenCLequalHist.cpp:289)
            if (nv12eq_create(0, o.width, o.height, 2, &ctx) != NV12EQ_OK) {
                std::fprintf(stderr, "nv12eq_create: %s\n", nv12eq_last_error_string(nullptr));
                c.errors++;
                Frame f;
                while (work.pop(f)) c.errors++;            // drop, as the reference does on processing errors
                return;
            }
            Frame f;
            while (work.pop(f)) {
                std::vector<uint8_t> out(f.data.size());
                const int st = process(ctx, o, f.data.data(), f.data.size(), out.data(), out.size());
                if (st != NV12EQ_OK) { c.errors++; continue; }   // processing_errors++ ; continue  (:132-137)
                c.processed++;
                deliver(f.seq, std::move(out));
            }
            nv12eq_destroy(ctx);
        });
    std::vector<uint8_t> src;
    for (int k = 0; k < o.frames; ++k) {
        synth_frame(src, o.width, o.height, (uint32_t)(k % 8));
        work.push(Frame{(uint64_t)k, src});
    }
    work.close();
    for (auto& t : workers) t.join();
}

void run_stream(const Options& o, Counters& c) {
    nv12eq_ctx* ctx = nullptr;
    if (nv12eq_create(0, o.width, o.height, 1, &ctx) != NV12EQ_OK) {
        std::fprintf(stderr, "nv12eq_create: %s\n", nv12eq_last_error_string(nullptr));
        c.errors++;
        return;
    }
    nv12eq_stream_config cfg{};
    cfg.op = o.op == "clahe" ? NV12EQ_OP_CLAHE : NV12EQ_OP_EQUALIZE;
    cfg.width = o.width; cfg.height = o.height; cfg.stride = o.width;
    cfg.uv_mode = o.op == "clahe" ? NV12EQ_UV_GRAY128 : NV12EQ_UV_COPY;
    cfg.clip_limit = o.clip; cfg.tiles_x = cfg.tiles_y = o.tile;
    cfg.depth = 8; cfg.full_policy = NV12EQ_FULL_BLOCK;    // max-size-buffers=8 (OpenCVequalHist.cpp:296)
    nv12eq_stream* s = nullptr;
    if (nv12eq_stream_open(ctx, &cfg, &s) != NV12EQ_OK) {
        std::fprintf(stderr, "nv12eq_stream_open: %s\n", nv12eq_last_error_string(ctx));
        c.errors++;
        nv12eq_destroy(ctx);
        return;
    }
    std::thread consumer([&] {
        std::vector<uint8_t> out((size_t)o.width * (o.height + o.height / 2));
        for (int k = 0; k < o.frames; ++k) {
            uint64_t seq = 0;
            if (nv12eq_stream_pop(s, out.data(), out.size(), &seq, 1) != NV12EQ_OK) { c.errors++; break; }
            if (seq != (uint64_t)k) c.out_of_order++;
            c.delivered++;
        }
    });
    std::vector<uint8_t> src;
    for (int k = 0; k < o.frames; ++k) {
        synth_frame(src, o.width, o.height, (uint32_t)(k % 8));
        if (nv12eq_stream_push(s, src.data(), src.size(), nullptr) == NV12EQ_OK) c.processed++; else c.errors++;
    }
    consumer.join();
    nv12eq_stream_stats st{};
    nv12eq_stream_get_stats(s, &st);
    std::printf("stream: max in flight %llu, mean latency %.2f ms, max %.2f ms\n", (unsigned long long)st.max_in_flight,
                st.delivered ? st.latency_us_sum / 1e3 / st.delivered : 0.0, st.latency_us_max / 1e3);
    nv12eq_stream_close(s);
    nv12eq_destroy(ctx);
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {   // same hand-rolled "--k v" loop as the reference (OpenCVequalHist.cpp:269-282)
        std::string a = argv[i];
        auto val = [&](const char*) { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--op") o.op = val("op");
        else if (a == "--width") o.width = std::atoi(val("w"));
        else if (a == "--height") o.height = std::atoi(val("h"));
        else if (a == "--frames") o.frames = std::atoi(val("f"));
        else if (a == "--workers") o.workers = std::max(1, std::min(8, std::atoi(val("n"))));
        else if (a == "--clipLimit") o.clip = std::atof(val("c"));
        else if (a == "--tile") o.tile = std::max(1, std::atoi(val("t")));
        else if (a == "--stream") o.stream = true;
    }
    std::printf("nv12eq %d: %s %dx%d, %d frames, %s\n", nv12eq_version(), o.op.c_str(), o.width, o.height, o.frames,
                o.stream ? "nv12eq_stream" : (std::to_string(o.workers) + " workers").c_str());
    Counters c;
    const auto t0 = std::chrono::steady_clock::now();
    if (o.stream) run_stream(o, c); else run_worker_pool(o, c);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("processed %llu, delivered in order %llu, out of order %llu, errors %llu, %.1f frames/s (incl. frame synthesis)\n",
                (unsigned long long)c.processed.load(), (unsigned long long)c.delivered.load(),
                (unsigned long long)c.out_of_order.load(), (unsigned long long)c.errors.load(), c.processed.load() / s);
    return c.errors.load() ? 1 : 0;
}
