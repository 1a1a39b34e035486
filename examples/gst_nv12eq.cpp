// gst_nv12eq.cpp -- GStreamer appsink -> libnv12eq -> appsrc shim.
//
// The reference's video programs are two GStreamer pipelines with an OpenCV worker pool between them
// (OpenCVequalHist.cpp:292-332 capture / encode pipelines, :71-98 appsink callback, :102-196 worker, :397-402 pool;
// clahevideo.cpp:178-201 the CLAHE frame body).  This program keeps the pipelines and the command line and replaces the pool with
// one nv12eq_stream: the appsink callback pushes the mapped NV12 buffer (stride and plane offset from GstVideoMeta when present,
// as nextimprovement.cpp:128-170 does), a delivery thread pops finished frames IN ORDER and pushes them into appsrc with the
// input's timestamps.  Back-pressure is the stream's drop-oldest policy, i.e. the leaky=downstream queues of the reference.
//
// Build (needs the GStreamer development packages, which this repo's build image does not have -- the file is compiled only by
// `make -C examples gst_nv12eq`, never by the test suite):
//   g++ -std=c++17 -O2 gst_nv12eq.cpp -I../include -L../opencv-opencl_b200 -lnv12eq
//       $(pkg-config --cflags --libs gstreamer-1.0 gstreamer-app-1.0 gstreamer-video-1.0) -o gst_nv12eq
// Run:  gst_nv12eq [--clahe] [--clip 2.0] [--tiles 8] [--width 1920] [--height 1080] [--fps 60] [--bitrate 20000] [--h265]
//                  [--src "<capture pipeline ending in appsink name=cv_sink>"] [--sink "<pipeline starting with appsrc name=my_src>"]
#include <gst/app/gstappsink.h>
#include <gst/app/gstappsrc.h>
#include <gst/gst.h>
#include <gst/video/video.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "nv12eq.h"

struct Shim {
    nv12eq_ctx* ctx = nullptr;
    nv12eq_stream* stream = nullptr;
    GstElement* appsrc = nullptr;
    int width = 1920, height = 1080;
    size_t frame_bytes = 0;
    std::mutex mu;
    std::deque<std::pair<uint64_t, GstClockTime>> stamps;   // (sequence number, pts) of pushed frames
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> in_frames{0}, out_frames{0}, dropped{0};
    std::vector<uint8_t> packed;                            // repacked input when the buffer's planes are not contiguous
};

// appsink callback: O(1) apart from the one copy into the stream's pinned staging (OpenCVequalHist.cpp:71-98)
static GstFlowReturn on_sample(GstAppSink* sink, gpointer user) {
    Shim* d = static_cast<Shim*>(user);
    GstSample* sample = gst_app_sink_pull_sample(sink);
    if (!sample) return GST_FLOW_ERROR;
    GstBuffer* buf = gst_sample_get_buffer(sample);
    GstMapInfo map;
    if (!buf || !gst_buffer_map(buf, &map, GST_MAP_READ)) { gst_sample_unref(sample); return GST_FLOW_ERROR; }
    const uint8_t* frame = map.data;
    size_t size = map.size;
    // plane offsets / strides of the capture buffer (nextimprovement.cpp:128-140): repack when they are not the packed layout
    if (GstVideoMeta* m = gst_buffer_get_video_meta(buf)) {
        const size_t sy = (size_t)m->stride[0], suv = (size_t)m->stride[1];
        if (sy != (size_t)d->width || suv != (size_t)d->width || m->offset[1] != (gsize)d->width * d->height) {
            d->packed.resize(d->frame_bytes);
            for (int r = 0; r < d->height; ++r) memcpy(&d->packed[(size_t)r * d->width], map.data + m->offset[0] + r * sy, d->width);
            for (int r = 0; r < d->height / 2; ++r)
                memcpy(&d->packed[(size_t)(d->height + r) * d->width], map.data + m->offset[1] + r * suv, d->width);
            frame = d->packed.data();
            size = d->packed.size();
        }
    }
    uint64_t seq = 0;
    const int rc = nv12eq_stream_push(d->stream, frame, size, &seq);
    if (rc == NV12EQ_OK) {
        std::lock_guard<std::mutex> lk(d->mu);
        d->stamps.emplace_back(seq, GST_BUFFER_PTS(buf));
        d->in_frames++;
    } else {
        d->dropped++;
    }
    gst_buffer_unmap(buf, &map);
    gst_sample_unref(sample);
    return GST_FLOW_OK;
}

// delivery thread: finished frames, in capture order, into the encoder pipeline (OpenCVequalHist.cpp:165-190)
static void deliver(Shim* d) {
    GstBuffer* out = nullptr;
    while (!d->stop.load()) {
        if (!out) out = gst_buffer_new_allocate(nullptr, d->frame_bytes, nullptr);
        GstMapInfo map;
        gst_buffer_map(out, &map, GST_MAP_WRITE);
        uint64_t seq = 0;
        const int rc = nv12eq_stream_pop(d->stream, map.data, map.size, &seq, /*block=*/0);   // polling: close() must not race a blocked pop
        gst_buffer_unmap(out, &map);
        if (rc == NV12EQ_ERR_EMPTY) { g_usleep(200); continue; }
        if (rc != NV12EQ_OK) continue;   // a failed frame is dropped like any other (the stream has reset itself)
        GstClockTime pts = GST_CLOCK_TIME_NONE;
        {
            std::lock_guard<std::mutex> lk(d->mu);
            while (!d->stamps.empty() && d->stamps.front().first < seq) d->stamps.pop_front();   // dropped by back-pressure
            if (!d->stamps.empty() && d->stamps.front().first == seq) { pts = d->stamps.front().second; d->stamps.pop_front(); }
        }
        GST_BUFFER_PTS(out) = pts;
        if (gst_app_src_push_buffer(GST_APP_SRC(d->appsrc), out) != GST_FLOW_OK) { out = nullptr; break; }   // takes ownership of `out`
        out = nullptr;
        d->out_frames++;
    }
    if (out) gst_buffer_unref(out);
}

int main(int argc, char** argv) {
    gst_init(&argc, &argv);
    Shim d;
    bool clahe = false, h265 = false;
    double clip = 2.0;
    int tiles = 8, fps = 60, bitrate = 20000;
    std::string src, sink;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* def) { return i + 1 < argc ? argv[++i] : def; };
        if (a == "--clahe") clahe = true;
        else if (a == "--h265") h265 = true;
        else if (a == "--clip") clip = atof(next("2.0"));
        else if (a == "--tiles") tiles = atoi(next("8"));
        else if (a == "--width") d.width = atoi(next("1920"));
        else if (a == "--height") d.height = atoi(next("1080"));
        else if (a == "--fps") fps = atoi(next("60"));
        else if (a == "--bitrate") bitrate = atoi(next("20000"));
        else if (a == "--src") src = next("");
        else if (a == "--sink") sink = next("");
    }
    d.frame_bytes = (size_t)d.width * d.height * 3 / 2;
    char tmp[2048];
    if (src.empty()) {   // the reference's capture pipeline (OpenCVequalHist.cpp:292-299)
        snprintf(tmp, sizeof tmp,
                 "v4l2src device=/dev/video0 io-mode=4 ! video/x-raw,format=NV12,width=%d,height=%d,framerate=60/1 ! "
                 "videorate drop-only=true max-rate=%d ! queue name=q_cam leaky=downstream max-size-buffers=8 max-size-time=0 max-size-bytes=0 ! "
                 "appsink name=cv_sink emit-signals=true max-buffers=1 drop=true sync=false",
                 d.width, d.height, fps);
        src = tmp;
    }
    if (sink.empty()) {  // the reference's encode pipeline (:305-331), software encoders instead of the board's OMX elements
        snprintf(tmp, sizeof tmp,
                 "appsrc name=my_src is-live=true format=GST_FORMAT_TIME do-timestamp=true ! video/x-raw,format=NV12,width=%d,height=%d,framerate=%d/1 ! "
                 "queue name=q_after_src leaky=downstream max-size-buffers=2 max-size-time=0 max-size-bytes=0 ! videoconvert ! "
                 "%s bitrate=%d ! %s ! udpsink host=127.0.0.1 port=5004 async=false",
                 d.width, d.height, fps, h265 ? "x265enc tune=zerolatency" : "x264enc tune=zerolatency", bitrate, h265 ? "rtph265pay" : "rtph264pay");
        sink = tmp;
    }
    GError* err = nullptr;
    GstElement* sink_pipe = gst_parse_launch(src.c_str(), &err);
    if (!sink_pipe) { fprintf(stderr, "capture pipeline: %s\n", err ? err->message : "?"); return 1; }
    GstElement* src_pipe = gst_parse_launch(sink.c_str(), &err);
    if (!src_pipe) { fprintf(stderr, "encode pipeline: %s\n", err ? err->message : "?"); return 1; }
    GstElement* appsink = gst_bin_get_by_name(GST_BIN(sink_pipe), "cv_sink");
    d.appsrc = gst_bin_get_by_name(GST_BIN(src_pipe), "my_src");
    if (!appsink || !d.appsrc) { fprintf(stderr, "pipelines need appsink name=cv_sink and appsrc name=my_src\n"); return 1; }

    if (nv12eq_create(0, d.width, d.height, 8, &d.ctx) != NV12EQ_OK) { fprintf(stderr, "nv12eq_create failed (no CUDA device?)\n"); return 1; }
    nv12eq_stream_config cfg{};
    cfg.op = clahe ? NV12EQ_OP_CLAHE : NV12EQ_OP_EQUALIZE;
    cfg.width = d.width; cfg.height = d.height; cfg.stride = d.width;
    cfg.uv_mode = NV12EQ_UV_COPY;                 // nextimprovement.cpp:160 (GRAY128 = OpenCVequalHist.cpp:162)
    cfg.clip_limit = clip; cfg.tiles_x = tiles; cfg.tiles_y = tiles;
    cfg.depth = 8;                                // max-size-buffers=8
    cfg.full_policy = NV12EQ_FULL_DROP_OLDEST;    // leaky=downstream / drop=true
    if (nv12eq_stream_open(d.ctx, &cfg, &d.stream) != NV12EQ_OK) { fprintf(stderr, "stream: %s\n", nv12eq_last_error_string(d.ctx)); return 1; }

    GstAppSinkCallbacks cbs{};
    cbs.new_sample = on_sample;
    gst_app_sink_set_callbacks(GST_APP_SINK(appsink), &cbs, &d, nullptr);
    std::thread out_thread(deliver, &d);
    gst_element_set_state(src_pipe, GST_STATE_PLAYING);
    gst_element_set_state(sink_pipe, GST_STATE_PLAYING);

    GstBus* bus = gst_element_get_bus(sink_pipe);
    for (;;) {   // once a second: the counters the reference prints (OpenCVequalHist.cpp:40-57)
        GstMessage* msg = gst_bus_timed_pop_filtered(bus, GST_SECOND, (GstMessageType)(GST_MESSAGE_ERROR | GST_MESSAGE_EOS));
        nv12eq_stream_stats st{};
        nv12eq_stream_get_stats(d.stream, &st);
        printf("in %llu out %llu dropped %llu (stream: pushed %llu delivered %llu back-pressure drops %llu)\n", (unsigned long long)d.in_frames.load(),
               (unsigned long long)d.out_frames.load(), (unsigned long long)d.dropped.load(), (unsigned long long)st.pushed,
               (unsigned long long)st.delivered, (unsigned long long)st.dropped_backpressure);
        if (msg) { gst_message_unref(msg); break; }
    }
    d.stop = true;
    gst_element_set_state(sink_pipe, GST_STATE_NULL);   // no more pushes
    out_thread.join();
    nv12eq_stream_close(d.stream);
    gst_app_src_end_of_stream(GST_APP_SRC(d.appsrc));
    gst_element_set_state(src_pipe, GST_STATE_NULL);
    nv12eq_destroy(d.ctx);
    gst_object_unref(bus);
    gst_object_unref(appsink);
    gst_object_unref(d.appsrc);
    gst_object_unref(sink_pipe);
    gst_object_unref(src_pipe);
    return 0;
}
