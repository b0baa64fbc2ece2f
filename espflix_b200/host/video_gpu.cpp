// espflix_b200/host/video_gpu.cpp — display side of the reference interface (src/video.cpp) on top of
// the C-ABI: video_init, push_video, video_isr (one scan line per call) and blit. A whole field is
// synthesised on the GPU by one K2 launch when the line counter wraps; video_isr then hands out its
// lines one by one, so a caller that drives the reference's I2S end-of-line interrupt loop
// (video.cpp:51-56) is served unchanged. No real-time pacing; the PTS -> field schedule of
// push_video is available offline (ef_set_video_pacing).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ef_player.h"
#include "espflix_b200.h"

volatile int _line_counter = 0;
volatile int _frame_counter = 0;
int16_t _hscroll = 0;
uint8_t _video_composite[VIDEO_COMPOSITE_HEIGHT * VIDEO_COMPOSITE_WIDTH];
int _video_composite_blend = 0;
int _video_composite_progress = 0;
int16_t _animate = 0, _animate_index = 0;  // poster scroll: push_video modes 2 / 3 (video.cpp:941-942, 1041)
int8_t _next_frame = -1;                  // video.cpp:936
uint32_t _next_frame_time = 0;
uint32_t _video_pts = 0, _pts_origin = 0, _video_frame_counter_origin = 0;   // video.cpp:945-948

namespace {
ef_ctx* g_ctx = nullptr;
int g_ntsc = 1, g_line_width = 912, g_line_count = 262;
Frame* g_frames = nullptr;
int g_current = -1;
std::vector<uint16_t> g_field;
std::vector<uint8_t> g_staging(EF_FRAME_BYTES);
ef_push_video_hook g_hook = nullptr;
void* g_hook_user = nullptr;
bool g_pacing = false, g_dirty = false;
ef_field_sink g_sink = nullptr;
void* g_sink_user = nullptr;
std::vector<uint16_t> g_emitted;

bool ensure_ctx()
{
    if (g_ctx) return true;
    ef_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_streams = 1; cfg.max_pictures = 1; cfg.max_slices_per_picture = 1; cfg.es_capacity = 4096; cfg.fields = 1;
    if (ef_create(&g_ctx, &cfg) != EF_OK) { fprintf(stderr, "video: %s\n", ef_last_error()); return false; }
    return true;
}

void upload(Frame* f, int fb = 0)
{
    for (int s = 0; s < FB_SLICES; s++) memcpy(g_staging.data() + (size_t)s * EF_STRIP_BYTES, f->_slices[s], EF_STRIP_BYTES);
    if (ef_write_frame(g_ctx, 0, fb, g_staging.data()) != EF_OK) fprintf(stderr, "video: %s\n", ef_last_error());
}
}  // namespace

// ease in / ease out of the poster scroll, one step per field and one at the flip (animate(), video.cpp:1076-1088)
static void animate()
{
    static const int16_t easd[16] = { 0, 8, 16, 24, 48, 72, 104, 136, 176, 216, 248, 280, 304, 328, 336, 344 };   // _easd
    if (_animate_index == 0) { _hscroll = 0; return; }
    if (_animate_index < 0) _hscroll = (int16_t)-easd[-(++_animate_index)];
    else _hscroll = easd[--_animate_index];
}

void ef_set_push_video_hook(ef_push_video_hook hook, void* user) { g_hook = hook; g_hook_user = user; }
void ef_set_video_pacing(int on, ef_field_sink sink, void* user) { g_pacing = on != 0; g_sink = sink; g_sink_user = user; }
void ef_video_set_frame_counter(int frame_counter) { _frame_counter = frame_counter; }

void video_init(int ntsc)                         // video.cpp:572
{
    if (!ensure_ctx()) return;
    g_ntsc = ntsc ? 1 : 0;
    ef_video_init(g_ctx, g_ntsc);
    ef_video_geometry(g_ctx, &g_line_width, &g_line_count);
    g_field.assign((size_t)g_line_width * g_line_count, 0);
    _line_counter = 0;
}

// ---- audio (video.cpp:957-1020, espflix.ino:73-136) ---------------------------------------------------------------
// push_audio() feeds a ring that the reference's audio thread drains frame by frame (decode_audio -> sbc_decoder ->
// write_pcm_16 -> pdm_second_order). Here the bytes are collected and ef_audio_drain() - the offline audio thread -
// decodes everything pushed since video_reset() in one batched GPU call and hands the frames not delivered yet to
// the sink, 128 PCM samples and 256 PDM words each, in order.
namespace {
std::vector<uint8_t> g_sbc;
size_t g_audio_delivered = 0;                  // frames already handed to the sink
ef_audio_sink g_audio_sink = nullptr;
void* g_audio_user = nullptr;
}  // namespace

void ef_set_audio_sink(ef_audio_sink sink, void* user) { g_audio_sink = sink; g_audio_user = user; }

void push_audio(const uint8_t* data, int len, int64_t, bool)   // video.cpp:1007
{
    if (len > 0) g_sbc.insert(g_sbc.end(), data, data + len);
}

int ef_audio_drain()
{
    if (g_sbc.empty()) return 0;
    const uint64_t off[2] = { 0, (uint64_t)g_sbc.size() };
    ef_audio_info info;
    if (ef_audio_decode(0, g_sbc.data(), off, 1, &info, nullptr, 0, nullptr) != EF_OK) { fprintf(stderr, "audio: %s\n", ef_last_error()); return 0; }
    if (info.frame_size <= 0 || info.n_frames <= g_audio_delivered) return 0;
    std::vector<int16_t> pcm((size_t)info.n_frames * 128);
    std::vector<uint16_t> pdm((size_t)info.n_frames * 256);
    if (ef_audio_decode(0, g_sbc.data(), off, 1, &info, pcm.data(), pcm.size(), pdm.data()) != EF_OK) { fprintf(stderr, "audio: %s\n", ef_last_error()); return 0; }
    int n = 0;
    for (size_t k = g_audio_delivered; k < info.n_frames; k++, n++)
        if (g_audio_sink) g_audio_sink(pcm.data() + k * 128, 128, pdm.data() + k * 256, g_audio_user);   // write_pcm_16(mono, 128, 1)
    g_audio_delivered = info.n_frames;
    return n;
}

void video_reset()                                  // video.cpp:1070
{
    _pts_origin = _video_frame_counter_origin = _video_pts = 0;
    g_sbc.clear(); g_audio_delivered = 0;          // _sbc_r = _sbc_w = _sbc_frame_size = 0
}
void video_pause(int) {}

void push_video(Frame* f, int front, int64_t pts, int mode)   // video.cpp:1023
{
    g_frames = f;
    if (!g_pacing) {                               // throughput build: the frame is current at once, nothing waits
        g_current = front;
        if (g_hook) g_hook(f, front, pts, mode, g_hook_user);
        return;
    }
    pts /= g_ntsc ? 1500 : 1800;                   // 90 kHz ticks -> field periods (60 / 50 Hz)
    _video_pts = (uint32_t)pts;
    if (_video_frame_counter_origin == 0) {
        _pts_origin = _video_pts;
        _video_frame_counter_origin = (uint32_t)_frame_counter;
    }
    uint32_t d = (_video_pts - _pts_origin) + _video_frame_counter_origin;    // field counter value at which this frame is due
    if (mode) { d = (uint32_t)_frame_counter; _animate = (int16_t)mode; }   // any non-zero mode: due now; 2 / 3 scroll the poster in
    if (d < (uint32_t)_frame_counter) {
        const int late = (int)((uint32_t)_frame_counter - d);
        printf("v late:%d\n", late);
        if (late > 2) {
            printf("resetting v timing\n");
            _video_frame_counter_origin = 0;
        }
    }
    _next_frame_time = d;
    _next_frame = (int8_t)front;
    // wait_events(VIDEO_READY): the line interrupt runs until it has flipped to this frame
    std::vector<uint16_t> line((size_t)g_line_width + 64);
    while (_next_frame != -1) video_isr(line.data());
    if (g_hook) g_hook(f, front, pts, mode, g_hook_user);
}

extern "C" void video_isr(volatile void* vbuf)    // video.cpp:1122
{
    if (!ensure_ctx()) return;
    if (g_field.empty()) video_init(1);
    const int i = _line_counter;
    {   // flip buffers in blanking (video.cpp:1165-1177): every line that is neither an active line of a
        // presented frame nor a vertical-sync line checks the queued frame
        const int active_top = 32 + (g_ntsc ? 0 : 32), active_bottom = active_top + 192;
        const int vsync_start = g_line_count - (g_ntsc ? 3 : 8);
        const bool active = i >= active_top && i < active_bottom && g_current != -1;
        if (!active && i < vsync_start && _next_frame != -1 && (uint32_t)_frame_counter >= _next_frame_time) {
            g_current = _next_frame;
            if (_animate == 2) _animate_index = -16; else if (_animate == 3) _animate_index = 16;   // video.cpp:1168-1171
            _animate = 0;
            _next_frame = -1;
            animate();
            g_dirty = true;                        // the lines below this one show the new frame
        }
    }
    if (i == 0 || g_dirty) {                       // new field (or new frame inside it): one K2 launch
        g_dirty = false;
        int fb = -2;                               // no frame presented yet: active lines are blank lines
        if (g_frames && g_current != -1) {
            upload(&g_frames[g_current], 0);
            if (_hscroll) upload(&g_frames[g_current ^ 1], 1);   // the frame scrolling in (video.cpp:1146-1154)
            fb = 0;
        }
        ef_video_set_scroll(g_ctx, _hscroll & ~7);
        ef_video_set_overlay(g_ctx, _video_composite, _video_composite_blend, _video_composite_progress);
        if (ef_composite_field(g_ctx, fb, _frame_counter, nullptr) != EF_OK || ef_read_field(g_ctx, 0, g_field.data()) != EF_OK)
            fprintf(stderr, "video: %s\n", ef_last_error());
    }
    memcpy((void*)vbuf, g_field.data() + (size_t)i * g_line_width, (size_t)g_line_width * 2);
    if (g_pacing && g_sink) {
        if (g_emitted.size() != g_field.size()) g_emitted.assign(g_field.size(), 0);
        memcpy(g_emitted.data() + (size_t)i * g_line_width, (const void*)vbuf, (size_t)g_line_width * 2);
    }
    _line_counter = i + 1;
    if (_line_counter == g_line_count) {           // end of field (video.cpp:1192-1197)
        if (g_pacing && g_sink) g_sink(g_emitted.data(), g_line_width, g_line_count, (uint32_t)_frame_counter, g_sink_user);
        _line_counter = 0; _frame_counter = _frame_counter + 1;
        if (_video_composite_blend > 0) --_video_composite_blend;
        animate();
    }
}

void blit(Frame* frame, uint16_t* dst, int line, int x, int width)   // video.cpp:690
{
    if (!ensure_ctx()) return;
    upload(frame);
    if (ef_blit(g_ctx, 0, 0, dst, line, x, width, _frame_counter) != EF_OK) fprintf(stderr, "video: %s\n", ef_last_error());
}
