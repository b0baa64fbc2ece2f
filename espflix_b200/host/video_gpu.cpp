// espflix_b200/host/video_gpu.cpp — display side of the reference interface (src/video.cpp) on top of
// the C-ABI: video_init, push_video, video_isr (one scan line per call) and blit. A whole field is
// synthesised on the GPU by one K2 launch when the line counter wraps; video_isr then hands out its
// lines one by one, so a caller that drives the reference's I2S end-of-line interrupt loop
// (video.cpp:51-56) is served unchanged. No real-time pacing.
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

namespace {
ef_ctx* g_ctx = nullptr;
int g_ntsc = 1, g_line_width = 912, g_line_count = 262;
Frame* g_frames = nullptr;
int g_current = -1;
std::vector<uint16_t> g_field;
std::vector<uint8_t> g_staging(EF_FRAME_BYTES);
ef_push_video_hook g_hook = nullptr;
void* g_hook_user = nullptr;

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

void ef_set_push_video_hook(ef_push_video_hook hook, void* user) { g_hook = hook; g_hook_user = user; }

void video_init(int ntsc)                         // video.cpp:572
{
    if (!ensure_ctx()) return;
    g_ntsc = ntsc ? 1 : 0;
    ef_video_init(g_ctx, g_ntsc);
    ef_video_geometry(g_ctx, &g_line_width, &g_line_count);
    g_field.assign((size_t)g_line_width * g_line_count, 0);
    _line_counter = 0;
}

void video_reset() {}
void video_pause(int) {}
void push_audio(const uint8_t*, int, int64_t, bool) {}   // audio side-chain is out of scope

void push_video(Frame* f, int front, int64_t pts, int mode)   // video.cpp:1023 without the wait on VIDEO_READY
{
    g_frames = f;
    g_current = front;
    if (g_hook) g_hook(f, front, pts, mode, g_hook_user);
}

extern "C" void video_isr(volatile void* vbuf)    // video.cpp:1122
{
    if (!ensure_ctx()) return;
    if (g_field.empty()) video_init(1);
    const int i = _line_counter;
    if (i == 0) {                                  // new field: one K2 launch
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
    _line_counter = i + 1;
    if (_line_counter == g_line_count) {           // end of field (video.cpp:1192-1197)
        _line_counter = 0; _frame_counter = _frame_counter + 1;
        if (_video_composite_blend > 0) --_video_composite_blend;
    }
}

void blit(Frame* frame, uint16_t* dst, int line, int x, int width)   // video.cpp:690
{
    if (!ensure_ctx()) return;
    upload(frame);
    if (ef_blit(g_ctx, 0, 0, dst, line, x, width, _frame_counter) != EF_OK) fprintf(stderr, "video: %s\n", ef_last_error());
}
