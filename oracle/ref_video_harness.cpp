// oracle/ref_video_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference composite synthesiser (/root/reference/src/video.cpp,
// linked with player.cpp/streamer.cpp/sbc_decoder.cpp for Frame and the OS shim) into
// oracle/_ref/libefref_vid.so. Protocol = SURVEY.md §8c "Composite golden vectors":
// video_init(std); copy an I420 frame into Frame fb[0] through get_y/get_cr/get_cb;
// set the ISR's file-scope state; call video_isr() _line_count times with two alternating
// line buffers; concatenate _line_width uint16 per line.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <atomic>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>

#include "player.h"   // reference headers: Frame, video_init, ...
#undef printf

static int g_quiet = 1;
extern "C" int efref_putchar(int c) { if (!g_quiet) fputc(c, stderr); return c; }
std::string to_string(int i) { return std::to_string(i); }

// ESP-only hooks that the simulator build leaves undefined (video.cpp:309-313, 962)
void video_init_hw(int, int) {}
void ir_sample() {}
void write_pcm_16(const int16_t*, int, int) {}

// file-scope state of video.cpp (video.cpp:533-552, 936-949)
extern volatile int _line_counter;
extern volatile int _frame_counter;
extern int _line_count, _line_width, _hsync, _hsync_long, _hsync_short, _burst_start, _burst_width, _active_start;
extern int8_t _next_frame, _current_frame;
extern uint32_t _next_frame_time;
void video_reset();
void push_video(Frame* f, int front, int64_t pts, int mode);   // video.cpp:1023
extern int16_t _hscroll;
extern int16_t _animate, _animate_index;                        // poster scroll state (video.cpp:941-942)
extern Frame* _frames;
extern std::mutex _event_guard;                                    // streamer.cpp:302-303 (desktop event group)
extern std::condition_variable _event_signal;
extern uint32_t _color_tab[256 * 3];
extern int16_t* _burst0;
extern int16_t* _burst1;
extern "C" void video_isr(volatile void* vbuf);
void blit(Frame* frame, uint16_t* dst, int line, int x, int width);   // video.cpp:690 (not in video.h)

static Frame* g_fb = nullptr;
static int g_std = -1;

static void load_i420(Frame* f, const uint8_t* s)
{
    for (int y = 0; y < 192; y++, s += 352) memcpy(f->get_y(y), s, 352);
    for (int y = 0; y < 96; y++, s += 176) memcpy(f->get_cr(y), s, 176);
    for (int y = 0; y < 96; y++, s += 176) memcpy(f->get_cb(y), s, 176);
}

extern "C" {

void efref_set_quiet(int q) { g_quiet = q; }

// video_init (video.cpp:572). ntsc=1 NTSC, 0 PAL. Safe to call again to switch standard.
void efref_video_init(int ntsc)
{
    if (!g_fb) { g_fb = new Frame[2]; g_fb[0].init(); g_fb[1].init(); }
    video_init(ntsc);
    g_std = ntsc;
}

// geometry readback: [line_width, line_count, hsync, hsync_long, hsync_short, burst_start, burst_width, active_start]
void efref_geometry(int* g)
{
    g[0] = _line_width; g[1] = _line_count; g[2] = _hsync; g[3] = _hsync_long;
    g[4] = _hsync_short; g[5] = _burst_start; g[6] = _burst_width; g[7] = _active_start;
}

void efref_color_tab(uint32_t* dst) { memcpy(dst, _color_tab, sizeof(_color_tab)); }

int efref_pal_burst(int16_t* b0, int16_t* b1)
{
    if (!_burst0) return 0;
    memcpy(b0, _burst0, _burst_width * 2); memcpy(b1, _burst1, _burst_width * 2);
    return _burst_width;
}

// One whole field via video_isr (video.cpp:1122). `i420` = frame shown (fb[0]); optional
// `i420_b` = the other frame (fb[1]) used when hscroll != 0. Returns samples written.
long efref_field(const uint8_t* i420, const uint8_t* i420_b, int frame_counter, int hscroll, uint16_t* out)
{
    load_i420(&g_fb[0], i420);
    if (i420_b) load_i420(&g_fb[1], i420_b);
    _frames = g_fb; _current_frame = 0; _next_frame = -1;
    _line_counter = 0; _frame_counter = frame_counter; _hscroll = (int16_t)hscroll;
    _video_composite_blend = 0;
    uint16_t* lb[2];
    lb[0] = (uint16_t*)calloc(_line_width + 64, 2);
    lb[1] = (uint16_t*)calloc(_line_width + 64, 2);
    int lines = _line_count, w = _line_width;
    for (int l = 0; l < lines; l++) {
        video_isr(lb[l & 1]);
        memcpy(out + (size_t)l * w, lb[l & 1], (size_t)w * 2);
    }
    free(lb[0]); free(lb[1]);
    return (long)lines * w;
}

// Whole field with the presentation extras of video_isr: two-frame horizontal scroll (_hscroll,
// video.cpp:1146-1154) and the time/progress overlay composite() (video.cpp:845-887).
long efref_field_ex(const uint8_t* i420, const uint8_t* i420_b, int frame_counter, int hscroll,
                    const uint8_t* bitmap, int blend, int progress, uint16_t* out)
{
    load_i420(&g_fb[0], i420);
    if (i420_b) load_i420(&g_fb[1], i420_b);
    _frames = g_fb; _current_frame = 0; _next_frame = -1;
    _line_counter = 0; _frame_counter = frame_counter; _hscroll = (int16_t)hscroll;
    if (bitmap) memcpy(_video_composite, bitmap, VIDEO_COMPOSITE_WIDTH * VIDEO_COMPOSITE_HEIGHT);
    _video_composite_blend = blend; _video_composite_progress = progress;
    uint16_t* lb[2];
    lb[0] = (uint16_t*)calloc(_line_width + 64, 2);
    lb[1] = (uint16_t*)calloc(_line_width + 64, 2);
    int lines = _line_count, w = _line_width;
    for (int l = 0; l < lines; l++) {
        video_isr(lb[l & 1]);
        memcpy(out + (size_t)l * w, lb[l & 1], (size_t)w * 2);
    }
    free(lb[0]); free(lb[1]);
    _video_composite_blend = 0; _hscroll = 0;
    return (long)lines * w;
}

// One blit() call (video.cpp:690) into a caller buffer: the north_star "line-blit entry point".
void efref_blit(const uint8_t* i420, int frame_counter, uint16_t* dst, int line, int x, int width)
{
    load_i420(&g_fb[0], i420);
    _frame_counter = frame_counter;
    blit(&g_fb[0], dst, line, x, width);
}


// PTS -> field pacing of push_video (video.cpp:1023-1057) against the frame flip of video_isr (:1165-1177),
// under the "instant decoder" model: a decoder thread pushes frame k the moment push_video(k-1) returns, and
// the line interrupt (this thread) only advances while a frame is queued or the decoder has finished - i.e.
// time passes only while the decoder waits for presentation. Deterministic, and the only model an offline
// throughput build can have. Returns the number of whole fields emitted (<= max_fields); flip_field/flip_line[k]
// = _frame_counter and line at which frame k became _current_frame. `out` (may be NULL) receives the fields.
long efref_paced_ex(const uint8_t* i420_frames, int n_frames, const int64_t* pts, const int* modes, int frame_counter0, int max_fields, int tail_fields,
                    uint16_t* out, uint32_t* flip_field, int* flip_line, int16_t* field_hscroll);
long efref_paced(const uint8_t* i420_frames, int n_frames, const int64_t* pts, const int* modes, int frame_counter0, int max_fields,
                 uint16_t* out, uint32_t* flip_field, int* flip_line)
{
    return efref_paced_ex(i420_frames, n_frames, pts, modes, frame_counter0, max_fields, 0, out, flip_field, flip_line, nullptr);
}

// the same with `tail_fields` more fields after the last flip (the poster scroll of push_video modes 2 / 3 runs for 16 fields,
// video.cpp:1076-1088) and, optionally, the _hscroll the active lines of every field were drawn with
long efref_paced_ex(const uint8_t* i420_frames, int n_frames, const int64_t* pts, const int* modes, int frame_counter0, int max_fields, int tail_fields,
                    uint16_t* out, uint32_t* flip_field, int* flip_line, int16_t* field_hscroll)
{
    video_reset();
    _frames = g_fb; _current_frame = -1; _next_frame = -1; _next_frame_time = 0;
    _line_counter = 0; _frame_counter = frame_counter0; _hscroll = 0; _video_composite_blend = 0;
    _animate = 0; _animate_index = 0;                               // a previous run may have stopped in the middle of a scroll
    std::atomic<int> done(0), pushed(0);
    std::thread decoder([&] {
        for (int k = 0; k < n_frames; k++) {
            load_i420(&g_fb[k & 1], i420_frames + (size_t)k * 101376);      // the back buffer, as the decoder would
            pushed = k + 1;
            push_video(g_fb, k & 1, pts[k], modes ? modes[k] : 0);          // blocks until the ISR flips (mode 1: MpegDecoder::flush_picture(1))
        }
        done = 1;
    });
    uint16_t* lb[2];
    lb[0] = (uint16_t*)calloc(_line_width + 64, 2);
    lb[1] = (uint16_t*)calloc(_line_width + 64, 2);
    const int lines = _line_count, w = _line_width;
    long fields = 0;
    int flips = 0, tail = -1;
    bool stuck = false;
    const int active_top = 32 + (g_std ? 0 : 32);
    while (fields < max_fields && !stuck) {
        for (int l = 0; l < lines && !stuck; l++) {
            if (l == active_top && field_hscroll) field_hscroll[fields] = _hscroll;
            const auto t0 = std::chrono::steady_clock::now();
            while (_next_frame == -1 && !done) {                            // wait for the decoder to queue its next frame
                // the reference's desktop set_events() notifies without holding _event_guard (streamer.cpp:335-343), so a
                // wake-up can fall between push_video's predicate check and its wait(): nudge the condition variable
                // (state untouched) under the lock so a parked decoder always re-reads _event_group
                { std::lock_guard<std::mutex> g(_event_guard); }
                _event_signal.notify_all();
                std::this_thread::yield();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { stuck = true; break; }
            }
            if (stuck) break;
            const int before = _current_frame;
            const uint32_t fc = (uint32_t)_frame_counter;
            video_isr(lb[l & 1]);
            if (_current_frame != before && flips < n_frames) { flip_field[flips] = fc; flip_line[flips] = l; flips++; }
            if (out) memcpy(out + ((size_t)fields * lines + l) * w, lb[l & 1], (size_t)w * 2);
        }
        fields++;
        if (done && _next_frame == -1) {                                    // last frame is on screen for this whole field
            if (tail < 0) tail = tail_fields;
            if (tail-- <= 0) break;
        }
    }
    if (!done) {                                                            // stopped early: never leave the decoder thread parked
        _frame_counter = 0x7FFFFFF0;
        for (long guard = 0; guard < 100000000L && !done; guard++) { _line_counter = 0; _next_frame_time = 0; video_isr(lb[0]); std::this_thread::yield(); }
    }
    decoder.join();
    free(lb[0]); free(lb[1]);
    return stuck ? -1 : fields;
}

}  // extern "C"
