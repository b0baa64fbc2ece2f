// espflix_b200/host/ef_player.h — host-side mirror of the reference's decode/display interface for
// the hot path, implemented on libespflix_b200.so. Same class names, method names, argument meaning
// and callback protocol as the reference (src/player.h:34-84 MpegDecoder, src/video.h:36-50 Frame +
// video_*/push_*, src/streamer.h:139 Buffer, src/video.cpp:690 blit, src/video.cpp:1122 video_isr),
// so a caller written against those headers compiles against this one. Only the members a caller of
// the path touches are declared; the bitstream/macroblock internals live on the GPU.
//
// Differences, all documented in INTEGRATION.md:
//   * run() returns when the end-of-stream Buffer (len == 0) arrives instead of parking forever in
//     pause() (player.cpp:1342); events/DECODER_RUN of streamer.h are the OS shim, not this path.
//   * push_video() here never blocks on the display (no real-time pacing).
//   * every picture must carry a PES PTS (true for every stream the reference's service makes);
//     quirk Q10 (no PTS -> decode in place) is not reproduced.
#ifndef EF_PLAYER_H
#define EF_PLAYER_H
#include <stdint.h>
#include <stdlib.h>

#define FB_WIDTH 352
#define FB_HEIGHT 192
#define FB_STRIDE (FB_WIDTH * 3 / 2)
#define FB_SLICE_HEIGHT 16
#define FB_SLICES (FB_HEIGHT / FB_SLICE_HEIGHT)

// 12 separately allocated strips of 16 rows x 528 bytes: 352 luma + 176 chroma per row, block-4
// chroma in strip rows 0-7, block-5 chroma in rows 8-15 (video.h:36-44, player.cpp:25-52)
class Frame {
public:
    uint8_t* _slices[FB_SLICES];
    void init();
    uint8_t* get_y(int y);
    uint8_t* get_cr(int y);
    uint8_t* get_cb(int y);
    void erase();
};

// unit of input: up to 8 transport packets; len <= 0 marks the end of the stream (streamer.h:139)
class Buffer {
public:
    uint32_t len;
    uint8_t data[8 * 188];
};

// callouts of the decoder, implemented by the display side (video.h:46-50)
void video_init(int ntsc);
void video_reset();
void video_pause(int p);
void push_video(Frame* f, int front, int64_t pts, int mode);
void push_audio(const uint8_t* data, int len, int64_t pts, bool pes_complete);
extern "C" void video_isr(volatile void* buf);                       // fills one scan line (video.cpp:1122)
void blit(Frame* frame, uint16_t* dst, int line, int x, int width);   // video.cpp:690
extern volatile int _line_counter;
extern volatile int _frame_counter;
extern int16_t _hscroll;                                             // two-frame scroll, multiple of 8 (video.cpp:940)

// time/progress overlay of the display side (video.h:52-57)
#define VIDEO_COMPOSITE_WIDTH 80
#define VIDEO_COMPOSITE_HEIGHT 16
#define VIDEO_COMPOSITE_PROGRESS_WIDTH (352 - VIDEO_COMPOSITE_WIDTH - 32)
extern uint8_t _video_composite[VIDEO_COMPOSITE_HEIGHT * VIDEO_COMPOSITE_WIDTH];
extern int _video_composite_blend;
extern int _video_composite_progress;

// test/tooling hook: observe every push_video() (the reference's harnesses stub push_video instead)
typedef void (*ef_push_video_hook)(Frame* f, int front, int64_t pts, int mode, void* user);
void ef_set_push_video_hook(ef_push_video_hook hook, void* user);

// Offline presentation pacing (video.cpp:1023-1057 against the flip in video_isr :1165-1177). When enabled,
// push_video() computes the display field from the PTS exactly like the reference, queues the frame, and then
// - instead of blocking on VIDEO_READY while a hardware interrupt runs - drives video_isr() itself until the
// frame has become the current one: time only passes while the decoder waits for presentation ("instant
// decoder"). Every completed field (line_count x line_width uint16) is handed to `sink` with the value of
// _frame_counter it was drawn under. mode != 0 shows the frame at once (as the reference) but the poster
// scroll animation (_easd) is not reproduced.
typedef void (*ef_field_sink)(const uint16_t* field, int line_width, int line_count, uint32_t frame_counter, void* user);
void ef_set_video_pacing(int on, ef_field_sink sink, void* user);
void ef_video_set_frame_counter(int frame_counter);     // _frame_counter at power-up is 0; tests start elsewhere

// Audio side (video.cpp:957-1020; espflix.ino:73-136). push_audio() collects the SBC bytes the demux routes to it;
// ef_audio_drain() is the offline audio thread: it decodes everything pushed since video_reset() on the GPU (the
// reference's decode_audio -> sbc_decoder) and hands every frame not yet delivered to the sink as 128 PCM samples
// (write_pcm_16(mono, 128, 1)) plus the 256 PDM words pdm_second_order makes of them. Returns the frames delivered.
typedef void (*ef_audio_sink)(const int16_t* pcm, int n_samples, const uint16_t* pdm, void* user);
void ef_set_audio_sink(ef_audio_sink sink, void* user);
int ef_audio_drain();

struct ef_decoder_impl;

class MpegDecoder {
public:
    Frame* _fb[2];
    int _fb_index;
    Frame* _reference;
    Frame* _current;
    int64_t _pts;
    int64_t _last_pts;

    MpegDecoder(Frame* fb0, Frame* fb1);
    ~MpegDecoder();
    void push_full(Buffer* b);      // producer -> decoder
    Buffer* pop_empty();            // decoder -> producer (4 Buffers circulate, player.cpp:367)
    void reset();
    void run();                     // body of the decoder thread
    int64_t get_pts();
    void flush_picture(int mode = 0);

private:
    ef_decoder_impl* _impl;
    MpegDecoder(const MpegDecoder&);
    MpegDecoder& operator=(const MpegDecoder&);
};

#endif
