// oracle/ref_decode_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference decoder (/root/reference/src/player.cpp + streamer.cpp,
// compiled where they lie by oracle/Makefile into oracle/_ref/libefref_dec.so) the way the
// reference's own app does (espflix.cpp:723-737 decode_next): pop_empty -> fill Buffer with
// <=8 TS packets -> push_full, decoder thread in MpegDecoder::run (player.cpp:1355).
// The three callbacks the decoder calls out to (video.h:46-50) are stubbed here so that
// nothing paces: push_video captures the presented Frame as I420.
//
// I420 dump layout: per pushed frame 192 rows get_y (352 B), 96 rows
// get_cr (176 B), 96 rows get_cb (176 B); frames in push order incl. the final
// flush_picture(1) (player.cpp:692, Q10).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <chrono>

#include "player.h"   // reference header (include path set by the Makefile)
#undef printf

static int g_quiet = 1;
extern "C" int efref_putchar(int c) { if (!g_quiet) fputc(c, stderr); return c; }

std::string to_string(int i) { return std::to_string(i); }   // streamer.h:128 (ESP-only definition)

namespace {
struct Capture {
    uint8_t* out = nullptr;      // I420 frames
    int64_t* pts = nullptr;
    size_t cap_frames = 0;
    size_t n = 0;
};
Capture g_cap;
const size_t kI420 = 352 * 192 * 3 / 2;

void grab(Frame* f, int64_t pts)
{
    if (g_cap.n < g_cap.cap_frames && g_cap.out) {
        uint8_t* d = g_cap.out + g_cap.n * kI420;
        for (int y = 0; y < 192; y++, d += 352) memcpy(d, f->get_y(y), 352);
        for (int y = 0; y < 96; y++, d += 176) memcpy(d, f->get_cr(y), 176);
        for (int y = 0; y < 96; y++, d += 176) memcpy(d, f->get_cb(y), 176);
        if (g_cap.pts) g_cap.pts[g_cap.n] = pts;
    }
    g_cap.n++;
}
}  // namespace

// video.h:46-50 — stubs (the real ones live in video.cpp and block on the video ISR)
void push_video(Frame* f, int front, int64_t pts, int mode) { (void)mode; grab(&f[front], pts); }
void push_audio(const uint8_t*, int, int64_t, bool) {}
void video_reset() {}

extern "C" {

void efref_set_quiet(int q) { g_quiet = q; }

// Decode `loops` back-to-back copies of a transport stream. Returns number of frames pushed
// (may exceed cap_frames; only the first cap_frames are stored). One call = one decoder
// instance; the reference keeps decoder scratch in process globals (player.cpp:732), so
// calls must not overlap within a process.
long efref_decode_ts(const uint8_t* ts, size_t len, int loops,
                     uint8_t* out_i420, size_t cap_frames, int64_t* pts_out)
{
    g_cap = Capture();
    g_cap.out = out_i420; g_cap.pts = pts_out; g_cap.cap_frames = cap_frames;

    Frame* fb = new Frame[2];
    fb[0].init(); fb[1].init();
    MpegDecoder* dec = new MpegDecoder(&fb[0], &fb[1]);
    clear_events(DECODER_PAUSED);
    set_events(DECODER_RUN);
    std::thread* th = new std::thread([dec] { dec->run(); });   // never returns (parks in pause())

    for (int l = 0; l < loops; l++) {
        size_t pos = 0;
        while (pos + 188 <= len) {
            Buffer* b = dec->pop_empty();
            if (!b) continue;
            size_t n = len - pos;
            if (n > sizeof(b->data)) n = sizeof(b->data);
            n -= n % 188;
            memcpy(b->data, ts + pos, n);
            b->len = (uint32_t)n;
            pos += n;
            dec->push_full(b);
        }
    }
    Buffer* b;
    while (!(b = dec->pop_empty())) {}
    b->len = 0;                       // end of stream -> decoder synthesises SEQUENCE_END (player.cpp:456,469)
    dec->push_full(b);
    while (!(get_events() & DECODER_PAUSED))
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    dec->flush_picture(1);            // present the last picture (Q10)
    th->detach();                     // thread stays parked on its own queue / event; leaked by design
    long n = (long)g_cap.n;
    g_cap = Capture();
    return n;
}

// Frame address map probe (video.h:36-44; player.cpp:25-52): offsets relative to _slices[s]
long efref_frame_offset(int which, int y)
{
    static Frame* f = nullptr;
    if (!f) { f = new Frame; f->init(); }
    uint8_t* p = which == 0 ? f->get_y(y) : which == 1 ? f->get_cr(y) : f->get_cb(y);
    for (int s = 0; s < FB_SLICES; s++)
        if (p >= f->_slices[s] && p < f->_slices[s] + FB_STRIDE * FB_SLICE_HEIGHT)
            return (long)s * 100000 + (long)(p - f->_slices[s]);
    return -1;
}

}  // extern "C"

#ifdef EFREF_MAIN
// CLI form (one decoder per process, Q11; the parked decoder thread would block libc's
// static destructors, so the process leaves through _exit):
//   efref_decode <in.ts> <out.i420|-> [loops]      prints {"frames":N,"seconds":S} on stdout
#include <unistd.h>
#include <vector>
int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.ts out.i420|- [loops]\n", argv[0]); return 2; }
    int loops = argc > 3 ? atoi(argv[3]) : 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> ts;
    uint8_t tmp[65536]; size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) ts.insert(ts.end(), tmp, tmp + n);
    fclose(f);
    bool dump = strcmp(argv[2], "-") != 0;
    size_t cap = dump ? 4096 : 0;
    std::vector<uint8_t> out(dump ? cap * kI420 : 16);
    auto t0 = std::chrono::steady_clock::now();
    long frames = efref_decode_ts(ts.data(), ts.size(), loops, dump ? out.data() : nullptr, cap, nullptr);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dump) {
        FILE* o = fopen(argv[2], "wb");
        if (!o) { perror(argv[2]); _exit(2); }
        size_t k = (size_t)frames < cap ? (size_t)frames : cap;
        fwrite(out.data(), kI420, k, o);
        fclose(o);
    }
    fprintf(stdout, "{\"frames\": %ld, \"seconds\": %.6f}\n", frames, s);
    fflush(stdout);
    _exit(0);
}
#endif
