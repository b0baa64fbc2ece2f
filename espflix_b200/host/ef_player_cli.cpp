// espflix_b200/host/ef_player_cli.cpp — drives the mirrored MpegDecoder exactly the way the reference's
// own app drives the original (espflix.cpp:723-737 decode_next; the test harness for the unmodified
// reference follows the same protocol):
// pop_empty -> fill Buffer with <= 8 TS packets -> push_full, decoder thread in run(), frames
// captured from push_video, final flush_picture(1). Usage: ef_player_cli in.ts out.i420 [field.u16 ntsc]
//   or, with the offline PTS -> field pacing: ef_player_cli in.ts out.i420 --paced fields.u16 ntsc frame_counter0 max_fields
//      (a 9th argument M = 2 | 3 presents the last picture as load_poster() does, flush_picture(M), and lets the poster
//       scroll run for 17 more fields: espflix.cpp:1068, video.cpp:1041)
//   or, with the audio of the program:        ef_player_cli in.ts out.i420 --audio out.pcm out.pdm
// Prints {"frames": N, "seconds": S, "frames_per_s": R, ...}: the level-1 drop-in rate of one stream (run() to join).
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "ef_player.h"

static std::vector<uint8_t> g_out;
static long g_frames = 0;

static std::vector<uint16_t> g_fields;
static long g_n_fields = 0, g_max_fields = 0;
static std::vector<uint32_t> g_flip_field;
static uint32_t g_last_fc = 0;

static void on_field(const uint16_t* field, int w, int lines, uint32_t fc, void*)
{
    if (g_n_fields < g_max_fields) g_fields.insert(g_fields.end(), field, field + (size_t)w * lines);
    g_n_fields++;
    g_last_fc = fc;
}

static std::vector<int16_t> g_pcm;
static std::vector<uint16_t> g_pdm;
static void on_audio(const int16_t* pcm, int n, const uint16_t* pdm, void*)
{
    g_pcm.insert(g_pcm.end(), pcm, pcm + n);
    g_pdm.insert(g_pdm.end(), pdm, pdm + 2 * n);
}

static void on_push(Frame* f, int front, int64_t, int, void*)
{
    Frame* fr = &f[front];
    size_t o = g_out.size();
    g_out.resize(o + 352 * 192 * 3 / 2);
    uint8_t* d = g_out.data() + o;
    for (int y = 0; y < 192; y++, d += 352) memcpy(d, fr->get_y(y), 352);
    for (int y = 0; y < 96; y++, d += 176) memcpy(d, fr->get_cr(y), 176);
    for (int y = 0; y < 96; y++, d += 176) memcpy(d, fr->get_cb(y), 176);
    g_frames++;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.ts out.i420 [field.u16 ntsc(1|0)]\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> ts;
    uint8_t tmp[65536]; size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) ts.insert(ts.end(), tmp, tmp + n);
    fclose(f);

    ef_set_push_video_hook(on_push, nullptr);
    const bool paced = argc >= 8 && !strcmp(argv[3], "--paced");
    if (paced) {
        video_init(atoi(argv[5]));
        ef_video_set_frame_counter(atoi(argv[6]));
        g_max_fields = atol(argv[7]);
        ef_set_video_pacing(1, on_field, nullptr);
    }
    const bool audio = argc >= 6 && !strcmp(argv[3], "--audio");
    if (audio) ef_set_audio_sink(on_audio, nullptr);
    Frame fb[2];
    fb[0].init(); fb[1].init();
    MpegDecoder dec(&fb[0], &fb[1]);
    const auto t0 = std::chrono::steady_clock::now();
    std::thread th([&] { dec.run(); });
    size_t pos = 0;
    while (pos + 188 <= ts.size()) {
        Buffer* b = dec.pop_empty();
        size_t k = ts.size() - pos;
        if (k > sizeof(b->data)) k = sizeof(b->data);
        k -= k % 188;
        memcpy(b->data, ts.data() + pos, k);
        b->len = (uint32_t)k;
        pos += k;
        dec.push_full(b);
    }
    Buffer* b = dec.pop_empty();
    b->len = 0;
    dec.push_full(b);
    th.join();
    const int last_mode = paced && argc >= 9 ? atoi(argv[8]) : 1;
    dec.flush_picture(last_mode);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (audio) {
        const int nf = ef_audio_drain();
        FILE* a = fopen(argv[4], "wb"); fwrite(g_pcm.data(), 2, g_pcm.size(), a); fclose(a);
        a = fopen(argv[5], "wb"); fwrite(g_pdm.data(), 2, g_pdm.size(), a); fclose(a);
        fprintf(stderr, "audio: %d frames\n", nf);
    }
    FILE* o = fopen(argv[2], "wb");
    fwrite(g_out.data(), 1, g_out.size(), o);
    fclose(o);
    if (paced) {
        // the field in which the last frame flipped is completed by the line interrupt, as it would be on air
        std::vector<uint16_t> line(1136 + 64);
        while (_line_counter != 0) video_isr(line.data());
        if (last_mode > 1)                         // the poster scroll: 17 more fields
            for (int f = 0; f < 17; f++) do video_isr(line.data()); while (_line_counter != 0);
        FILE* ff = fopen(argv[4], "wb");
        fwrite(g_fields.data(), 2, g_fields.size(), ff);
        fclose(ff);
        printf("{\"frames\": %ld, \"fields\": %ld, \"seconds\": %.6f}\n", g_frames, g_n_fields, secs);
        return 0;
    }
    if (argc >= 5 && !audio) {                     // one field of the last presented frame through video_isr
        const int ntsc = atoi(argv[4]);
        video_init(ntsc);
        const int w = ntsc ? 912 : 1136, lines = ntsc ? 262 : 312;
        std::vector<uint16_t> field((size_t)w * lines), line(w + 64);
        for (int l = 0; l < lines; l++) { video_isr(line.data()); memcpy(field.data() + (size_t)l * w, line.data(), (size_t)w * 2); }
        FILE* ff = fopen(argv[3], "wb");
        fwrite(field.data(), 2, field.size(), ff);
        fclose(ff);
    }
    printf("{\"frames\": %ld, \"seconds\": %.6f, \"frames_per_s\": %.1f, \"audio_frames\": %zu}\n", g_frames, secs, g_frames / secs, g_pcm.size() / 128);
    return 0;
}
