// espflix_b200/host/player_gpu.cpp — MpegDecoder/Frame (reference: src/player.cpp) on top of the
// C-ABI. The host keeps only the transport side: Buffer queues, TS/PES demux (more()/demux(),
// player.cpp:381-493: video PID into the access-unit cutter, audio PIDs to push_audio with the reference's
// "no PTS mutes the stream" state), and the push_video/buffer-swap protocol of flush_picture() (player.cpp:692).
// Every picture is decoded by the CUDA kernels, and not one by one: complete access units are collected and a whole
// batch (kBatchPictures) goes through ONE ef_submit_es_host -> ef_index -> ef_decode_all_to_host; the callbacks
// then fire in the reference's order with the reference's Frame contents (the decoder runs ahead of presentation
// by at most one batch, as the reference's decoder thread runs ahead of its display by one frame).
#include "ef_player.h"

#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

#include "espflix_b200.h"

// ---- Frame (player.cpp:25-52) ---------------------------------------------------------------------
void Frame::init()
{
    for (int i = 0; i < FB_SLICES; i++) {
        _slices[i] = (uint8_t*)malloc(FB_STRIDE * FB_SLICE_HEIGHT + 4);
        memset(_slices[i], 0, FB_STRIDE * FB_SLICE_HEIGHT + 4);
    }
}
uint8_t* Frame::get_y(int y) { return _slices[y >> 4] + (y & 15) * FB_STRIDE; }
uint8_t* Frame::get_cr(int y) { return _slices[y >> 3] + (y & 7) * FB_STRIDE + FB_WIDTH; }
uint8_t* Frame::get_cb(int y) { return _slices[y >> 3] + ((y & 7) + 8) * FB_STRIDE + FB_WIDTH; }
void Frame::erase()
{
    for (int i = 0; i < FB_SLICES; i++) memset(_slices[i], 0x30, FB_STRIDE * FB_SLICE_HEIGHT + 4);
}

namespace {
struct BufferQueue {
    std::deque<Buffer*> q;
    std::mutex m;
    std::condition_variable cv;
    void push(Buffer* b) { { std::lock_guard<std::mutex> l(m); q.push_back(b); } cv.notify_one(); }
    Buffer* pop() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !q.empty(); }); Buffer* b = q.front(); q.pop_front(); return b; }
    void drain_into(BufferQueue& other) { std::lock_guard<std::mutex> l(m); while (!q.empty()) { other.push(q.front()); q.pop_front(); } }
};

int be16(const uint8_t* d) { return (d[0] << 8) | d[1]; }

int64_t pes_timestamp(const uint8_t* d, int flags)       // parse_pts, player.cpp:299
{
    flags = (flags >> 2) & 0x30;
    if ((d[0] & 0xF0) != flags) return -1;
    int64_t n = ((int64_t)(d[0] & 0x0E)) << 29;
    n += (int64_t)(be16(d + 1) >> 1) << 15;
    return n + (be16(d + 3) >> 1);
}
}  // namespace

constexpr int kBatchPictures = 32;             // access units per GPU submit
constexpr size_t kBatchBytes = 768 << 10;       // ... or this much elementary stream, whichever comes first (es_capacity is 1 MB)

struct ef_decoder_impl {
    ef_ctx* ctx = nullptr;
    BufferQueue full, empty;
    Buffer pool[4];
    std::vector<uint8_t> es;                   // video ES not yet decoded
    std::vector<std::pair<size_t, int64_t>> pes;   // (offset in es, pts) of PES headers seen
    std::vector<size_t> unit_end;              // complete access units at the front of `es`: end offsets
    std::vector<int64_t> unit_pts;             // PTS latched by the picture header of each unit (-1: keep)
    uint8_t* frames = nullptr;                 // pinned: kBatchPictures decoded pictures in the strip layout
    bool in_picture = false;                   // a picture start code has been seen in `es`
    size_t scan_from = 0, unit_start = 0;
    int64_t audio_pts = -1;                    // _audio_pts (player.cpp:418-431)
};

MpegDecoder::MpegDecoder(Frame* fb0, Frame* fb1)
{
    _fb[0] = fb0; _fb[1] = fb1;
    _fb_index = 0;
    _reference = _fb[_fb_index++ & 1];          // player.cpp:359
    _current = _fb[_fb_index & 1];
    _last_pts = _pts = -1;
    _impl = new ef_decoder_impl();
    ef_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.n_streams = 1; cfg.max_pictures = kBatchPictures; cfg.max_slices_per_picture = 64;
    cfg.es_capacity = 1 << 20; cfg.fields = 0;
    if (ef_create(&_impl->ctx, &cfg) != EF_OK || ef_host_alloc((void**)&_impl->frames, (size_t)kBatchPictures * EF_FRAME_BYTES) != EF_OK) {
        fprintf(stderr, "MpegDecoder: %s\n", ef_last_error());   // no CPU decoder to fall back to
        abort();
    }
    for (int i = 0; i < 4; i++) _impl->empty.push(&_impl->pool[i]);
}

MpegDecoder::~MpegDecoder()
{
    ef_host_free(_impl->frames);
    ef_destroy(_impl->ctx);
    delete _impl;
}

void MpegDecoder::push_full(Buffer* b) { _impl->full.push(b); }
Buffer* MpegDecoder::pop_empty() { return _impl->empty.pop(); }
int64_t MpegDecoder::get_pts() { return _last_pts; }

void MpegDecoder::reset()                       // player.cpp:439
{
    _impl->full.drain_into(_impl->empty);
    _impl->es.clear(); _impl->pes.clear(); _impl->unit_end.clear(); _impl->unit_pts.clear();
    _impl->in_picture = false; _impl->scan_from = 0; _impl->unit_start = 0; _impl->audio_pts = -1;
    ef_reset(_impl->ctx);
    _fb_index = 0;
    _reference = _fb[_fb_index++ & 1];
    _current = _fb[_fb_index & 1];
    video_reset();
    _last_pts = -1;
}

void MpegDecoder::flush_picture(int mode)       // player.cpp:692
{
    if (_last_pts != -1 || mode) {
        push_video(_fb[0], _fb_index & 1, _last_pts, mode);
        _reference = _fb[_fb_index++ & 1];
        _current = _fb[_fb_index & 1];
    }
    if (!mode) _last_pts = _pts;
}

// Decode the collected access units in ONE submit and replay the reference's per-picture protocol: at a picture
// header the decoder latches the PTS, presents the previous picture (flush_picture, player.cpp:692-702) and swaps
// buffers; the slices then fill _current.
static void decode_batch(MpegDecoder* d, ef_decoder_impl* im)
{
    const int n = (int)im->unit_end.size();
    if (!n) return;
    const size_t bytes = im->unit_end.back();
    const uint64_t off[2] = { 0, (uint64_t)bytes };
    int rc = ef_submit_es_host(im->ctx, im->es.data(), off, nullptr);
    if (rc == EF_OK) rc = ef_index(im->ctx, nullptr);
    if (rc == EF_OK) rc = ef_decode_all_to_host(im->ctx, n, im->frames, 1, nullptr);
    if (rc == EF_OK) rc = ef_sync(im->ctx, nullptr);
    if (rc != EF_OK) fprintf(stderr, "MpegDecoder: %s\n", ef_last_error());   // reference convention: print and keep going
    for (int k = 0; k < n; k++) {
        if (im->unit_pts[k] != -1) d->_pts = im->unit_pts[k];
        d->flush_picture(0);
        if (rc == EF_OK) {
            const uint8_t* src = im->frames + (size_t)k * EF_FRAME_BYTES;
            Frame* f = d->_current;
            for (int s = 0; s < FB_SLICES; s++) memcpy(f->_slices[s], src + (size_t)s * EF_STRIP_BYTES, EF_STRIP_BYTES);
        }
    }
    // drop what has been decoded, keep PES marks and scan positions relative to the new origin
    im->es.erase(im->es.begin(), im->es.begin() + (long)bytes);
    std::vector<std::pair<size_t, int64_t>> keep;
    for (auto& pp : im->pes) if (pp.first >= bytes) keep.push_back({ pp.first - bytes, pp.second });
    im->pes.swap(keep);
    im->unit_end.clear(); im->unit_pts.clear();
    im->scan_from -= bytes; im->unit_start -= bytes;
}

// cut complete access units off `es`: a unit ends where, after a picture start code, the next sequence / GOP /
// picture / sequence-end start code begins (marker(), player.cpp:1318)
static void drain(MpegDecoder* d, ef_decoder_impl* im, bool final)
{
    std::vector<uint8_t>& es = im->es;
    size_t i = im->scan_from;
    while (i + 4 <= es.size()) {
        if (es[i] != 0 || es[i + 1] != 0 || es[i + 2] != 1) { i++; continue; }
        const uint8_t code = es[i + 3];
        const bool header = code == 0x00 || code == 0xB3 || code == 0xB8 || code == 0xB7;
        if (header && im->in_picture) {
            im->unit_end.push_back(i);
            im->unit_start = i;
            im->in_picture = false;
        }
        if (code == 0x00) {                                       // picture(): the PTS of the PES this header lies in
            int64_t pts = -1;
            for (auto& pp : im->pes) if (pp.first <= i + 3) pts = pp.second;
            im->unit_pts.push_back(pts);
            im->in_picture = true;
        }
        i += 4;
    }
    im->scan_from = i;                     // everything before i has been examined; a start code straddling the next Buffer starts at or after i
    if (final && im->in_picture && es.size() > im->unit_start) {
        im->unit_end.push_back(es.size());
        im->unit_start = es.size();
        im->in_picture = false;
        im->scan_from = es.size();
    }
    if (final || (int)im->unit_end.size() >= kBatchPictures || (!im->unit_end.empty() && es.size() >= kBatchBytes)) {
        // a picture header already seen beyond the last complete unit keeps its latched PTS for the next batch
        std::vector<int64_t> carry(im->unit_pts.begin() + (long)im->unit_end.size(), im->unit_pts.end());
        im->unit_pts.resize(im->unit_end.size());
        decode_batch(d, im);
        im->unit_pts = carry;
    }
}

void MpegDecoder::run()                          // player.cpp:1355
{
    ef_decoder_impl* im = _impl;
    for (;;) {
        Buffer* b = im->full.pop();
        if ((int32_t)b->len <= 0) {              // end of stream (player.cpp:469)
            im->empty.push(b);
            drain(this, im, true);
            return;
        }
        for (uint32_t mark = 0; mark + 188 <= b->len; mark += 188) {        // more(), player.cpp:459
            const uint8_t* p = b->data + mark;
            if (p[0] != 0x47) { fprintf(stderr, "ts lost sync\n"); continue; }
            const int pid = ((p[1] << 8) + p[2]) & 0x1fff;
            const uint8_t* data = p + 4;
            if (p[3] & 0x20) data = p + 5 + p[4];
            if (!(p[3] & 0x10)) continue;
            const uint8_t* end = p + 188;
            const uint8_t* payload = data;
            int64_t pts = -1;
            const bool pus = p[1] & 0x40;
            if (pus) {                           // demux(), player.cpp:387-406
                const uint8_t* h = data + 6;
                const int flags = be16(h);
                payload = h + 3 + h[2];
                if (flags & 0x0080) pts = pes_timestamp(h + 3, flags);
            }
            if (pid == 0x100) {
                if (pus) im->pes.push_back({ im->es.size(), pts });
                if (payload < end) im->es.insert(im->es.end(), payload, end);
            } else if (pid == 0x101 || pid == 0x102) {
                if (pus) im->audio_pts = pts;    // a PES without a (well-formed) PTS mutes the stream until the next one with (player.cpp:418-431)
                if (im->audio_pts != -1 && payload < end) push_audio(payload, (int)(end - payload), pts, false);
            }
        }
        im->empty.push(b);
        drain(this, im, false);
    }
}
