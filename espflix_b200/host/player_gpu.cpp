// espflix_b200/host/player_gpu.cpp — MpegDecoder/Frame (reference: src/player.cpp) on top of the
// C-ABI. The host keeps only the transport side: Buffer queues, TS/PES demux of the video PID
// (more()/demux(), player.cpp:381-493), access-unit cutting and the push_video/buffer-swap protocol
// of flush_picture() (player.cpp:692). Every picture is decoded by the CUDA kernels.
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

struct ef_decoder_impl {
    ef_ctx* ctx = nullptr;
    BufferQueue full, empty;
    Buffer pool[4];
    std::vector<uint8_t> es;                   // video ES not yet decoded
    std::vector<std::pair<size_t, int64_t>> pes;   // (offset in es, pts) of PES headers seen
    std::vector<uint8_t> staging;
    bool in_picture = false;                   // a picture start code has been seen in `es`
    size_t scan_from = 0;
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
    cfg.device = 0; cfg.n_streams = 1; cfg.max_pictures = 4; cfg.max_slices_per_picture = 64;
    cfg.es_capacity = 1 << 20; cfg.fields = 0;
    if (ef_create(&_impl->ctx, &cfg) != EF_OK) {
        fprintf(stderr, "MpegDecoder: %s\n", ef_last_error());   // no CPU decoder to fall back to
        abort();
    }
    _impl->staging.resize(EF_FRAME_BYTES);
    for (int i = 0; i < 4; i++) _impl->empty.push(&_impl->pool[i]);
}

MpegDecoder::~MpegDecoder()
{
    ef_destroy(_impl->ctx);
    delete _impl;
}

void MpegDecoder::push_full(Buffer* b) { _impl->full.push(b); }
Buffer* MpegDecoder::pop_empty() { return _impl->empty.pop(); }
int64_t MpegDecoder::get_pts() { return _last_pts; }

void MpegDecoder::reset()                       // player.cpp:439
{
    _impl->full.drain_into(_impl->empty);
    _impl->es.clear(); _impl->pes.clear();
    _impl->in_picture = false; _impl->scan_from = 0;
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

// decode one access unit (headers + exactly one picture) and land it in the host Frame _current
static void decode_unit(MpegDecoder* d, ef_decoder_impl* im, const uint8_t* p, size_t n)
{
    const uint64_t off[2] = { 0, (uint64_t)n };
    int rc = ef_submit_es_host(im->ctx, p, off, nullptr);
    if (rc == EF_OK) rc = ef_index(im->ctx, nullptr);
    if (rc == EF_OK) rc = ef_decode_picture(im->ctx, 0, nullptr);
    if (rc == EF_OK) rc = ef_read_frame(im->ctx, 0, d->_fb_index & 1, im->staging.data());
    if (rc != EF_OK) { fprintf(stderr, "MpegDecoder: %s\n", ef_last_error()); return; }   // reference convention: print and keep going
    Frame* f = d->_current;
    for (int s = 0; s < FB_SLICES; s++) memcpy(f->_slices[s], im->staging.data() + (size_t)s * EF_STRIP_BYTES, EF_STRIP_BYTES);
}

// cut complete access units off the front of `es`: a unit ends where, after a picture start code,
// the next sequence / GOP / picture / sequence-end start code begins (marker(), player.cpp:1318)
static void drain(MpegDecoder* d, ef_decoder_impl* im, bool final)
{
    std::vector<uint8_t>& es = im->es;
    size_t unit_start = 0, i = im->scan_from;
    while (i + 4 <= es.size()) {
        if (es[i] != 0 || es[i + 1] != 0 || es[i + 2] != 1) { i++; continue; }
        const uint8_t code = es[i + 3];
        const bool header = code == 0x00 || code == 0xB3 || code == 0xB8 || code == 0xB7;
        if (header && im->in_picture) {
            decode_unit(d, im, es.data() + unit_start, i - unit_start);
            unit_start = i;
            im->in_picture = false;
        }
        if (code == 0x00) {                                       // picture(): present the previous one, swap, latch the PTS
            int64_t pts = d->_pts;
            for (auto& pp : im->pes) if (pp.first <= i + 3) pts = pp.second;
            if (pts != -1) d->_pts = pts;
            d->flush_picture(0);
            im->in_picture = true;
        }
        i += 4;
    }
    if (final && im->in_picture && es.size() > unit_start) {
        decode_unit(d, im, es.data() + unit_start, es.size() - unit_start);
        unit_start = es.size();
        im->in_picture = false;
        i = es.size();
    }
    // drop what has been decoded, keep PES marks relative to the new origin
    if (unit_start) {
        es.erase(es.begin(), es.begin() + (long)unit_start);
        std::vector<std::pair<size_t, int64_t>> keep;
        for (auto& pp : im->pes) if (pp.first >= unit_start) keep.push_back({ pp.first - unit_start, pp.second });
        im->pes.swap(keep);
    }
    im->scan_from = i >= unit_start ? i - unit_start : 0;   // everything before i has been examined; a start code
                                                            // straddling the next Buffer starts at or after i
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
                if (payload < end) push_audio(payload, (int)(end - payload), pts, false);
            }
        }
        im->empty.push(b);
        drain(this, im, false);
    }
}
