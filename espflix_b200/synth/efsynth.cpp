// espflix_b200/synth/efsynth.cpp — deterministic synthetic MPEG-1 I+P stream generator.
//
// Workload generator for the parity tests and bench.py (SURVEY.md §8d "Synthetic stream spec"):
// 352x192, N=12 GOPs (1 I + 11 P), no B pictures, 12 / 5 / 1 slices per picture, TS wrapper with
// PID 0x100 and one PES (PTS = 129003 + 3003 k) per picture, i.e. the wire format the reference's
// indexer asks ffmpeg for (indexer/indexer.cpp:306-309). It is an ENCODER: nothing here is on the
// decode path and nothing here comes from the reference except the stream syntax it accepts.
// The reconstruction loop mirrors the decoder arithmetic (reference quirks Q1-Q5) so that the
// motion search runs on the pictures a decoder will actually hold.
//
// Coverage flags add what the reference's own fixtures lack: macroblock-level
// quantiser changes, full_pel_forward, f_code > 1 with long vectors, 16-bit escapes, saturating
// coefficients (EFS_OVERDRIVE, outside the reference's defined domain), custom quantiser matrices (incl. entries small enough for quirk Q2), intra
// macroblocks in P pictures and skipped runs at row ends.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../csrc/ef_iso11172_tables.h"

namespace {

enum {
    EFS_MBQUANT = 1,       // change quantizer_scale at macroblock level (types 0x11/0x12/0x1A)
    EFS_FULLPEL = 2,       // full_pel_forward_vector = 1
    EFS_FCODE3 = 4,        // forward_f_code = 3, longer vectors
    EFS_BIGLEVELS = 8,     // qscale 1 + boosted contrast -> 16-bit escapes and saturation
    EFS_MATRICES = 16,     // load custom intra / non-intra matrices (Q4) with small entries (Q2)
    EFS_INTRA_IN_P = 32,   // force some intra macroblocks inside P pictures
    EFS_STATIC = 64,       // mostly static scene -> long skipped runs
    EFS_OVERDRIVE = 128,   // force +-255 levels at the two top frequencies of some intra blocks: coefficient saturation at
                           // +-2048. One saturated coefficient alone swings samples by +-400..500, i.e. OUTSIDE [-256,511], the
                           // domain of the reference's clamp table (it reads out of bounds there) - decoder-vs-oracle tests only.
};

const int W = 352, H = 192, MBW = 22, MBH = 12;

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 6364136223846793005ULL + 1442695040888963407ULL) { next(); next(); }
    uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
    int range(int lo, int hi) { return lo + (int)(next() % (uint32_t)(hi - lo + 1)); }
    float unit() { return (next() & 0xFFFFFF) / 16777216.0f; }
};

struct BitWriter {
    std::vector<uint8_t>& out;
    uint32_t acc = 0; int n = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(uint32_t v, int bits) {
        for (int i = bits - 1; i >= 0; i--) {
            acc = (acc << 1) | ((v >> i) & 1);
            if (++n == 8) { out.push_back((uint8_t)acc); acc = 0; n = 0; }
        }
    }
    void code(const char* c) { for (; *c; c++) put(*c == '1', 1); }
    void align() { while (n) put(0, 1); }
    void start_code(int c) { align(); out.push_back(0); out.push_back(0); out.push_back(1); out.push_back((uint8_t)c); }
};

const char* lookup(const ef_vlc_code* t, int n, int value)
{
    for (int i = 0; i < n; i++) if (t[i].value == value) return t[i].code;
    return nullptr;
}

struct Plane { int w, h; std::vector<uint8_t> p; uint8_t& at(int x, int y) { return p[(size_t)y * w + x]; } uint8_t at(int x, int y) const { return p[(size_t)y * w + x]; } };
struct Picture { Plane y, cb, cr; Picture() { y = {W, H, std::vector<uint8_t>(W * H)}; cb = {W / 2, H / 2, std::vector<uint8_t>(W * H / 4)}; cr = cb; } };

// smooth random "world" larger than the picture; pictures are moving windows onto it
struct World {
    int w, h; std::vector<float> y, cb, cr;
    static void smooth(std::vector<float>& dst, int w, int h, Rng& r, int cell, float lo, float hi)
    {
        int gw = w / cell + 3, gh = h / cell + 3;
        std::vector<float> g((size_t)gw * gh);
        for (auto& v : g) v = lo + (hi - lo) * r.unit();
        dst.resize((size_t)w * h);
        for (int yy = 0; yy < h; yy++)
            for (int xx = 0; xx < w; xx++) {
                float fx = (float)xx / cell, fy = (float)yy / cell;
                int ix = (int)fx, iy = (int)fy; fx -= ix; fy -= iy;
                fx = fx * fx * (3 - 2 * fx); fy = fy * fy * (3 - 2 * fy);
                float a = g[(size_t)iy * gw + ix], b = g[(size_t)iy * gw + ix + 1];
                float c = g[(size_t)(iy + 1) * gw + ix], d = g[(size_t)(iy + 1) * gw + ix + 1];
                dst[(size_t)yy * w + xx] = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy;
            }
    }
    World(Rng& r, float contrast)
    {
        w = W + 160; h = H + 160;
        smooth(y, w, h, r, 8, 128 - 110 * contrast, 128 + 107 * contrast);
        std::vector<float> fine; smooth(fine, w, h, r, 3, -18 * contrast, 18 * contrast);
        for (size_t i = 0; i < y.size(); i++) y[i] += fine[i];
        smooth(cb, w / 2, h / 2, r, 8, 128 - 100 * contrast, 128 + 100 * contrast);
        smooth(cr, w / 2, h / 2, r, 8, 128 - 100 * contrast, 128 + 100 * contrast);
    }
    static float samp(const std::vector<float>& p, int w, int h, float x, float y)
    {
        x = x < 0 ? 0 : x > w - 2 ? (float)(w - 2) : x;
        y = y < 0 ? 0 : y > h - 2 ? (float)(h - 2) : y;
        int ix = (int)x, iy = (int)y; float fx = x - ix, fy = y - iy;
        const float* q = &p[(size_t)iy * w + ix];
        return (q[0] * (1 - fx) + q[1] * fx) * (1 - fy) + (q[w] * (1 - fx) + q[w + 1] * fx) * fy;
    }
    void render(Picture& pic, float ox, float oy, float px, float py, int lo, int hi, int noise, Rng& nr) const
    {
        for (int yy = 0; yy < H; yy++)
            for (int xx = 0; xx < W; xx++) {
                float sx = xx + ox, sy = yy + oy;
                if (xx >= px && xx < px + 64 && yy >= py && yy < py + 64) { sx = xx - px + 20.0f; sy = yy - py + 200.0f; }   // moving patch
                int v = (int)lrintf(samp(y, w, h, sx, sy));
                if (noise) v += nr.range(-noise, noise);       // temporal grain: keeps P residuals non-trivial
                pic.y.at(xx, yy) = (uint8_t)(v < lo ? lo : v > hi ? hi : v);
            }
        for (int yy = 0; yy < H / 2; yy++)
            for (int xx = 0; xx < W / 2; xx++) {
                float sx = xx + ox / 2, sy = yy + oy / 2;
                int a = (int)lrintf(samp(cb, w / 2, h / 2, sx, sy)), b = (int)lrintf(samp(cr, w / 2, h / 2, sx, sy));
                pic.cb.at(xx, yy) = (uint8_t)(a < 16 ? 16 : a > 240 ? 240 : a);
                pic.cr.at(xx, yy) = (uint8_t)(b < 16 ? 16 : b > 240 ? 240 : b);
            }
    }
};

// ---- transform -------------------------------------------------------------------------
float g_cos[8][8];
void init_dct() { for (int k = 0; k < 8; k++) for (int n = 0; n < 8; n++) g_cos[k][n] = (k ? 0.5f : 0.35355339f) * cosf((2 * n + 1) * k * (float)M_PI / 16); }
void fdct(const int* px, float* out)
{
    float tmp[64];
    for (int r = 0; r < 8; r++) for (int k = 0; k < 8; k++) { float s = 0; for (int n = 0; n < 8; n++) s += g_cos[k][n] * px[r * 8 + n]; tmp[r * 8 + k] = s; }
    for (int c = 0; c < 8; c++) for (int k = 0; k < 8; k++) { float s = 0; for (int n = 0; n < 8; n++) s += g_cos[k][n] * tmp[n * 8 + c]; out[k * 8 + c] = s; }
}

// decoder-side reconstruction of one 1-D pass (same arithmetic a decoder of this syntax applies)
void idct_pass(int* v, int stride, bool fin)
{
    int i0 = v[0], i1 = v[stride], i2 = v[2 * stride], i3 = v[3 * stride], i4 = v[4 * stride], i5 = v[5 * stride], i6 = v[6 * stride], i7 = v[7 * stride];
    int b3 = i2 + i6, b4 = i5 - i3, t1 = i1 + i7, t2 = i3 + i5, b6 = i1 - i7, b7 = t1 + t2;
    int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    int x0 = x4 - (((t1 - t2) * 362 + 128) >> 8);
    int x1 = i0 - i4, x2 = (((i2 - i6) * 362 + 128) >> 8) - b3, x3 = i0 + i4;
    int y3 = x1 + x2, y4 = x3 + b3, y5 = x1 - x2, y6 = x3 - b3, y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
    int o[8] = { b7 + y4, x4 + y3, y5 - x0, y6 - y7, y6 + y7, x0 + y5, y3 - x4, y4 - b7 };
    for (int i = 0; i < 8; i++) v[i * stride] = fin ? (o[i] + 128) >> 8 : o[i];
}
int clamp248(int v) { return v < 0 ? 0 : v > 248 ? 248 : v; }

struct Encoder {
    Rng rng; int flags; int slices; int base_q;
    std::vector<uint8_t>& es; BitWriter bw;
    uint8_t intra_q[64], inter_q[64];      // as the DECODER will index them (raster index into stream-order bytes, Q4)
    Picture recon[2]; int cur = 1;          // ping-pong like the decoder
    int ptype = 1, full_pel = 0, f_code = 1;
    int qscale = 6;
    // slice state
    int dc_pred[3], mv_pred[2];

    Encoder(uint64_t seed, int fl, int sl, int q, std::vector<uint8_t>& out) : rng(seed ^ 0x9E3779B97F4A7C15ULL), flags(fl), slices(sl), base_q(q), es(out), bw(out)
    {
        memcpy(intra_q, ef_default_intra_q, 64); memset(inter_q, 16, 64);
    }

    void sequence_header()
    {
        bw.start_code(0xB3);
        bw.put(W, 12); bw.put(H, 12); bw.put(1, 4); bw.put(4, 4); bw.put(0x3FFFF, 18); bw.put(1, 1); bw.put(20, 10); bw.put(0, 1);
        if (flags & EFS_MATRICES) {
            uint8_t mi[64], mn[64];
            for (int i = 0; i < 64; i++) { mi[i] = (uint8_t)(i == 0 ? 8 : rng.range(1, 48)); mn[i] = (uint8_t)rng.range(1, 40); }
            for (int i = 0; i < 6; i++) { mi[rng.range(1, 63)] = 1; mn[rng.range(0, 63)] = 1; }     // tiny entries -> Q2 (v == 0 -> +1)
            bw.put(1, 1); for (int i = 0; i < 64; i++) bw.put(mi[i], 8);
            bw.put(1, 1); for (int i = 0; i < 64; i++) bw.put(mn[i], 8);
            memcpy(intra_q, mi, 64); memcpy(inter_q, mn, 64);    // decoder indexes the stream-order bytes with the raster index
        } else { bw.put(0, 1); bw.put(0, 1); }
        bw.align();
    }
    void gop_header(int k) { bw.start_code(0xB8); bw.put((1u << 12) | (uint32_t)(k & 63), 25); bw.put(1, 1); bw.put(0, 1); bw.put(0, 5); bw.align(); }
    void picture_header(int tref, int type)
    {
        ptype = type;
        bw.start_code(0x00);
        bw.put(tref & 1023, 10); bw.put(type, 3); bw.put(0xFFFF, 16);
        if (type == 2) { bw.put(full_pel, 1); bw.put(f_code, 3); }
        bw.put(0, 1); bw.align();
    }

    // ---- block coding ----
    void put_dc(int blk, int diff)
    {
        int a = abs(diff), size = 0;
        while (a >> size) size++;
        bw.code(lookup(blk < 4 ? ef_vlc_dc_luma : ef_vlc_dc_chroma, 9, size));
        if (size) bw.put((uint32_t)(diff > 0 ? diff : diff + (1 << size) - 1), size);
    }
    void put_coef(int run, int level, bool first)
    {
        int a = abs(level);
        const char* c = (a <= 40 && run <= 31) ? lookup(ef_vlc_dct, EF_VLC_DCT_COUNT, (run << 8) | a) : nullptr;
        if (run == 0 && a == 1) { bw.code(first ? "1" : "11"); bw.put(level < 0, 1); return; }
        if (c) { bw.code(c); bw.put(level < 0, 1); return; }
        bw.code("000001"); bw.put(run, 6);
        if (level >= -127 && level <= 127) bw.put((uint32_t)level & 0xFF, 8);
        else if (level > 0) { bw.put(0, 8); bw.put(level, 8); }
        else { bw.put(128, 8); bw.put((uint32_t)(level + 256) & 0xFF, 8); }
    }

    // quantise + emit one block, reconstruct exactly as a decoder would; returns false if block has no coefficients (non-intra)
    // px: source (intra) or residual (inter); pred: prediction (inter) ; out: 64 reconstructed pixels
    struct Coded { int levels[64]; int n; };       // zig-zag order levels
    bool quantise(const int* px, bool intra, Coded& c)
    {
        float f[64]; fdct(px, f);
        const uint8_t* q = intra ? intra_q : inter_q;
        bool any = false;
        const bool boost = (flags & EFS_OVERDRIVE) && rng.range(0, 7) == 0;
        for (int i = 0; i < 64; i++) {
            int zz = ef_zigzag[i];
            int lv;
            if (intra && i == 0) lv = (int)lrintf(f[0] / 8.0f);
            else {
                float s = f[zz] * 8.0f / (float)(qscale * q[zz]);
                lv = intra ? (int)lrintf(s) : (int)s;        // dead zone for residuals
            }
            if (boost && intra && i >= 62) lv = (i & 1) ? 255 : -255;     // EFS_OVERDRIVE
            lv = lv > 255 ? 255 : lv < -255 ? -255 : lv;
            c.levels[i] = lv;
            if (lv && !(intra && i == 0)) any = true;
        }
        return any || intra;
    }
    void reconstruct(const Coded& c, bool intra, int dc_value, const int* pred, uint8_t* out)
    {
        int b[64]; memset(b, 0, sizeof(b));
        const uint8_t* q = intra ? intra_q : inter_q;
        int n = 0;
        if (intra) { b[0] = dc_value << 8; n = 1; }
        int last = 0;
        for (int i = intra ? 1 : 0; i < 64; i++) {
            if (!c.levels[i]) continue;
            int zz = ef_zigzag[i], v = c.levels[i] << 1;
            if (!intra) v += v < 0 ? -1 : 1;
            v = (v * qscale * q[zz]) / 16;
            if ((v & 1) == 0) v -= v > 0 ? 1 : -1;
            if (v > 2047) v = 2047; else if (v < -2048) v = -2048;
            b[zz] = v * ef_aan_prescale[zz];
            last = i + 1;
        }
        if (last > n) n = last;
        if (n == 1) {
            int dc = b[0] >> 8;
            for (int i = 0; i < 64; i++) out[i] = intra ? (uint8_t)dc : (uint8_t)clamp248(dc + pred[i]);
            return;
        }
        for (int i = 0; i < 8; i++) idct_pass(b + i, 8, false);
        for (int i = 0; i < 64; i += 8) idct_pass(b + i, 1, true);
        for (int i = 0; i < 64; i++) out[i] = (uint8_t)clamp248(b[i] + (intra ? 0 : pred[i]));
    }
    void emit_block(const Coded& c, bool intra, int blk, int& dc_value)
    {
        int start = 0;
        if (intra) {
            int comp = blk < 4 ? 0 : blk - 3;
            int dc = c.levels[0] < 0 ? 0 : c.levels[0] > 255 ? 255 : c.levels[0];
            int diff = dc - dc_pred[comp];
            diff = diff > 255 ? 255 : diff < -255 ? -255 : diff;
            put_dc(blk, diff);
            dc_pred[comp] += diff; dc_value = dc_pred[comp];
            start = 1;
        }
        int run = 0; bool first = !intra;
        for (int i = start; i < 64; i++) {
            if (!c.levels[i]) { run++; continue; }
            put_coef(run, c.levels[i], first);
            first = false; run = 0;
        }
        bw.code("10");
    }

    // ---- prediction (decoder arithmetic) ----
    static void mc_block(const Plane& ref, int hx, int hy, int size, int* out)   // half-pel position of the top-left sample
    {
        int xh = hx & 1, yh = hy & 1, x0 = hx >> 1, y0 = hy >> 1;
        for (int y = 0; y < size; y++)
            for (int x = 0; x < size; x++) {
                int a = ref.at(x0 + x, y0 + y);
                if (xh && yh) a = (a + ref.at(x0 + x + 1, y0 + y) + ref.at(x0 + x, y0 + y + 1) + ref.at(x0 + x + 1, y0 + y + 1) + 2) >> 2;
                else if (xh) a = (a + ref.at(x0 + x + 1, y0 + y) + 1) >> 1;
                else if (yh) a = (a + ref.at(x0 + x, y0 + y + 1) + 1) >> 1;
                out[y * size + x] = a;
            }
    }
    static int sad16(const Plane& src, int sx, int sy, const int* pred)
    {
        int s = 0;
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) s += abs((int)src.at(sx + x, sy + y) - pred[y * 16 + x]);
        return s;
    }
    // motion search in half-pel units (or full-pel units x2 when full_pel), vectors kept in frame and in f_code range
    void search(const Picture& src, const Picture& ref, int mbx, int mby, int& bh, int& bv, int range_px)
    {
        int step = full_pel ? 2 : 1;
        int limit = (16 << (f_code - 1)) * (full_pel ? 2 : 1);    // decoder range in half-pel units after scaling
        int pred[256];
        auto legal = [&](int h, int v) {
            int hx = mbx * 32 + h, hy = mby * 32 + v;
            if (hx < 0 || hy < 0) return false;
            if ((hx >> 1) + 16 + (hx & 1) > W || (hy >> 1) + 16 + (hy & 1) > H) return false;
            return h >= -limit && h <= limit - step && v >= -limit && v <= limit - step;
        };
        bh = bv = 0; mc_block(ref.y, mbx * 32, mby * 32, 16, pred);
        int best = sad16(src.y, mbx * 16, mby * 16, pred) - 64;
        for (int v = -range_px * 2; v <= range_px * 2; v += 2)
            for (int h = -range_px * 2; h <= range_px * 2; h += 2) {
                if (!legal(h, v)) continue;
                mc_block(ref.y, mbx * 32 + h, mby * 32 + v, 16, pred);
                int s = sad16(src.y, mbx * 16, mby * 16, pred);
                if (s < best) { best = s; bh = h; bv = v; }
            }
        if (!full_pel) {
            int ch = bh, cv = bv;
            for (int v = cv - 1; v <= cv + 1; v++)
                for (int h = ch - 1; h <= ch + 1; h++) {
                    if (!legal(h, v)) continue;
                    mc_block(ref.y, mbx * 32 + h, mby * 32 + v, 16, pred);
                    int s = sad16(src.y, mbx * 16, mby * 16, pred);
                    if (s < best) { best = s; bh = h; bv = v; }
                }
        }
    }
    void put_mv_component(int delta_units, int& pred_units)     // units: half-pel (or full-pel if full_pel)
    {
        int r = f_code - 1, f = 1 << r, range = 16 * f;
        int d = delta_units;
        if (d < -range) d += 2 * range; else if (d > range - 1) d -= 2 * range;
        if (d == 0) bw.code(lookup(ef_vlc_mv, EF_VLC_MV_COUNT, 0));
        else {
            int a = abs(d) - 1, code = (a >> r) + 1, res = a & (f - 1);
            bw.code(lookup(ef_vlc_mv, EF_VLC_MV_COUNT, d < 0 ? -code : code));
            if (r) bw.put(res, r);
        }
        pred_units += d;
        if (pred_units > range - 1) pred_units -= 2 * range; else if (pred_units < -range) pred_units += 2 * range;
    }

    void put_increment(int inc)
    {
        while (inc > 33) { bw.code(lookup(ef_vlc_mba, EF_VLC_MBA_COUNT, 35)); inc -= 33; }
        bw.code(lookup(ef_vlc_mba, EF_VLC_MBA_COUNT, inc));
    }

    void gather(const Plane& p, int x0, int y0, int* out) { for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) out[y * 8 + x] = p.at(x0 + x, y0 + y); }
    void scatter(Plane& p, int x0, int y0, const uint8_t* in) { for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) p.at(x0 + x, y0 + y) = in[y * 8 + x]; }

    void encode_picture(const Picture& src, int type)
    {
        Picture& out = recon[cur];
        const Picture& ref = recon[cur ^ 1];
        std::vector<int> first_rows;
        if (slices == 12) for (int r = 0; r < 12; r++) first_rows.push_back(r);
        else if (slices == 5) first_rows = {0, 2, 5, 7, 10};
        else first_rows = {0};
        int search_px = (flags & EFS_FCODE3) ? 12 : 7;
        for (size_t si = 0; si < first_rows.size(); si++) {
            int row0 = first_rows[si], row1 = si + 1 < first_rows.size() ? first_rows[si + 1] : MBH;
            qscale = base_q;
            if (flags & EFS_BIGLEVELS) qscale = 1;
            bw.start_code(row0 + 1);
            bw.put(qscale, 5); bw.put(0, 1);
            dc_pred[0] = dc_pred[1] = dc_pred[2] = 128; mv_pred[0] = mv_pred[1] = 0;
            int pending = 0;                      // skipped macroblocks not yet signalled
            int total = (row1 - row0) * MBW;
            for (int idx = 0; idx < total; idx++) {
                int mbx = idx % MBW, mby = row0 + idx / MBW;
                bool firstmb = idx == 0, lastmb = idx == total - 1;
                int want_q = qscale;
                if ((flags & EFS_MBQUANT) && rng.range(0, 5) == 0) want_q = rng.range(1, 31);
                if ((flags & EFS_BIGLEVELS)) want_q = 1;

                bool intra = type == 1;
                int mh = 0, mv = 0;                // vector in coded units
                int predY[256], predC[2][64];
                if (type == 2) {
                    if ((flags & EFS_INTRA_IN_P) && rng.range(0, 9) == 0) intra = true;
                    else {
                        int bh, bv; search(src, ref, mbx, mby, bh, bv, search_px);
                        mh = full_pel ? bh / 2 : bh; mv = full_pel ? bv / 2 : bv;
                        int hx = mbx * 32 + bh, hy = mby * 32 + bv;
                        mc_block(ref.y, hx, hy, 16, predY);
                        mc_block(ref.cb, hx >> 1, hy >> 1, 8, predC[0]);      // floor (decoder quirk Q3)
                        mc_block(ref.cr, hx >> 1, hy >> 1, 8, predC[1]);
                    }
                }
                // quantise all six blocks
                int saved_q = qscale; qscale = want_q;
                Coded cb6[6]; bool coded[6]; int cbp = 0;
                for (int b = 0; b < 6; b++) {
                    int px[64];
                    const Plane& sp = b < 4 ? src.y : b == 4 ? src.cb : src.cr;
                    int x0 = b < 4 ? mbx * 16 + (b & 1) * 8 : mbx * 8, y0 = b < 4 ? mby * 16 + (b >> 1) * 8 : mby * 8;
                    gather(sp, x0, y0, px);
                    if (!intra) {
                        const int* pr = b < 4 ? predY : predC[b - 4];
                        int stride = b < 4 ? 16 : 8, ox = b < 4 ? (b & 1) * 8 : 0, oy = b < 4 ? (b >> 1) * 8 : 0;
                        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) px[y * 8 + x] -= pr[(oy + y) * stride + ox + x];
                    }
                    coded[b] = quantise(px, intra, cb6[b]);
                    if (coded[b]) cbp |= 0x20 >> b;
                }
                bool has_mv = !intra && (mh || mv);
                bool skippable = type == 2 && !intra && !cbp && !has_mv && !firstmb && !lastmb;
                if (skippable) {
                    qscale = saved_q;
                    pending++;
                    // decoder: predict_zero + predictor reset
                    for (int b = 0; b < 6; b++) {
                        Plane& dp = b < 4 ? out.y : b == 4 ? out.cb : out.cr; const Plane& rp = b < 4 ? ref.y : b == 4 ? ref.cb : ref.cr;
                        int x0 = b < 4 ? mbx * 16 + (b & 1) * 8 : mbx * 8, y0 = b < 4 ? mby * 16 + (b >> 1) * 8 : mby * 8;
                        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) dp.at(x0 + x, y0 + y) = rp.at(x0 + x, y0 + y);
                    }
                    continue;
                }
                if (pending) { dc_pred[0] = dc_pred[1] = dc_pred[2] = 128; mv_pred[0] = mv_pred[1] = 0; }
                put_increment(pending + 1); pending = 0;
                bool quant = want_q != saved_q && (intra || cbp);
                if (!quant) qscale = saved_q;      // cannot signal a change without coefficients: requantise below
                if (!quant && want_q != saved_q) {  // redo quantisation with the old scale
                    cbp = 0;
                    for (int b = 0; b < 6; b++) {
                        int px[64];
                        const Plane& sp = b < 4 ? src.y : b == 4 ? src.cb : src.cr;
                        int x0 = b < 4 ? mbx * 16 + (b & 1) * 8 : mbx * 8, y0 = b < 4 ? mby * 16 + (b >> 1) * 8 : mby * 8;
                        gather(sp, x0, y0, px);
                        if (!intra) {
                            const int* pr = b < 4 ? predY : predC[b - 4];
                            int stride = b < 4 ? 16 : 8, ox = b < 4 ? (b & 1) * 8 : 0, oy = b < 4 ? (b >> 1) * 8 : 0;
                            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) px[y * 8 + x] -= pr[(oy + y) * stride + ox + x];
                        }
                        coded[b] = quantise(px, intra, cb6[b]);
                        if (coded[b]) cbp |= 0x20 >> b;
                    }
                }
                int mbtype;
                if (intra) mbtype = quant ? 0x11 : 0x01;
                else if (cbp) mbtype = (has_mv || rng.range(0, 3) == 0 ? 0x08 : 0) | 0x02 | (quant ? 0x10 : 0);
                else mbtype = 0x08;
                if (!intra && (mbtype & 0x08)) has_mv = true;   // vector is transmitted (possibly zero)
                bw.code(lookup(type == 1 ? ef_vlc_mbtype_i : ef_vlc_mbtype_p, type == 1 ? 2 : 7, mbtype));
                if (mbtype & 0x10) bw.put(qscale, 5);
                if (intra) { mv_pred[0] = mv_pred[1] = 0; }
                else {
                    dc_pred[0] = dc_pred[1] = dc_pred[2] = 128;
                    if (mbtype & 0x08) { put_mv_component(mh - mv_pred[0], mv_pred[0]); put_mv_component(mv - mv_pred[1], mv_pred[1]); }
                    else mv_pred[0] = mv_pred[1] = 0;
                }
                if (mbtype & 0x02) bw.code(lookup(ef_vlc_cbp, EF_VLC_CBP_COUNT, cbp));
                for (int b = 0; b < 6; b++) {
                    Plane& dp = b < 4 ? out.y : b == 4 ? out.cb : out.cr;
                    int x0 = b < 4 ? mbx * 16 + (b & 1) * 8 : mbx * 8, y0 = b < 4 ? mby * 16 + (b >> 1) * 8 : mby * 8;
                    int pr[64]; uint8_t rec[64];
                    if (!intra) {
                        const int* p = b < 4 ? predY : predC[b - 4];
                        int stride = b < 4 ? 16 : 8, ox = b < 4 ? (b & 1) * 8 : 0, oy = b < 4 ? (b >> 1) * 8 : 0;
                        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pr[y * 8 + x] = p[(oy + y) * stride + ox + x];
                    }
                    if (cbp & (0x20 >> b)) {
                        int dcv = 0;
                        emit_block(cb6[b], intra, b, dcv);
                        reconstruct(cb6[b], intra, dcv, pr, rec);
                    } else for (int i = 0; i < 64; i++) rec[i] = (uint8_t)pr[i];
                    scatter(dp, x0, y0, rec);
                }
            }
            bw.align();
        }
        cur ^= 1;
    }
};

}  // namespace

extern "C" {

// Generate one stream: n_pictures pictures in GOPs of `gop` (I then P...). slices in {12,5,1}.
// noise: amplitude of per-picture luma grain (0 = none) used to set the P-picture bit budget.
// Returns ES bytes written (0 if es_cap too small). pic_off (n_pictures+1 entries, optional): byte
// offset in the ES where each picture's access unit starts (sequence/GOP headers included).
size_t efs_generate(uint64_t seed, int n_pictures, int gop, int slices, int qscale, int flags, int noise,
                    uint8_t* es_out, size_t es_cap, uint32_t* pic_off)
{
    static bool once = false;
    if (!once) { init_dct(); once = true; }
    std::vector<uint8_t> es;
    es.reserve((size_t)n_pictures * 16384);
    Encoder enc(seed, flags, slices, qscale, es);
    Rng rng(seed);
    float contrast = (flags & EFS_BIGLEVELS) ? 1.6f : 0.55f;
    World world(rng, contrast);
    float vx = (flags & EFS_STATIC) ? 0.0f : (flags & EFS_FCODE3) ? 5.5f : 1.5f + (seed % 3) * 0.5f;
    float vy = (flags & EFS_STATIC) ? 0.0f : (flags & EFS_FCODE3) ? -3.5f : 0.5f + (seed % 2) * 1.0f;
    if (flags & EFS_FULLPEL) { enc.full_pel = 1; vx = 2.0f; vy = 1.0f; }
    if (flags & EFS_FCODE3) enc.f_code = 3;
    Picture src;
    for (int k = 0; k < n_pictures; k++) {
        if (pic_off) pic_off[k] = (uint32_t)es.size();
        bool is_i = (k % gop) == 0;
        if (is_i) { enc.sequence_header(); enc.gop_header(k); }
        float ox = 60 + vx * k, oy = 60 + vy * k;
        float px = 40 + 0.5f * k * 3, py = 30 + 0.5f * k;           // patch moves at half-pel multiples
        if (flags & EFS_STATIC) { px = 40 + (k & 1) * 0.5f; py = 30; }
        int hi = (flags & EFS_BIGLEVELS) ? 255 : 235, lo = (flags & EFS_BIGLEVELS) ? 0 : 16;
        world.render(src, ox, oy, px, py, lo, hi, noise, rng);
        enc.picture_header(k % gop, is_i ? 1 : 2);
        enc.encode_picture(src, is_i ? 1 : 2);
    }
    enc.bw.align();
    if (pic_off) pic_off[n_pictures] = (uint32_t)es.size();
    if (es.size() > es_cap) return 0;
    memcpy(es_out, es.data(), es.size());
    return es.size();
}

// Wrap an ES into 188-byte TS packets, PID 0x100, one PES (flags 0x8080, PTS = 129003 + 3003 k)
// per picture. Returns TS bytes (0 if ts_cap too small).
size_t efs_wrap_ts(const uint8_t* es, const uint32_t* pic_off, int n_pictures, uint8_t* ts_out, size_t ts_cap)
{
    std::vector<uint8_t> ts;
    int cc = 0;
    for (int k = 0; k < n_pictures; k++) {
        std::vector<uint8_t> pes;
        int64_t pts = 129003 + 3003LL * k;
        const uint8_t hdr[] = { 0, 0, 1, 0xE0, 0, 0, 0x80, 0x80, 5,
            (uint8_t)(0x21 | ((pts >> 29) & 0x0E)), (uint8_t)(pts >> 22), (uint8_t)(0x01 | ((pts >> 14) & 0xFE)), (uint8_t)(pts >> 7), (uint8_t)(0x01 | ((pts << 1) & 0xFE)) };
        pes.insert(pes.end(), hdr, hdr + sizeof(hdr));
        pes.insert(pes.end(), es + pic_off[k], es + pic_off[k + 1]);
        size_t pos = 0; bool first = true;
        while (pos < pes.size()) {
            size_t left = pes.size() - pos;
            uint8_t pkt[188];
            pkt[0] = 0x47; pkt[1] = (uint8_t)((first ? 0x40 : 0) | 0x01); pkt[2] = 0x00;
            size_t payload = 184;
            if (left >= 184) { pkt[3] = (uint8_t)(0x10 | (cc & 15)); memcpy(pkt + 4, &pes[pos], 184); }
            else {
                size_t af = 184 - left;            // adaptation field incl. its length byte
                pkt[3] = (uint8_t)(0x30 | (cc & 15));
                pkt[4] = (uint8_t)(af - 1);
                if (af > 1) { pkt[5] = 0; memset(pkt + 6, 0xFF, af - 2); }
                memcpy(pkt + 4 + af, &pes[pos], left);
                payload = left;
            }
            cc++; first = false; pos += payload;
            ts.insert(ts.end(), pkt, pkt + 188);
        }
    }
    if (ts.size() > ts_cap) return 0;
    memcpy(ts_out, ts.data(), ts.size());
    return ts.size();
}

}  // extern "C"
