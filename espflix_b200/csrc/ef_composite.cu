// espflix_b200/csrc/ef_composite.cu — K2, the composite field kernel: one launch synthesises a whole
// NTSC (262 x 912) or PAL (312 x 1136) field of uint16 DAC samples for every stream.
//
// Replaces one field's worth of video_isr() calls (video.cpp:1122-1198): sync(), burst() /
// burst_pal(), blit() (video.cpp:690-804), blanking(), pal_sync(). The reference rewrites only part
// of a two-entry ping-pong line buffer on active lines; the closed form of what the DAC sees is
// (SURVEY.md §8a a19):
//   NTSC active line: sync [0,64) | burst [64,104) | BLACK [104,160) | blit [160,864) | BLACK [864,912)
//   PAL  active line: sync [0,80) | BLACK [80,96) | burst [96,140) | BLACK [140,280) | blit [280,984) | BLACK [984,1136)
// Pure streaming, HBM bound: each thread produces 8 samples (one 16-byte store) from 4 luma pixels
// and 2+2 chroma bytes; the packed 4x8-bit expressions of blit() are reproduced verbatim because
// the low byte of every sample carries deterministic carry "junk" that parity depends on.
#include "ef_common.cuh"

namespace {

__device__ __forceinline__ uint32_t chroma_word(const uint32_t* tab, uint32_t u, uint32_t v, int vt)
{
    return ((tab[u & 255] + tab[vt + (v & 255)]) & 0xFCFCFCFCu) >> 2;      // CHROMA_EVEN / CHROMA_ODD, video.cpp:670
}

// The chroma LUTs have a closed form (tests/test_tables.py::test_chroma_lut_closed_form checks all 768 entries of
// both standards): the four subcarrier phases of a table entry are 48, 48 +- r(c) with
//   r(c) = sgn(128 - c) * ((16 |128 - c| + 11) / 22)          [= round_half_away((128 - c) * 24 / 33), espflix.cpp:1091-1180]
// clamped to [0, 127]: U rides on the sine phases (bytes 1, 0), V on the cosine phases (bytes 3, 2; swapped on the PAL
// lines that use cos_v_neg). EF_K2_ARITH computes CHROMA_EVEN / CHROMA_ODD from that instead of two dependent gathers.
#ifndef EF_K2_ARITH
#define EF_K2_ARITH 0
#endif
__device__ __forceinline__ int chroma_r(uint32_t c)
{
    const int d = 128 - (int)(c & 255u);
    const int q = ((16 * abs(d) + 11) * 745) >> 14;                       // / 22, exact for numerators <= 2059 (tests/test_tables.py)
    return d < 0 ? -q : q;
}
__device__ __forceinline__ uint32_t chroma_word_arith(uint32_t u, uint32_t v, bool v_neg)
{
    const int ru = chroma_r(u);
    int rv = chroma_r(v);
    if (v_neg) rv = -rv;
    const uint32_t b3 = (uint32_t)__viaddmin_s32_relu(48, rv, 127), b2 = (uint32_t)__viaddmin_s32_relu(48, -rv, 127);
    const uint32_t b1 = (uint32_t)__viaddmin_s32_relu(48, ru, 127), b0 = (uint32_t)__viaddmin_s32_relu(48, -ru, 127);
    const uint32_t w = (b3 << 24) + (b2 << 16) + (b1 << 8) + b0 + 0x30303030u;       // + the constant phase of the other table
    return (w & 0xFCFCFCFCu) >> 2;
}

// compile-time geometry of the two standards (video_init / pal_init, video.cpp:572-630; values
// probe-verified against the reference in tests/golden/composite_pins.json)
template <bool kNtsc> struct Geo;
template <> struct Geo<true> {
    static constexpr int W = 912, LINES = 262, HSYNC = 64, HSYNC_LONG = 840, HSYNC_SHORT = 0, BURST_START = 64, BURST_W = 40,
                         TOP = 32, VSYNC = 259, BLIT = 160;
};
template <> struct Geo<false> {
    static constexpr int W = 1136, LINES = 312, HSYNC = 80, HSYNC_LONG = 536, HSYNC_SHORT = 32, BURST_START = 96, BURST_W = 44,
                         TOP = 64, VSYNC = 304, BLIT = 280;
};

template <bool kNtsc>
__device__ __forceinline__ uint32_t blank_sample(const int16_t* pal_burst, int line, int x)
{
    using G = Geo<kNtsc>;
    const uint32_t SYNC = 0x0000, BLANKING = 0x1400, BLACK = 0x1800;       // IRE(-40), IRE(0), IRE(7.5): video.cpp:520-525
    if (line >= G::VSYNC) {
        if (kNtsc) return x < G::HSYNC_LONG ? SYNC : BLANKING;             // blanking(buf, true)
        const uint32_t types = 0x00233000u;                                // _sync_type[8] = {0,0,0,3,3,2,0,0}, one nibble each
        const int t = (types >> ((line - G::VSYNC) * 4)) & 15;
        const int half = G::W >> 1;
        const int second = x >= half;
        const int xx = second ? x - half : x;
        const int lng = second ? (t & 1) : (t & 2);
        return xx < (lng ? G::HSYNC_LONG : G::HSYNC_SHORT) ? SYNC : BLANKING; // pal_sync2
    }
    if (x < G::HSYNC) return SYNC;
    const int i = x - G::BURST_START;
    if (i >= 0 && i < G::BURST_W) {
        if (kNtsc) return (i & 1) ? BLANKING : ((i & 2) ? 0x0A00u : 0x1E00u);   // burst(), video.cpp:806: 10 cycles of [1E00 1400 0A00 1400]
        return (uint16_t)pal_burst[(((line + 1) & 1) ? 0 : 64) + (i ^ 1)];     // burst_pal(): pair-swapped, table chosen by (_line_counter after ++) & 1
    }
    return BLACK;
}

__constant__ uint32_t c_dither[8] = {                                     // dither4x4, video.cpp:673
    0x00020301, 0x03010002, 0x02030100, 0x01000203, 0x03010002, 0x00020301, 0x01000203, 0x02030100 };

}  // namespace

// eight samples (one 16-byte chunk) of a blank / vsync line, or of an active line outside the blit span
// (which still shows what blanking() last left in the ping-pong buffer: sync, burst, BLACK). Most chunks lie
// wholly inside one constant region: decide per chunk, go per sample only at the burst, on the vsync lines
// and under the overlay.
template <bool kNtsc>
__device__ __forceinline__ uint4 blank_chunk(const EfDev& D, const EfPresent& pr, int line, int x0)
{
    using G = Geo<kNtsc>;
    const uint32_t SYNC2 = 0x00000000u, BLACK2 = 0x18001800u;
    uint32_t w[4];
    const int ol = line - (G::TOP + EF_H + 2);                             // overlay line 0..15 (video.cpp:1183-1189)
    if (pr.blend != 0 && ol >= 0 && ol < 16 && x0 >= G::BLIT + 16 && x0 < G::BLIT + 16 + 160 + 16 + 480) {
        // composite(), video.cpp:845-887: 80 bitmap bytes -> 160 samples, then (lines 3..8) a 480-sample progress bar
        int scale = 255 / 4;
        if (pr.blend != -1 && pr.blend < 32) scale = (scale * pr.blend) >> 5;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int s = x0 + 2 * i - (G::BLIT + 16);                     // both samples of a word share the source byte / bar step
            uint32_t v = 0x1800u;
            if (s < 160) v = 0x1800u + (uint32_t)pr.bitmap[ol * 80 + (s >> 1)] * (uint32_t)scale;
            else if (s >= 176 && ol >= 3 && ol <= 8) v = 0x1800u + ((uint32_t)scale << ((((s - 176) >> 2) * 2 < pr.progress) ? 8 : 7));
            w[i] = (v & 0xFFFFu) | (v << 16);
        }
    } else if (line < G::VSYNC && x0 + 8 <= G::HSYNC) w[0] = w[1] = w[2] = w[3] = SYNC2;
    else if (line < G::VSYNC && (x0 >= G::BURST_START + G::BURST_W || (x0 >= G::HSYNC && x0 + 8 <= G::BURST_START))) w[0] = w[1] = w[2] = w[3] = BLACK2;
    else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            w[i] = blank_sample<kNtsc>(D.pal_burst, line, x0 + 2 * i) | (blank_sample<kNtsc>(D.pal_burst, line, x0 + 2 * i + 1) << 16);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// four luma pixels + two chroma phase words -> eight samples (video.cpp:716-733, verbatim packed arithmetic)
__device__ __forceinline__ uint4 blit_quad(uint32_t y4, uint32_t dither, uint32_t ca, uint32_t cb, uint32_t& lum)
{
    uint32_t p0 = (y4 + dither) & 0xFCFCFCFCu;
    uint32_t p1 = ((p0 >> 1) + (p0 >> 9)) & 0xFCFCFCFCu;
    p0 >>= 2; p1 >>= 2;
    const uint32_t l = (((p0 & 0xFF) + lum) >> 1) & 0xFF;
    lum = p0 >> 24;
    return make_uint4(((l << 24) | ((p0 & 0xFF) << 8)) + ca, ((p1 << 24) | (p0 & 0xFF00)) + (ca << 8),
                      ((p1 << 16) | (p0 >> 8)) + cb, (((p1 << 8) & 0xFF000000u) | (p0 >> 16)) + (cb << 8));
}

#ifndef EF_K2_LEGACY
// ---- K2 v6: a CTA synthesises a BAND of 16 consecutive lines of one stream in shared memory and hands the finished
// rows to the TMA engine (cp.async.bulk.global.shared::cta): the field leaves the SM as bulk stores instead of 16-byte
// STG, which took the output off the L1/TEX pipe - the v4 kernel was bound there (86 % busy at 42 % DRAM throughput,
// profiles/r01_k2_details.txt). Bands coincide with macroblock rows of the tiled frame (the active area starts at
// line 32 / 64), so the loads follow the tile layout: a warp takes two 8-pixel groups x 16 rows = 256 contiguous luma
// bytes and 2 x 64 contiguous chroma bytes of ONE tile (2 + 1 + 1 L1 wavefronts instead of 8 + 8 + 8 when its lanes
// ran along a scan line). What is NOT picture is constant and is not recomputed per line (v5 did, and was bound by
// instruction issue: profiles/r02_k2v5_ncu.txt):
//   * a blank band synthesises each DISTINCT line once (NTSC: one; PAL: two burst phases; the vertical sync types; the
//     overlay lines) and bulk-stores that one shared-memory row to every field line that shows it;
//   * an active band synthesises sync / burst / black margins for two rows (the two PAL burst phases) and stores every
//     line as three bulk copies: left margin and right margin from the template row, the blit span from the line's own.
// Shared-memory rows are padded by 16 bytes: the 16-byte stores of 8 consecutive rows then hit 32 distinct banks.
// grid: x = band (17 NTSC / 20 PAL), y = stream; 352 threads = 11 warps x 2 tile columns = the 22 macroblocks of a row.
#ifndef EF_K2_THREADS
#define EF_K2_THREADS 352
#endif
constexpr int kK2Threads = EF_K2_THREADS;             // 352 (2 tile-column tasks per warp) or 704 (1)
constexpr int kK2Tasks = 22 / (kK2Threads / 32);
static_assert(kK2Tasks * (kK2Threads / 32) == 22, "the 22 macroblock columns of a row must divide over the warps");

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

// which blank lines look alike: lines of one class are sample-identical
template <bool kNtsc>
__device__ __forceinline__ int blank_class(const EfPresent& pr, int line)
{
    using G = Geo<kNtsc>;
    if (line >= G::VSYNC) return kNtsc ? 1 : 16 + (int)((0x00233000u >> ((line - G::VSYNC) * 4)) & 15u);   // _sync_type[], as in blank_sample()
    const int ol = line - (G::TOP + EF_H + 2);
    if (pr.blend != 0 && ol >= 0 && ol < 16) return 64 + ol;              // overlay rows differ line by line
    return kNtsc ? 0 : ((line + 1) & 1) + 2;                              // PAL: the burst table alternates (burst_pal())
}

template <bool kNtsc>
__global__ void __launch_bounds__(kK2Threads)
ef_composite_kernel(const EfDev* __restrict__ Dp, int fb_sel, int frame_counter, const EfPresent pr)
{
    using G = Geo<kNtsc>;
    constexpr int CPL = G::W / 8;                                          // 16-byte chunks per line
    constexpr int LB = G::W * 2, LS = LB + 16;                             // bytes per line in the field / in shared memory
    constexpr int BLIT_C = G::BLIT / 8, BLIT_N = 2 * EF_W / 8;             // the blit span in chunks: [BLIT_C, BLIT_C + 88)
    constexpr int MARGIN_N = CPL - BLIT_N;                                 // chunks of a line outside it
    __shared__ __align__(128) uint8_t sm[16 * LS];
    __shared__ int s_src[16], s_distinct[16], s_n;
    const EfDev& D = *Dp;
    const uint32_t* tab = D.color_tab;                                     // 3 KB chroma LUT through L1
    const int stream = (int)blockIdx.y;
    const int line0 = (int)blockIdx.x * 16;
    const int nlines = min(16, G::LINES - line0);
    const int fl0 = line0 - G::TOP;                                        // frame line of the band's first line (bands = macroblock rows)
    const bool active = fl0 >= 0 && fl0 < EF_H && fb_sel != -2;            // -2: no frame presented yet (video.cpp:1140)
    uint16_t* out0 = D.fields + (size_t)stream * D.field_stride + (size_t)line0 * G::W;

    if (!active) {
        // ---- blank band: every distinct line once ------------------------------------------------------------------
        if (threadIdx.x == 0) {
            int n = 0;
            for (int l = 0; l < nlines; l++) {
                const int c = blank_class<kNtsc>(pr, line0 + l);
                int src = l;
                for (int k = 0; k < l; k++) if (blank_class<kNtsc>(pr, line0 + k) == c) { src = k; break; }
                s_src[l] = src;
                if (src == l) s_distinct[n++] = l;
            }
            s_n = n;
        }
        __syncthreads();
        const int n = s_n;
        for (int i = (int)threadIdx.x; i < n * CPL; i += kK2Threads) {
            const int l = s_distinct[i / CPL], c = i % CPL;
            *(uint4*)(sm + l * LS + c * 16) = blank_chunk<kNtsc>(D, pr, line0 + l, c * 8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy
        __syncthreads();
        if ((int)threadIdx.x < nlines) {
            bulk_store(out0 + (size_t)threadIdx.x * G::W, sm + s_src[threadIdx.x] * LS, LB);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory must stay valid until the engine has read it
        }
        return;
    }

    // ---- active band: margins of the two template rows (rows 0 and 1: even / odd field line) ----------------------------
    for (int i = (int)threadIdx.x; i < 2 * MARGIN_N; i += kK2Threads) {
        const int l = i / MARGIN_N, k = i % MARGIN_N;
        const int c = k >= BLIT_C ? k + BLIT_N : k;
        *(uint4*)(sm + l * LS + c * 16) = blank_chunk<kNtsc>(D, pr, line0 + l, c * 8);
    }
    {
        const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
        const int row = lane & 15, fl = fl0 + row;                         // frame line 0..191
        int fb0 = fb_sel >= 0 ? fb_sel : (int)((D.base_pics[stream] + D.n_pics[stream]) & 1u);
        // two-frame horizontal scroll (video.cpp:1146-1154): blit(f, dst, i, h, 352-h) then blit(f^1, dst + (352-h)*2, i, 0, h)
        int h = pr.hscroll;
        if (h < 0) { h += EF_W; fb0 ^= 1; }
        const int split = (EF_W - h) >> 3;                                 // first 8-pixel group drawn by the second blit
        const uint32_t dither = c_dither[(fl & 3) + ((frame_counter & 1) << 2)];
        const int cy = fl >> 1, ncy = cy + (fl == 191 ? 0 : 1);            // odd lines average with the next chroma row (video.cpp:704-716)
        const int yoff = (fl >> 4) * EF_MBW_MAX * EF_TILE + (fl & 15) * 16;
        const int coff = (cy >> 3) * EF_MBW_MAX * EF_TILE + (cy & 7) * 8 + 256;      // get_cr(line>>1); get_cb is the next 64-byte plane
        const int noff = (fl & 1) ? (ncy >> 3) * EF_MBW_MAX * EF_TILE + (ncy & 7) * 8 + 256 : coff;
        const int vt = (fl & 1) ? 512 : 256;
        const uint8_t* fbase = D.frames + ef_frame_offset(stream, 0);
        // the warp's two tasks (destination groups gd and gd + 22): all loads first, then the arithmetic
        uint32_t u4[kK2Tasks], v4[kK2Tasks], un[kK2Tasks], vn[kK2Tasks], lumw[kK2Tasks];
        uint2 y8[kK2Tasks];
        bool cstart[kK2Tasks];
#pragma unroll
        for (int t = 0; t < kK2Tasks; t++) {
            const int gd = warp * 2 + (lane >> 4) + t * (44 / kK2Tasks);   // destination group 0..43
            const bool second = gd >= split;
            const int g = second ? gd - split : gd + (h >> 3);             // 8-pixel group of the source frame
            cstart[t] = second ? g == 0 : gd == 0;                         // blit() starts its luma carry at 0
            const uint8_t* f = fbase + (size_t)(fb0 ^ (int)second) * EF_FRAME;
            const int tcol = (g >> 1) * EF_TILE, sub = g & 1;
            u4[t] = *(const uint32_t*)(f + coff + tcol + sub * 4);
            v4[t] = *(const uint32_t*)(f + coff + tcol + sub * 4 + 64);
            un[t] = *(const uint32_t*)(f + noff + tcol + sub * 4);
            vn[t] = *(const uint32_t*)(f + noff + tcol + sub * 4 + 64);
            y8[t] = *(const uint2*)(f + yoff + tcol + sub * 8);
            const int q = cstart[t] ? 0 : 2 * g - 1;                       // previous 4-pixel group
            lumw[t] = *(const uint32_t*)(f + yoff + (q >> 2) * EF_TILE + (q & 3) * 4);
        }
#pragma unroll
        for (int t = 0; t < kK2Tasks; t++) {
            const int gd = warp * 2 + (lane >> 4) + t * (44 / kK2Tasks);
            uint32_t u = u4[t], v = v4[t];
            if (fl & 1) {
                u = ((u >> 1) & 0x7F7F7F7Fu) + ((un[t] >> 1) & 0x7F7F7F7Fu);
                v = ((v >> 1) & 0x7F7F7F7Fu) + ((vn[t] >> 1) & 0x7F7F7F7Fu);
            }
            uint32_t lum = cstart[t] ? 0u : ((((lumw[t] + dither) & 0xFCFCFCFCu) >> 2) >> 24);   // carry = last pixel of the previous group
#if EF_K2_ARITH
            const bool vneg = !kNtsc && vt == 512;                         // PAL odd lines: cos_v_neg (video.cpp:584-591)
            const uint4 o0 = blit_quad(y8[t].x, dither, chroma_word_arith(u, v, vneg), chroma_word_arith(u >> 8, v >> 8, vneg), lum);
            const uint4 o1 = blit_quad(y8[t].y, dither, chroma_word_arith(u >> 16, v >> 16, vneg), chroma_word_arith(u >> 24, v >> 24, vneg), lum);
#else
            const uint4 o0 = blit_quad(y8[t].x, dither, chroma_word(tab, u, v, vt), chroma_word(tab, u >> 8, v >> 8, vt), lum);
            const uint4 o1 = blit_quad(y8[t].y, dither, chroma_word(tab, u >> 16, v >> 16, vt), chroma_word(tab, u >> 24, v >> 24, vt), lum);
#endif
            uint8_t* o = sm + row * LS + (G::BLIT + gd * 16) * 2;
            *(uint4*)o = o0;
            *(uint4*)(o + 16) = o1;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 16) {
        const int l = (int)threadIdx.x;
        uint16_t* out = out0 + (size_t)l * G::W;
        const uint8_t* tmpl = sm + (l & 1) * LS;
        bulk_store(out, tmpl, G::BLIT * 2);
        bulk_store(out + G::BLIT, sm + l * LS + G::BLIT * 2, 2 * EF_W * 2);
        bulk_store(out + G::BLIT + 2 * EF_W, tmpl + (G::BLIT + 2 * EF_W) * 2, (G::W - G::BLIT - 2 * EF_W) * 2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
#else
// grid: x = threads of one field / 256, y = stream. One thread = two adjacent 16-byte chunks (32 bytes out,
// 8 luma pixels in: one 8-byte luma load, one 4-byte load per chroma plane). The pairing is phased so that
// the blit span starts on a pair boundary (NTSC chunk 20: pairs (2j, 2j+1); PAL chunk 35: pairs (2j-1, 2j)).
// A warp covers 16 consecutive pairs of TWO consecutive lines (lanes 0-15 the even line, lanes 16-31 the
// odd one): in the tiled frame the two luma rows of a tile share 32-byte sectors and both lines read the same
// chroma row, so every sector the warp fetches is used whole. Stores are 512 contiguous bytes per half-warp.
template <bool kNtsc>
__global__ void __launch_bounds__(256)
ef_composite_kernel(const EfDev* __restrict__ Dp, int fb_sel, int frame_counter, const EfPresent pr)
{
    using G = Geo<kNtsc>;
    constexpr int CPL = G::W / 8;                                          // 16-byte chunks per line (even)
    constexpr int PHASE = (G::BLIT / 8) & 1;                               // 0: pairs (2j, 2j+1); 1: pairs (2j-1, 2j), single chunks at both ends
    constexpr int ITEMS = CPL / 2 + PHASE;                                 // work items per line
    constexpr int GROUPS = (ITEMS + 15) / 16;                              // warps per line pair
    const EfDev& D = *Dp;
#ifndef EF_K2_LUT_SMEM
    const uint32_t* tab = D.color_tab;                                     // 3 KB chroma LUT through L1 (measured faster than a shared-memory copy per CTA)
#else
    __shared__ uint32_t tab[768];
    for (int i = threadIdx.x; i < 768; i += 256) tab[i] = D.color_tab[i];
    __syncthreads();
#endif
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    const uint32_t stream = blockIdx.y;
    const int pair = (int)(t / (GROUPS * 32)), r = (int)(t - (uint32_t)pair * (GROUPS * 32));
    const int lane = r & 31;
    const int line = 2 * pair + (lane >> 4), j = (r >> 5) * 16 + (lane & 15);
    if (pair >= G::LINES / 2 || j >= ITEMS) return;
    const int c0 = 2 * j - PHASE;                                          // first chunk of the item; c0 == -1 / c0 + 1 == CPL: single chunk
    uint16_t* out = D.fields + (size_t)stream * D.field_stride + (size_t)line * G::W;

    const int fl = line - G::TOP;                                          // frame line 0..191 on active lines
    const bool active = fl >= 0 && fl < EF_H && fb_sel != -2;              // -2: no frame presented yet (video.cpp:1140)
    const int x0 = c0 * 8;
    if (active && x0 >= G::BLIT && x0 < G::BLIT + 2 * EF_W) {
        int fb = fb_sel >= 0 ? fb_sel : (int)((D.base_pics[stream] + D.n_pics[stream]) & 1u);
        // two-frame horizontal scroll (video.cpp:1146-1154): blit(f, dst, i, h, 352-h) then blit(f^1, dst + (352-h)*2, i, 0, h)
        int h = pr.hscroll;
        if (h < 0) { h += EF_W; fb ^= 1; }
        const int gd = (x0 - G::BLIT) >> 4;                                // 8-pixel group of the destination, 0..43
        const int split = (EF_W - h) >> 3;                                 // first group drawn by the second blit
        const bool second = gd >= split;
        if (second) fb ^= 1;
        const int g = second ? gd - split : gd + (h >> 3);                 // 8-pixel group of the source frame
        const bool call_start = second ? g == 0 : gd == 0;                 // blit() starts its luma carry at 0
        const uint8_t* f = D.frames + ef_frame_offset((int)stream, fb);
        const uint32_t dither = c_dither[(fl & 3) + ((frame_counter & 1) << 2)];
        // tiled frame (ef_common.cuh): the 8 luma pixels of group g are half a tile row, the 4 chroma samples half a chroma tile row
        const int cy = fl >> 1;
        const int tcol = (g >> 1) * EF_TILE;
        const uint8_t* yrow = f + (fl >> 4) * EF_MBW_MAX * EF_TILE + (fl & 15) * 16;
        const uint8_t* crow = f + (cy >> 3) * EF_MBW_MAX * EF_TILE + (cy & 7) * 8 + 256 + (g & 1) * 4;   // get_cr(line>>1); get_cb is the next 64-byte plane
        uint32_t u4 = *(const uint32_t*)(crow + tcol);
        uint32_t v4 = *(const uint32_t*)(crow + tcol + 64);
        int vt = 256;
        if (fl & 1) {                                                      // odd lines average with the next chroma row (video.cpp:704-716)
            const int n = cy + (fl == 191 ? 0 : 1);
            const uint8_t* nrow = f + (n >> 3) * EF_MBW_MAX * EF_TILE + (n & 7) * 8 + 256 + (g & 1) * 4;
            u4 = ((u4 >> 1) & 0x7F7F7F7Fu) + ((*(const uint32_t*)(nrow + tcol) >> 1) & 0x7F7F7F7Fu);
            v4 = ((v4 >> 1) & 0x7F7F7F7Fu) + ((*(const uint32_t*)(nrow + tcol + 64) >> 1) & 0x7F7F7F7Fu);
            vt = 512;
        }
        const uint2 y8 = *(const uint2*)(yrow + tcol + (g & 1) * 8);
        uint32_t lum = 0;                                                  // carry = last pixel of the previous group, 0 at the start of a blit() call
        if (!call_start) {
            const int q = 2 * g - 1;                                       // previous 4-pixel group
            lum = ((((*(const uint32_t*)(yrow + (q >> 2) * EF_TILE + (q & 3) * 4) + dither) & 0xFCFCFCFCu) >> 2) >> 24);
        }
        const uint4 o0 = blit_quad(y8.x, dither, chroma_word(tab, u4, v4, vt), chroma_word(tab, u4 >> 8, v4 >> 8, vt), lum);
        const uint4 o1 = blit_quad(y8.y, dither, chroma_word(tab, u4 >> 16, v4 >> 16, vt), chroma_word(tab, u4 >> 24, v4 >> 24, vt), lum);
        __stcs((uint4*)(out + x0), o0);                                    // streaming stores: the field is not re-read by this kernel
        __stcs((uint4*)(out + x0 + 8), o1);
    } else {
        if (c0 >= 0) __stcs((uint4*)(out + x0), blank_chunk<kNtsc>(D, pr, line, x0));
        if (c0 + 1 < CPL) __stcs((uint4*)(out + x0 + 8), blank_chunk<kNtsc>(D, pr, line, x0 + 8));
    }
}

#endif

// single blit() call into a device buffer (line-blit entry point; used by ef_blit): one thread per 4 luma pixels
__global__ void ef_blit_kernel(const EfDev* __restrict__ Dp, int stream, int fb, int fl, int x, int width, int frame_counter, uint16_t* __restrict__ dst)
{
    const EfDev& D = *Dp;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;                   // group index relative to x
    x &= ~3;
    if (q * 4 >= width) return;
    const uint8_t* f = D.frames + ef_frame_offset(stream, fb);
    const uint32_t dither = c_dither[(fl & 3) + ((frame_counter & 1) << 2)];
    const int g = (x >> 2) + q;                                            // absolute 4-pixel group index
    const int cy = fl >> 1;
    const int ycol = (g >> 2) * EF_TILE + (g & 3) * 4, ccol = (g >> 2) * EF_TILE + 256 + (g & 3) * 2;
    const uint8_t* yrow = f + (fl >> 4) * EF_MBW_MAX * EF_TILE + (fl & 15) * 16;
    const uint8_t* crow = f + (cy >> 3) * EF_MBW_MAX * EF_TILE + (cy & 7) * 8;
    uint32_t u2 = *(const uint16_t*)(crow + ccol);
    uint32_t v2 = *(const uint16_t*)(crow + ccol + 64);
    int vt = 256;
    if (fl & 1) {
        const int n = cy + (fl == 191 ? 0 : 1);
        const uint8_t* nrow = f + (n >> 3) * EF_MBW_MAX * EF_TILE + (n & 7) * 8;
        const uint32_t ub = *(const uint16_t*)(nrow + ccol);
        const uint32_t vb = *(const uint16_t*)(nrow + ccol + 64);
        u2 = ((u2 >> 1) & 0x7F7Fu) + ((ub >> 1) & 0x7F7Fu);
        v2 = ((v2 >> 1) & 0x7F7Fu) + ((vb >> 1) & 0x7F7Fu);
        vt = 512;
    }
    const uint32_t* tab = D.color_tab;
    const uint32_t ca = chroma_word(tab, u2, v2, vt), cb = chroma_word(tab, u2 >> 8, v2 >> 8, vt);
    uint32_t lum = 0;
    if (q > 0) lum = ((((*(const uint32_t*)(yrow + ((g - 1) >> 2) * EF_TILE + ((g - 1) & 3) * 4) + dither) & 0xFCFCFCFCu) >> 2) >> 24);
    uint32_t p0 = (*(const uint32_t*)(yrow + ycol) + dither) & 0xFCFCFCFCu;
    uint32_t p1 = ((p0 >> 1) + (p0 >> 9)) & 0xFCFCFCFCu;
    p0 >>= 2; p1 >>= 2;
    lum = (((p0 & 0xFF) + lum) >> 1) & 0xFF;
    uint32_t* o = (uint32_t*)(dst + q * 8);
    o[0] = ((lum << 24) | ((p0 & 0xFF) << 8)) + ca;
    o[1] = ((p1 << 24) | (p0 & 0xFF00)) + (ca << 8);
    o[2] = ((p1 << 16) | (p0 >> 8)) + cb;
    o[3] = (((p1 << 8) & 0xFF000000u) | (p0 >> 16)) + (cb << 8);
}

cudaError_t ef_launch_composite(const EfDev* dev, int n_streams, const EfGeometry& g, int fb, int frame_counter, const EfPresent& pr, cudaStream_t stream)
{
#ifndef EF_K2_LEGACY
    const dim3 bands((unsigned)(g.line_count + 15) / 16, (unsigned)n_streams);
    if (g.ntsc) ef_composite_kernel<true><<<bands, kK2Threads, 0, stream>>>(dev, fb, frame_counter, pr);
    else ef_composite_kernel<false><<<bands, kK2Threads, 0, stream>>>(dev, fb, frame_counter, pr);
    return cudaGetLastError();
#else
    const unsigned items = (unsigned)(g.line_width >> 4) + (((unsigned)g.blit_start >> 3) & 1u);   // 32-byte work items per line (see the kernel)
    const unsigned groups = (items + 15) / 16;
    const unsigned threads = groups * 32 * (unsigned)(g.line_count / 2);   // both standards have an even line count
    const dim3 grid((threads + 255) / 256, (unsigned)n_streams);
    if (g.ntsc) ef_composite_kernel<true><<<grid, 256, 0, stream>>>(dev, fb, frame_counter, pr);
    else ef_composite_kernel<false><<<grid, 256, 0, stream>>>(dev, fb, frame_counter, pr);
    return cudaGetLastError();
#endif
}

cudaError_t ef_launch_blit(const EfDev* dev, int stream_index, int fb, int line, int x, int width, int frame_counter, uint16_t* dst, cudaStream_t stream)
{
    const int groups = (width + 3) / 4;
    ef_blit_kernel<<<(groups + 63) / 64, 64, 0, stream>>>(dev, stream_index, fb, line, x, width, frame_counter, dst);
    return cudaGetLastError();
}
