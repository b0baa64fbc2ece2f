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

__device__ __forceinline__ uint16_t blank_sample(const EfGeometry& g, const int16_t* pal_burst, int line, int x)
{
    const uint16_t SYNC = 0x0000, BLANKING = 0x1400, BLACK = 0x1800;       // IRE(-40), IRE(0), IRE(7.5): video.cpp:520-525
    if (line >= g.vsync_start) {
        if (g.ntsc) return x < g.hsync_long ? SYNC : BLANKING;             // blanking(buf, true)
        const uint32_t types = 0x00233000u;                                // _sync_type[8] = {0,0,0,3,3,2,0,0}, one nibble each
        const int t = (types >> ((line - g.vsync_start) * 4)) & 15;
        const int half = g.line_width >> 1;
        const int second = x >= half;
        const int xx = second ? x - half : x;
        const int lng = second ? (t & 1) : (t & 2);
        return xx < (lng ? g.hsync_long : g.hsync_short) ? SYNC : BLANKING; // pal_sync2
    }
    if (x < g.hsync) return SYNC;
    if (g.ntsc) {
        const int i = x - g.hsync;                                         // burst(), video.cpp:806: 10 cycles of [1E00 1400 0A00 1400]
        if (i < 40) return (i & 1) ? BLANKING : ((i & 2) ? 0x0A00 : 0x1E00);
    } else {
        const int i = x - g.burst_start;                                   // burst_pal(): pair-swapped copy, table chosen by (_line_counter after ++) & 1
        if (i >= 0 && i < g.burst_width) return (uint16_t)pal_burst[(((line + 1) & 1) ? 0 : 64) + (i ^ 1)];
    }
    return BLACK;
}

__constant__ uint32_t c_dither[8] = {                                     // dither4x4, video.cpp:673
    0x00020301, 0x03010002, 0x02030100, 0x01000203, 0x03010002, 0x00020301, 0x01000203, 0x02030100 };

}  // namespace

__global__ void __launch_bounds__(256)
ef_composite_kernel(const EfDev* __restrict__ Dp, int fb_sel, int frame_counter)
{
    __shared__ uint32_t tab[768];
    const EfDev& D = *Dp;
    for (int i = threadIdx.x; i < 768; i += blockDim.x) tab[i] = D.color_tab[i];
    __syncthreads();

    const EfGeometry g = D.geo;
    const int cpl = g.line_width >> 3;                                     // 16-byte chunks per line
    const uint32_t chunks = (uint32_t)(cpl * g.line_count);
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stream = (uint32_t)(gid / chunks);
    const uint32_t c = (uint32_t)(gid % chunks);
    if (stream >= (uint32_t)D.n_streams) return;

    const int line = (int)(c / (uint32_t)cpl), k = (int)(c % (uint32_t)cpl);
    const int x0 = k * 8;
    uint16_t* out = D.fields + (size_t)stream * D.field_stride + (size_t)line * g.line_width + x0;

    const int fl = line - g.active_top;                                    // frame line 0..191 on active lines
    const bool active = fl >= 0 && fl < EF_H && fb_sel != -2;      // -2: no frame presented yet (video.cpp:1140)
    uint32_t w[4];
    if (active && x0 >= g.blit_start && x0 < g.blit_start + 2 * EF_W) {
        const int fb = fb_sel >= 0 ? fb_sel : (int)((D.base_pics[stream] + D.n_pics[stream]) & 1u);
        const uint8_t* f = D.frames + ef_frame_offset((int)stream, fb);
        const int q = (x0 - g.blit_start) >> 3;                            // 4-pixel group index 0..87
        const uint32_t dither = c_dither[(fl & 3) + ((frame_counter & 1) << 2)];
        // tiled frame (ef_common.cuh): 4 luma pixels of group q sit in tile q>>2, the 2 chroma samples too
        const int cy = fl >> 1;
        const int ycol = (q >> 2) * EF_TILE + (q & 3) * 4, ccol = (q >> 2) * EF_TILE + 256 + (q & 3) * 2;
        const uint8_t* yrow = f + (fl >> 4) * EF_MBW_MAX * EF_TILE + (fl & 15) * 16;
        const uint8_t* crow = f + (cy >> 3) * EF_MBW_MAX * EF_TILE + (cy & 7) * 8;       // get_cr(line>>1); get_cb is the next 64-byte plane
        uint32_t u2 = *(const uint16_t*)(crow + ccol);
        uint32_t v2 = *(const uint16_t*)(crow + ccol + 64);
        int vt = 256;
        if (fl & 1) {                                                      // odd lines average with the next chroma row (video.cpp:704-716)
            const int n = cy + (fl == 191 ? 0 : 1);
            const uint8_t* nrow = f + (n >> 3) * EF_MBW_MAX * EF_TILE + (n & 7) * 8;
            const uint32_t ub = *(const uint16_t*)(nrow + ccol);
            const uint32_t vb = *(const uint16_t*)(nrow + ccol + 64);
            u2 = ((u2 >> 1) & 0x7F7Fu) + ((ub >> 1) & 0x7F7Fu);
            v2 = ((v2 >> 1) & 0x7F7Fu) + ((vb >> 1) & 0x7F7Fu);
            vt = 512;
        }
        const uint32_t ca = chroma_word(tab, u2, v2, vt), cb = chroma_word(tab, u2 >> 8, v2 >> 8, vt);
        uint32_t lum = 0;                                                  // carry = pixel 3 of the previous group, 0 at the line start
        if (q > 0) lum = ((((*(const uint32_t*)(yrow + ((q - 1) >> 2) * EF_TILE + ((q - 1) & 3) * 4) + dither) & 0xFCFCFCFCu) >> 2) >> 24);
        uint32_t p0 = (*(const uint32_t*)(yrow + ycol) + dither) & 0xFCFCFCFCu;   // video.cpp:716-733, verbatim packed arithmetic
        uint32_t p1 = ((p0 >> 1) + (p0 >> 9)) & 0xFCFCFCFCu;
        p0 >>= 2; p1 >>= 2;
        lum = (((p0 & 0xFF) + lum) >> 1) & 0xFF;
        w[0] = ((lum << 24) | ((p0 & 0xFF) << 8)) + ca;
        w[1] = ((p1 << 24) | (p0 & 0xFF00)) + (ca << 8);
        w[2] = ((p1 << 16) | (p0 >> 8)) + cb;
        w[3] = (((p1 << 8) & 0xFF000000u) | (p0 >> 16)) + (cb << 8);
    } else {
        // blank and vsync lines, and the part of an active line outside the blit span (which still
        // shows what blanking() last left in the ping-pong buffer: sync, burst, BLACK)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint16_t a = blank_sample(g, D.pal_burst, line, x0 + 2 * i);
            const uint16_t b = blank_sample(g, D.pal_burst, line, x0 + 2 * i + 1);
            w[i] = (uint32_t)a | ((uint32_t)b << 16);
        }
    }
    *(uint4*)out = make_uint4(w[0], w[1], w[2], w[3]);
}

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

cudaError_t ef_launch_composite(const EfDev* dev, int n_streams, const EfGeometry& g, int fb, int frame_counter, cudaStream_t stream)
{
    const uint64_t chunks = (uint64_t)(g.line_width >> 3) * g.line_count * (uint64_t)n_streams;
    const uint64_t blocks = (chunks + 255) / 256;
    ef_composite_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dev, fb, frame_counter);
    return cudaGetLastError();
}

cudaError_t ef_launch_blit(const EfDev* dev, int stream_index, int fb, int line, int x, int width, int frame_counter, uint16_t* dst, cudaStream_t stream)
{
    const int groups = (width + 3) / 4;
    ef_blit_kernel<<<(groups + 63) / 64, 64, 0, stream>>>(dev, stream_index, fb, line, x, width, frame_counter, dst);
    return cudaGetLastError();
}
