// espflix_b200/csrc/ef_decode.cu — K1, the fused macroblock kernel (one launch per picture index
// over the whole batch of streams).
//
// Replaces, for every stream at once, MpegDecoder::slice() and everything under it
// (player.cpp:733-1316): macroblock-header VLC, motion vectors, block() coefficient VLC +
// dequantisation, idct(), mocomp()/predict_zero(), copy_block/add_block with the [0,248] clamp.
//
// Mapping (B200-first, not a translation of the reference's single scalar loop):
//   * unit of work = one slice of one stream (EfWork). VLC parsing is serial inside a slice, so the
//     parallelism is streams x slices: every LANE of a warp parses a different slice, one
//     macroblock per iteration. The macroblock header is parsed by all lanes together; the
//     coefficients are then decoded by a FLAT per-lane state machine (one symbol per step, lanes
//     move through their blocks independently): CLZ-indexed shared-memory table -> (run, level,
//     length) in one load, dequantised on the spot (quirk Q2 included) and appended as a 32-bit
//     (block, scan position, value) entry to the lane's list in shared memory (96 entries; the rare
//     longer macroblock spills to a per-lane HBM area).
//   * then the whole WARP reconstructs the 32 macroblocks one after another: scatter the list into
//     a dense 6x64 scratch (values already de-zigzagged and AAN-prescaled by the parser), the
//     reference's integer AAN IDCT in place with one lane per block column, then per block row
//     (4 luma blocks = 32 lanes), half-pel motion compensation and the clamped add with one lane
//     per 8-pixel row segment.
//   * frame stores are MACROBLOCK-TILED in HBM (ef_common.cuh): a macroblock is 384 contiguous
//     bytes, so the warp's stores are two full 128-byte lines + one more for chroma, no partial
//     sectors; motion-compensated reads touch <= 4 tiles.
//   * lanes that finish a slice pull the next one from a global cursor, so lane occupancy stays
//     high until the picture's work list is empty (persistent CTAs, one per SM, 14 warps).
// Bit-exactness notes: Q1 clamp [0,248]; Q2 oddification maps 0 -> +1; Q3 chroma
// vector = floor(luma position / 2); Q4 matrices indexed in raster order (done at index time);
// Q5 single-coefficient blocks bypass the IDCT with floor; Q6 first macroblock of a slice lands in
// column 0; Q7 intra DC-only blocks are replicated unclamped.
#include "ef_common.cuh"

namespace {

constexpr int kWarpsPerCta = EF_K1_WARPS;
constexpr int kListEntries = EF_K1_LIST;        // per-lane coefficient list capacity in shared memory
#ifdef EF_K1_HALF
constexpr int kLanes = 16;                      // tuning experiment: 16 slices per warp
#else
constexpr int kLanes = 32;                      // slices (parser lanes) per warp
#endif
constexpr int kListBytes = kLanes * kListEntries * 4;
constexpr int kHdrStride = 44;                  // bytes per lane header; 11 words (odd) -> conflict-free
constexpr int kHdrDc = 0;                       // int32 [6] intra DC (pixel scale)
constexpr int kHdrInfo = 24;                    // bit0 valid, 1 intra, 2-7 coded blocks, 8-13 n==1 mask, 14-19 abort mask (bit b = block b), 20-24 mb_x, 25-28 mb_y
constexpr int kHdrCnt = 28;                     // list entries | skip_before << 16
constexpr int kHdrMv = 32;                      // (int16 h) | (int16 v) << 16, half-pel units
constexpr int kHdrBytes = kLanes * kHdrStride;
constexpr int kDenseStride = 72;                // words per block in the dense scratch: 64 + 8 pad -> the 4 luma blocks hit distinct banks
constexpr int kDenseBytes = 6 * kDenseStride * 4;   // int32 [6][72] prescaled coefficients, raster order; also the IDCT transpose buffer
constexpr int kStageBytes = 4 * EF_TILE;        // motion-compensation staging: up to 2 x 2 reference tiles per macroblock
constexpr int kWarpBytes = kListBytes + kHdrBytes + kDenseBytes + kStageBytes + 16;   // + the warp's mbarrier

struct SharedTables {                           // same layout as the head of EfTables
    uint16_t dct[26 * 32];
    uint16_t mba[8 * 32];
    uint16_t mv[7 * 32];
    uint16_t cbp[512];
    uint8_t ptype[64];
    uint8_t qdef[128];
    uint16_t zp[64];                            // scan position -> raster index | AAN prescale << 8
};
constexpr int kTableBytes = (sizeof(SharedTables) + 15) & ~15;

// ---------------------------------------------------------------------------------------------
// bit reader (FILL_BITS/peek_bits/get_bits, player.cpp:348-352, 495-514): MSB-first. `hi` holds
// the current 32-bit word, `lo` the next one, `nx` the one after (prefetched), pos = bits of `hi`
// already consumed. peek() is a single funnel shift. Reads run at most 12 bytes past the slice
// plus the 8 prefetched ones (into the next start code); the ES blob carries 256 bytes of zero padding at its end.
// ---------------------------------------------------------------------------------------------
struct BitReader {
    const uint32_t* words;   // the whole ES blob as aligned 32-bit words (cudaMalloc alignment)
    uint32_t idx;            // next word to fetch
    uint32_t hi, lo, nx_raw, nx2_raw;   // two prefetched words, still little-endian: their loads are not waited for until they are needed
    int pos;

    __device__ __forceinline__ uint32_t fetch_raw()
    {
        const uint32_t* a = words + idx;
        const uint32_t v = __ldg(a);
        if ((idx & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 32));   // 128 bytes ahead of this slice's read position
        idx++;
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* blob, uint64_t byte_off)
    {
        words = (const uint32_t*)blob;
        idx = (uint32_t)(byte_off >> 2);
        pos = (int)(byte_off & 3) * 8;
        hi = __byte_perm(fetch_raw(), 0, 0x0123); lo = __byte_perm(fetch_raw(), 0, 0x0123); nx_raw = fetch_raw(); nx2_raw = fetch_raw();
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, pos); }
    __device__ __forceinline__ void skip(int n)
    {
        pos += n;
        if (pos >= 32) { pos -= 32; hi = lo; lo = __byte_perm(nx_raw, 0, 0x0123); nx_raw = nx2_raw; nx2_raw = fetch_raw(); }
    }
    __device__ __forceinline__ uint32_t get(int n)     // 1 <= n <= 32
    {
        uint32_t v = peek() >> (32 - n);
        skip(n);
        return v;
    }
};

// per-lane slice parser state
struct SliceState {
    BitReader br;
    uint8_t* cur;
    const uint8_t* ref;
    const uint8_t* qtab;     // scan-order quantiser tables [intra 64 | non-intra 64]: the shared-memory defaults or the stream's own in HBM
    int mbw, mbh;
    int mb_x, mb_y;          // last macroblock handled
    int first;               // next macroblock is the first of the slice (Q6)
    int ptype, full_pel, r_size;
    int qscale;
    int dc_y, dc_cr, dc_cb;  // reference names: cr = block 4, cb = block 5 (player.cpp:728)
    int mv_h, mv_v;
};

__device__ __forceinline__ int motion_component(BitReader& br, const uint16_t* mvtab, int m, int r_size, bool& bad)
{
    // motion_vector(), player.cpp:891
    uint32_t bits = br.peek();
    int lz = __clz(bits);
    if (lz > 6) { bad = true; return m; }
    uint32_t e = mvtab[lz * 32 + ((bits << (lz + 1)) >> 27)];
    int len = e & 15;
    if (!len) { bad = true; return m; }
    int code = (int)(e >> 4) - 16;
    br.skip(len);
    int d = code;
    if (code != 0 && r_size != 0) {
        d = ((abs(code) - 1) << r_size) + (int)br.get(r_size) + 1;
        if (code < 0) d = -d;
    }
    int scale = 1 << r_size;
    m += d;
    if (m > (scale << 4) - 1) m -= scale << 5;
    else if (m < -(scale << 4)) m += scale << 5;
    return m;
}

// One 8-point pass of the reference IDCT (player.cpp:938-995), on registers.
template <bool kFinal>
__device__ __forceinline__ void idct8(int (&v)[8])
{
    int b1 = v[4];
    int b3 = v[2] + v[6];
    int b4 = v[5] - v[3];
    int t1 = v[1] + v[7];
    int t2 = v[3] + v[5];
    int b6 = v[1] - v[7];
    int b7 = t1 + t2;
    int m0 = kFinal ? v[0] + 128 : v[0];            // every output carries m0 exactly once: the final (x + 128) >> 8 needs one add
    int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    int x0 = x4 - (((t1 - t2) * 362 + 128) >> 8);
    int x1 = m0 - b1;
    int x2 = (((v[2] - v[6]) * 362 + 128) >> 8) - b3;
    int x3 = m0 + b1;
    int y3 = x1 + x2, y4 = x3 + b3, y5 = x1 - x2, y6 = x3 - b3;
    int y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
    v[0] = b7 + y4; v[1] = x4 + y3; v[2] = y5 - x0; v[3] = y6 - y7;
    v[4] = y6 + y7; v[5] = x0 + y5; v[6] = y3 - x4; v[7] = y4 - b7;
    if (kFinal) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] >>= 8;
    }
}

// dequantise one level (block(), player.cpp:1110-1119): v = 2*level (+-1 when not intra);
// v = (v * qscale * q) / 16 with C truncation; oddify (Q2: 0 becomes +1); clamp to [-2048, 2047].
__device__ __forceinline__ int dequant(int level, int intra, int qsq)
{
    int v = level << 1;
    if (!intra) v += v < 0 ? -1 : 1;
    v = (v * qsq) / 16;
    if ((v & 1) == 0) v -= v > 0 ? 1 : -1;
    return max(-2048, min(2047, v));
}

__device__ __forceinline__ uint32_t pin4(uint32_t pred, int r0, int r1, int r2, int r3)
{
    // PIN(b + s) for four pixels (add_block, player.cpp:1189; _pin clamps to [0,248], Q1):
    // max(min(pred + res, 248), 0) is one DPX instruction per pixel. (The packed s16x2 form would
    // halve this but its 16-bit add could wrap for residuals beyond +-32,512; kept exact for any int.)
    const int p0 = __viaddmin_s32_relu((int)(pred & 0xFF), r0, 248);
    const int p1 = __viaddmin_s32_relu((int)((pred >> 8) & 0xFF), r1, 248);
    const int p2 = __viaddmin_s32_relu((int)((pred >> 16) & 0xFF), r2, 248);
    const int p3 = __viaddmin_s32_relu((int)(pred >> 24), r3, 248);
    return __byte_perm(__byte_perm(p0, p1, 0x0040), __byte_perm(p2, p3, 0x0040), 0x5410);
}

// (a+b+1)>>1 on four packed bytes (mocomp cases 1 and 2, player.cpp:777-805)
__device__ __forceinline__ uint32_t avg2x4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu); }

// (a+b+c+d+2)>>2 on four packed bytes (mocomp case 3, player.cpp:806)
__device__ __forceinline__ uint32_t avg4x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    const uint32_t m = 0x00FF00FFu;
    uint32_t e = (a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u;
    uint32_t o = ((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u;
    return ((e >> 2) & m) | (((o >> 2) & m) << 8);
}

// ---- motion-compensation source: TMA-engine bulk copies of whole reference tiles into shared memory ----
// A macroblock's prediction window (17 x 17 luma + 2 x 9 x 9 chroma at a half-pel vector) lies in at
// most 2 x 2 tiles of the reference frame, and a tile (384 B: Y, block-4, block-5 chroma) is contiguous
// in HBM, so one elected lane issues 1, 2 or 4 cp.async.bulk copies per macroblock, completion is
// signalled on a per-warp mbarrier, and all lanes then read their pixels from shared memory. The copies
// are issued before the IDCT work of the macroblock and waited for after it.
struct PredWords { uint32_t a0, a1, a2, b0, b1, b2; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t sbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes) : "memory");
}
// one or two horizontally adjacent reference tiles (384 / 768 contiguous bytes) HBM -> shared memory
// through the TMA engine; sdst / sbar are shared-window addresses
__device__ __forceinline__ void bulk_tiles(uint32_t sdst, const uint8_t* gsrc, uint32_t bytes, uint32_t sbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(sbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t sbar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(sbar), "r"(parity) : "memory");
    } while (!ok);
}

// Words of this lane's 8-pixel segment from the staged tiles. (x, y) = first pixel in plane
// coordinates, (tx0, ty0) = tile of the window's top-left corner. kLuma: 16-pixel tile rows, else
// 8-pixel chroma tile rows of plane 0/1.
template <bool kLuma>
__device__ __forceinline__ void pred_words_staged(const uint8_t* st, int plane, int x, int y, int tx0, int ty0, bool yh, PredWords& w)
{
    constexpr int TS = kLuma ? 16 : 8, SH = kLuma ? 4 : 3;
    const int xa = x & ~3, xi = xa & (TS - 1), yi = y & (TS - 1);
    // word 0, then +4 bytes each, hopping to the next staged tile column (+384) at the end of a tile row;
    // a hop out of the 2 x 2 window can only happen for the third word when it is not used: keep it inside
    const int o0 = ((y >> SH) - ty0) * (2 * EF_TILE) + ((xa >> SH) - tx0) * EF_TILE + (kLuma ? 0 : 256 + plane * 64) + yi * TS + xi;
    const int o1 = o0 + (xi + 4 < TS ? 4 : EF_TILE + 4 - TS);
    const int xj = (xi + 4) & (TS - 1);
    int o2 = o1 + (xj + 4 < TS ? 4 : EF_TILE + 4 - TS);
    if (((xa + 8) >> SH) - tx0 > 1) o2 = o1;
    w.a0 = *(const uint32_t*)(st + o0); w.a1 = *(const uint32_t*)(st + o1); w.a2 = *(const uint32_t*)(st + o2);
    if (yh) {
        const int d = yi == TS - 1 ? 2 * EF_TILE - (TS - 1) * TS : TS;    // next row: same tile, or the tile below
        w.b0 = *(const uint32_t*)(st + o0 + d); w.b1 = *(const uint32_t*)(st + o1 + d); w.b2 = *(const uint32_t*)(st + o2 + d);
    }
}

// Same words straight from HBM with coordinates clamped into the frame: only for vectors that
// point outside the picture, where the reference reads whatever lies there (plain byte semantics of mocomp()).
template <bool kLuma>
__device__ __forceinline__ void pred_words_clamped(const uint8_t* ref, int plane, int x, int y, bool yh, PredWords& w)
{
    constexpr int W = kLuma ? EF_W : EF_W / 2, H = kLuma ? EF_H : EF_H / 2, TS = kLuma ? 16 : 8, SH = kLuma ? 4 : 3;
    const int plane_off = kLuma ? 0 : 256 + plane * 64;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        if (r == 1 && !yh) break;
        const int yy = max(0, min(H - 1, y + r));
        const int xa = max(0, min(W - 4, x & ~3)), xb = min(W - 4, xa + 4), xc = min(W - 4, xa + 8);
        const int rowbase = (yy >> SH) * EF_MBW_MAX * EF_TILE + plane_off + (yy & (TS - 1)) * TS;
        const uint32_t v0 = *(const uint32_t*)(ref + rowbase + (xa >> SH) * EF_TILE + (xa & (TS - 1)));
        const uint32_t v1 = *(const uint32_t*)(ref + rowbase + (xb >> SH) * EF_TILE + (xb & (TS - 1)));
        const uint32_t v2 = *(const uint32_t*)(ref + rowbase + (xc >> SH) * EF_TILE + (xc & (TS - 1)));
        if (r == 0) { w.a0 = v0; w.a1 = v1; w.a2 = v2; } else { w.b0 = v0; w.b1 = v1; w.b2 = v2; }
    }
}

// eight predicted pixels from the loaded words (the four cases of mocomp(), player.cpp:767-820)
__device__ __forceinline__ void pred_finish(const PredWords& w, int x, int xh, int yh, uint32_t& o0, uint32_t& o1)
{
    const int sh = (x & 3) * 8;
    const uint32_t p0 = __funnelshift_r(w.a0, w.a1, sh), p1 = __funnelshift_r(w.a1, w.a2, sh);
    if (xh) {
        const uint32_t q0 = __funnelshift_rc(w.a0, w.a1, sh + 8), q1 = __funnelshift_rc(w.a1, w.a2, sh + 8);
        if (yh) {
            o0 = avg4x4(p0, q0, __funnelshift_r(w.b0, w.b1, sh), __funnelshift_rc(w.b0, w.b1, sh + 8));
            o1 = avg4x4(p1, q1, __funnelshift_r(w.b1, w.b2, sh), __funnelshift_rc(w.b1, w.b2, sh + 8));
        } else {
            o0 = avg2x4(p0, q0);
            o1 = avg2x4(p1, q1);
        }
    } else if (yh) {
        o0 = avg2x4(p0, __funnelshift_r(w.b0, w.b1, sh));
        o1 = avg2x4(p1, __funnelshift_r(w.b1, w.b2, sh));
    } else {
        o0 = p0; o1 = p1;
    }
}

// ---------------------------------------------------------------------------------------------
// macroblock header of this lane's slice (player.cpp:1266-1307). Returns false when the slice
// ended. On success: hdr is filled, `cbp` = blocks to parse (bit b = block b), `intra` set.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool parse_header(SliceState& s, uint8_t* hdr, const SharedTables& T, int& cbp_out, int& intra_out)
{
    BitReader& br = s.br;
    uint32_t bits = br.peek();
    if ((bits >> 9) == 0) return false;             // slice_done(): next 23 bits are zero (player.cpp:1238)

    int increment = 0;                              // macroblock_address_increment (player.cpp:1267-1275)
    for (;;) {
        bits = br.peek();
        int lz = __clz(bits);
        if (lz > 7) return false;
        uint32_t e = T.mba[lz * 32 + ((bits << (lz + 1)) >> 27)];
        int len = e & 15, val = (int)(e >> 4);
        if (!len) return false;
        br.skip(len);
        if (val == 34) continue;                    // stuffing
        if (val == 35) { increment += 33; continue; }   // escape
        increment += val;
        break;
    }
    int skip_before = 0;
    if (s.first) { s.first = 0; increment = 1; }    // inc_mb ignores its argument for the first macroblock (Q6)
    else if (increment > 1) { s.dc_y = s.dc_cr = s.dc_cb = 128; s.mv_h = s.mv_v = 0; skip_before = increment - 1; }
    s.mb_x += increment;
    while (s.mb_x >= s.mbw) { s.mb_x -= s.mbw; s.mb_y++; }     // inc_mb(), player.cpp:823
    if (s.mb_y >= s.mbh) return false;              // the reference would write past the frame here

    int mb_type;                                    // macroblock_type (player.cpp:1292)
    bits = br.peek();
    if (s.ptype == 1) {
        if (bits >> 31) { mb_type = 0x01; br.skip(1); }
        else if (bits >> 30) { mb_type = 0x11; br.skip(2); }
        else return false;
    } else {
        uint32_t e = T.ptype[bits >> 26];
        if (!(e & 7)) return false;
        mb_type = (int)(e >> 3);
        br.skip(e & 7);
    }
    const int intra = mb_type & 1;
    if (mb_type & 0x10) s.qscale = (int)br.get(5);
    int mvh = 0, mvv = 0;
    if (intra) { s.mv_h = s.mv_v = 0; }
    else {
        s.dc_y = s.dc_cr = s.dc_cb = 128;
        if (mb_type & 0x08) {
            bool bad = false;
            s.mv_h = motion_component(br, T.mv, s.mv_h, s.r_size, bad);
            s.mv_v = motion_component(br, T.mv, s.mv_v, s.r_size, bad);
            if (bad) return false;
        } else s.mv_h = s.mv_v = 0;
        mvh = s.mv_h << s.full_pel;                 // predict(), player.cpp:878
        mvv = s.mv_v << s.full_pel;
    }
    int cbp = intra ? 63 : 0;
    if (mb_type & 0x02) {
        bits = br.peek();
        uint32_t e = T.cbp[bits >> 23];
        if (!(e & 15)) return false;
        cbp = (int)(e >> 4);
        br.skip(e & 15);
    }
    cbp = (int)(__brev((unsigned)cbp) >> 26);       // bit b = block b (the VLC value has block 0 in bit 5)
    *(uint32_t*)(hdr + kHdrInfo) = 1u | ((uint32_t)intra << 1) | ((uint32_t)cbp << 2) | ((uint32_t)s.mb_x << 20) | ((uint32_t)s.mb_y << 25);
    *(uint32_t*)(hdr + kHdrCnt) = (uint32_t)skip_before << 16;
    *(uint32_t*)(hdr + kHdrMv) = ((uint32_t)mvh & 0xFFFFu) | ((uint32_t)mvv << 16);
    cbp_out = cbp; intra_out = intra;
    return true;
}

// dct_dc_size + differential of an intra block (player.cpp:1010-1068); returns the DC (pixel scale)
__device__ __forceinline__ int parse_dc(SliceState& s, int blk)
{
    BitReader& br = s.br;
    const uint32_t bits = br.peek();
    int dc_size, used, dc;
    if (blk < 4) {
        dc = s.dc_y;
        if (!(bits >> 31)) { dc_size = 1 + (int)((bits >> 30) & 1); used = 2; }
        else if (!((bits >> 30) & 1)) { dc_size = ((bits >> 29) & 1) ? 3 : 0; used = 3; }
        else { int ones = min(9, __clz(~bits)); dc_size = ones + 2; used = dc_size - 1; }
    } else {
        dc = blk == 4 ? s.dc_cr : s.dc_cb;
        if (!(bits >> 31)) { dc_size = (int)((bits >> 30) & 1); used = 2; }
        else { int ones = min(10, __clz(~bits)); dc_size = ones + 1; used = min(dc_size, 10); }
    }
    br.skip(used);
    if (dc_size) {
        int delta = (int)br.get(dc_size);
        if (delta & (1 << (dc_size - 1))) dc += delta;
        else dc += (int)((0xFFFFFFFFu << dc_size) | (uint32_t)(delta + 1));
        if (blk < 4) s.dc_y = dc; else if (blk == 4) s.dc_cr = dc; else s.dc_cb = dc;
    }
    return dc;
}

}  // namespace

__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
ef_decode_kernel(const EfDev* __restrict__ Dp, int pic)
{
    extern __shared__ __align__(16) uint8_t smem[];
    SharedTables& T = *(SharedTables*)smem;
    const EfDev& D = *Dp;
    {   // stage the tables (the first sizeof(SharedTables) bytes of EfTables have the same layout)
        const uint32_t* src = (const uint32_t*)D.tables;
        uint32_t* dst = (uint32_t*)smem;
        for (int i = threadIdx.x; i < (int)(sizeof(SharedTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* wbase = smem + kTableBytes + (size_t)warp * kWarpBytes;
    uint32_t* list = (uint32_t*)wbase + (lane % kLanes) * kListEntries;
    uint8_t* hdr = wbase + kListBytes + (lane % kLanes) * kHdrStride;
    int* dense = (int*)(wbase + kListBytes + kHdrBytes);
    uint8_t* stage = wbase + kListBytes + kHdrBytes + kDenseBytes;
    uint64_t* bar = (uint64_t*)(stage + kStageBytes);
    const uint32_t sstage = smem_u32(stage), sbar = smem_u32(bar);
    uint32_t bar_phase = 0;
    if (lane == 0) mbar_init(bar, 1);
    uint32_t* ovf = D.k1_overflow + ((size_t)(blockIdx.x * kWarpsPerCta + warp) * 32 + lane) * (384 - kListEntries);
    for (int i = lane; i < kDenseBytes / 4; i += 32) ((uint32_t*)dense)[i] = 0;
    __syncthreads();

    const uint32_t total = D.pic_total[pic];
    const EfWork* work = D.work + D.pic_base[pic];
    uint32_t* cursor = D.cursor + pic;

    // per-lane constants of the reconstruction mapping
    const int col = lane & 7;                       // column owned in the IDCT column pass
    const int cblk = lane >> 3;                     // block (0..3) owned in the column pass
    const int prow = lane >> 1, phalf = lane & 1;   // luma pixel row / 8-pixel half owned for prediction + store
    const int rblk = (prow >> 3) * 2 + phalf;       // block that those pixels belong to
    const int rrow = prow & 7;                      // row of that block
    const int crow = lane & 7, cplane = (lane >> 3) & 1;   // chroma row / plane owned by lanes 0..15

    SliceState s;
    s.first = 0; s.cur = nullptr; s.ref = nullptr; s.qtab = T.qdef; s.mbw = 0;
    bool active = false, exhausted = false;
    // First round: slices are dealt out statically, 32 consecutive ones per warp, warps interleaved across
    // the SMs - with fewer slices than lanes (4,096 pictures x 12 slices vs 71 K lanes) every SM then runs
    // the same number of full warps instead of whichever warps reach the cursor first. Later rounds (lanes
    // whose slice ended) pull from the global cursor, which counts from the end of the first round.
#ifdef EF_K1_HALF
    const uint32_t first_round = gridDim.x * (uint32_t)kWarpsPerCta * 16u;
#else
    const uint32_t first_round = gridDim.x * (uint32_t)kWarpsPerCta * 32u;
#endif
    bool first_fill = true;
#ifdef EF_K1_HALF   // tuning experiment: only 16 slices per warp (twice the warps for the same batch)
    if (lane >= 16) exhausted = true;
#endif

    for (;;) {
        // ---- refill idle lanes with new slices -------------------------------------------------
        unsigned need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            uint32_t base;
            if (first_fill) {
#ifdef EF_K1_HALF
                base = ((uint32_t)warp * gridDim.x + blockIdx.x) * 16u;
#else
                base = ((uint32_t)warp * gridDim.x + blockIdx.x) * 32u;
#endif
                first_fill = false;
            } else {
                base = 0;
                int leader = __ffs(need) - 1;
                if (lane == leader) base = first_round + atomicAdd(cursor, (uint32_t)__popc(need));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
            }
            if (!active && !exhausted) {
                uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1));
                if (idx >= total) exhausted = true;
                else {
                    const EfWork w = work[idx];
                    const int code = w.info & 255;
                    s.ptype = (w.info >> 8) & 7; s.full_pel = (w.info >> 11) & 1; s.r_size = (w.info >> 12) & 7;
                    const EfSeq* seq = D.seq + (size_t)w.stream * (D.max_seq + 1) + (w.info >> 16);
                    s.mbw = min((int)seq->mb_width, EF_MBW_MAX);
                    s.mbh = min((int)seq->mb_height, EF_MBH_MAX);
                    s.qtab = seq->custom ? (const uint8_t*)seq->q_scan : (const uint8_t*)T.qdef;
                    const uint32_t fb = (D.base_pics[w.stream] + (uint32_t)pic + 1u) & 1u;     // flush_picture(), player.cpp:692
                    s.cur = D.frames + ef_frame_offset((int)w.stream, (int)fb);
                    s.ref = D.frames + ef_frame_offset((int)w.stream, (int)(fb ^ 1u));
                    s.br.init(D.es, D.es_off[w.stream] + w.es_off);
                    s.mb_y = code - 2; s.mb_x = s.mbw - 1;   // slice(), player.cpp:1255: the first increment lands on column 0 of row code-1
                    s.first = 1;
                    s.dc_y = s.dc_cr = s.dc_cb = 128; s.mv_h = s.mv_v = 0;
                    active = code >= 1 && code <= s.mbh && s.mbw > 0 && seq->valid;
                    if (active) {
                        s.qscale = (int)s.br.get(5);
                        while (s.br.get(1)) s.br.skip(8);    // extra_information_slice
                    }
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !active)) break;

        // ---- phase 1a: every lane parses the header of the next macroblock of its slice ---------
        bool have = false;
        int cbp_rem = 0, intra = 0;
        if (active) {
            have = parse_header(s, hdr, T, cbp_rem, intra);
            if (!have) active = false;
        }

        // ---- phase 1b: flat coefficient state machine, one VLC symbol per lane per step ----------
        int cnt = 0, n1mask = 0, abortmask = 0, blk = 0, n = 0;
        uint32_t* wptr = list;
        bool busy = have && cbp_rem != 0, start = true;
        const uint8_t* qrow = s.qtab + (intra ? 0 : 64);
        while (__any_sync(0xFFFFFFFFu, busy)) {
            if (busy) {
                BitReader& br = s.br;
                if (start) {                                   // next coded block of this macroblock
                    blk = __ffs(cbp_rem) - 1;
                    cbp_rem &= cbp_rem - 1;
                    n = 0;
                    if (intra) { ((int*)(hdr + kHdrDc))[blk] = parse_dc(s, blk); n = 1; }
                    start = false;
                }
                const uint32_t bits = br.peek();
                const int lz = min(__clz(bits), 12);               // row 12 / 25 = not a code
                // index = row * 32 + the five bits after the leading one; (bits << lz) >> 26 is "1xxxxx" = 32 + those bits
                const int ctx = n == 0 ? 13 * 32 - 32 : -32;       // first-coefficient context lives in rows 13..25
                const uint32_t e = T.dct[ctx + lz * 32 + (int)((bits << lz) >> 26)];
                int len = e & 31, run = (e >> 5) & 31, level = (int)(e >> 10);
                bool end_block = false, derail = false;
                if (level) {
                    if ((bits >> (32 - len)) & 1) level = -level;
                } else if (len == 2) {                         // '10': end of block (player.cpp:1075)
                    end_block = true;
                    if (n == 1) n1mask |= 1 << blk;            // Q5
                } else if (run == 1) {                         // escape: 6-bit run, 8- or 16-bit level (player.cpp:1092)
                    run = (int)((bits >> 20) & 63);
                    const int b = (int)((bits >> 12) & 255);
                    if (b == 0) { level = (int)((bits >> 4) & 255); len = 28; }
                    else if (b == 128) { level = (int)((bits >> 4) & 255) - 256; len = 28; }
                    else { level = (int)(int8_t)b; len = 20; }
                } else derail = true;                          // not a code: the reference derails here
                if (derail) {                                  // give up on this and the remaining blocks, end the slice
                    abortmask |= (1 << blk) | cbp_rem;
                    s.mb_y = s.mbh;
                    busy = false;
                } else {
                    br.skip(len);
                    if (!end_block) {
                        n += run;
                        if (n >= 64) { abortmask |= 1 << blk; end_block = true; }      // block() returns -1: nothing is stored
                        else {
                            const int v = dequant(level, intra, s.qscale * (int)qrow[n]);
                            const uint32_t zp = T.zp[n];                                   // zz = zig_zag[n]; b[zz] = v * scale_dct_q[zz] (player.cpp:1108, 1121)
                            const uint32_t ent = ((uint32_t)(v * (int)(zp >> 8)) & 0x3FFFFu) | ((zp & 63u) << 18) | ((uint32_t)blk << 24);
                            *wptr++ = ent;
                            if (++cnt == kListEntries) wptr = ovf;                 // spill the rest of a very long macroblock to HBM
                            n++;
                        }
                    }
                    if (end_block) { start = true; busy = cbp_rem != 0; }
                }
            }
        }
        if (have) {
            *(uint32_t*)(hdr + kHdrInfo) |= ((uint32_t)n1mask << 8) | ((uint32_t)abortmask << 14);
            *(uint32_t*)(hdr + kHdrCnt) |= (uint32_t)cnt;
        }
        unsigned todo = __ballot_sync(0xFFFFFFFFu, have);
        __syncwarp();

        // ---- phase 2: the warp reconstructs those macroblocks one by one -----------------------
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint8_t* H = wbase + kListBytes + r * kHdrStride;
            const uint32_t info = *(const uint32_t*)(H + kHdrInfo);
            const uint32_t cntw = *(const uint32_t*)(H + kHdrCnt);
            const uint32_t mvw = *(const uint32_t*)(H + kHdrMv);
            uint8_t* cur = (uint8_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)s.cur, r);
            const uint8_t* ref = (const uint8_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)s.ref, r);
            const uint32_t* rovf = (const uint32_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)ovf, r);
            const int mbw = __shfl_sync(0xFFFFFFFFu, s.mbw, r);
            const bool intra_r = (info >> 1) & 1;
            const int cbp = (info >> 2) & 63, n1m = (info >> 8) & 63, abm = (info >> 14) & 63;
            const int mx = (info >> 20) & 31, my = (info >> 25) & 15;
            const int entries = cntw & 0xFFFF, skip_before = cntw >> 16;
            const int live = cbp & ~abm;

            const int mvh = (int)(int16_t)(mvw & 0xFFFF), mvv = (int)(int16_t)(mvw >> 16);
            const int tile = ef_tile_offset(mx, my);

            // ---- prediction loads first: 8 luma pixels per lane, 8 chroma pixels for lanes 0..15 ---
            const int hx = mx * 32 + mvh, hy = my * 32 + mvv;                       // predict(), player.cpp:882
            const int cx = hx >> 1, cy = hy >> 1;                                   // Q3: floor
            const int lx = (hx >> 1) + phalf * 8, ly = (hy >> 1) + prow;
            const int kx = cx >> 1, ky = (cy >> 1) + crow;
            const int X0 = hx >> 1, Y0 = hy >> 1, tx0 = X0 >> 4, ty0 = Y0 >> 4;
            // whole prediction window inside the picture (always, for streams the reference accepts)
            const bool inside = hx >= 0 && hy >= 0 && X0 + 16 + (hx & 1) <= EF_W && Y0 + 16 + (hy & 1) <= EF_H;
            if (!intra_r && inside && lane == 0) {
                const bool two_x = ((X0 + 15 + (hx & 1)) >> 4) != tx0, two_y = ((Y0 + 15 + (hy & 1)) >> 4) != ty0;
                const uint8_t* src = ref + ef_tile_offset(tx0, ty0);
                // (the staging area is only ever written by these copies and read with plain loads that have
                // all completed before the __syncwarp() that ended the previous macroblock)
                const uint32_t row_bytes = two_x ? 2 * EF_TILE : EF_TILE;           // tiles of one row are contiguous in HBM
                mbar_expect_tx(sbar, row_bytes << (int)two_y);
                bulk_tiles(sstage, src, row_bytes, sbar);
                if (two_y) bulk_tiles(sstage + 2 * EF_TILE, src + EF_MBW_MAX * EF_TILE, row_bytes, sbar);
            }

            // expand the coefficient list into the dense scratch
            const uint32_t* rl = (const uint32_t*)wbase + r * kListEntries;
            for (int j = lane; j < entries; j += 32) {
                const uint32_t ent = j < kListEntries ? rl[j] : rovf[j - kListEntries];
                const int eb = (ent >> 24) & 7;
                if (!((abm >> eb) & 1)) dense[eb * kDenseStride + ((ent >> 18) & 63)] = ((int)(ent << 14)) >> 14;   // 18-bit signed value
            }

            // skipped macroblocks: predict_zero() copies them from the reference frame (player.cpp:1283-1288)
            if (skip_before) {
                int sx = mx, sy = my;
                for (int k = 0; k < skip_before; k++) {
                    if (--sx < 0) { sx = mbw - 1; sy--; }
                    if (sy < 0) break;
                    const int to = ef_tile_offset(sx, sy);
                    if (lane < 24) *(uint4*)(cur + to + lane * 16) = *(const uint4*)(ref + to + lane * 16);
                }
            }
            __syncwarp();                                                           // dense[] complete

            // ---- residual: luma set (blocks 0-3), then chroma set (blocks 4,5) ---------------------
            const int* dcs = (const int*)(H + kHdrDc);
            int resY[8], resC[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { resY[i] = 0; resC[i] = 0; }

#pragma unroll
            for (int set = 0; set < 2; set++) {
                const int setmask = set == 0 ? 0x0F : 0x30;
                if (!(live & setmask)) continue;                         // warp-uniform
                const int bk = set == 0 ? cblk : 4 + (cblk & 1);
                const bool lane_on = set == 0 || lane < 16;
                // column pass, in place: lane (block, column) owns the 8 words dense[block][0..7][column]
                if (lane_on) {
                    int* db = dense + bk * kDenseStride + col;
                    int v[8];
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) v[rr] = db[rr * 8];
                    if (intra_r && col == 0) v[0] = (int)((uint32_t)dcs[bk] << 8);      // b[0] <<= 8, player.cpp:1065
                    idct8<false>(v);
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) db[rr * 8] = v[rr];
                }
                __syncwarp();
                // row pass: luma lanes own (rblk, rrow); chroma lanes 0..15 own (4 + lane/8, lane%8)
                const int orow = set == 0 ? rrow : crow;
                const int oblk = set == 0 ? rblk : 4 + cplane;
                const bool ocoded = lane_on && (live >> oblk) & 1;
                int w[8];
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = 0;
                if (lane_on) {
                    int4* rowp = (int4*)(dense + oblk * kDenseStride + orow * 8);
                    const int4 a = rowp[0], b = rowp[1];
                    rowp[0] = make_int4(0, 0, 0, 0); rowp[1] = make_int4(0, 0, 0, 0);    // leave the scratch zeroed for the next macroblock
                    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                }
                if (ocoded) {
                    if (!((n1m >> oblk) & 1)) idct8<true>(w);
                    else {                                               // n == 1: dc = b[0] >> 8 (Q5); after the column pass every row holds b[0] in column 0
                        const int dc = intra_r ? dcs[oblk] : w[0] >> 8;
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] = dc;
                    }
                }
                if (set == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) resY[i] = w[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) resC[i] = w[i];
                }
                __syncwarp();
            }

            // ---- finish the prediction, combine + store (copy_block / copy_block_dc / add_block / add_block_dc)
            uint32_t py0 = 0, py1 = 0, pc0 = 0, pc1 = 0;
            if (!intra_r) {
                PredWords wy, wc;
                if (inside) {
                    mbar_wait(sbar, bar_phase);
                    bar_phase ^= 1;
                    pred_words_staged<true>(stage, 0, lx, ly, tx0, ty0, hy & 1, wy);
                    if (lane < 16) pred_words_staged<false>(stage, cplane, kx, ky, tx0, ty0, cy & 1, wc);
                } else {
                    pred_words_clamped<true>(ref, 0, lx, ly, hy & 1, wy);
                    if (lane < 16) pred_words_clamped<false>(ref, cplane, kx, ky, cy & 1, wc);
                }
                pred_finish(wy, lx, hx & 1, hy & 1, py0, py1);
                if (lane < 16) pred_finish(wc, kx, cx & 1, cy & 1, pc0, pc1);
            }
            {
                const bool coded = (cbp >> rblk) & 1, aborted = (abm >> rblk) & 1, n1 = (n1m >> rblk) & 1;
                uint32_t o0 = py0, o1 = py1;
                bool store = true;
                if (coded && !aborted) {
                    if (intra_r && n1) {                                 // copy_block_dc: replicated, not clamped (Q7)
                        uint32_t d = (uint32_t)resY[0]; d |= d << 8; d |= d << 16;
                        o0 = o1 = d;
                    } else {
                        o0 = pin4(py0, resY[0], resY[1], resY[2], resY[3]);
                        o1 = pin4(py1, resY[4], resY[5], resY[6], resY[7]);
                    }
                } else if (intra_r) store = false;                        // aborted intra block: destination untouched
                if (store) *(uint2*)(cur + tile + prow * 16 + phalf * 8) = make_uint2(o0, o1);
            }
            if (lane < 16) {
                const int bk = 4 + cplane;
                const bool coded = (cbp >> bk) & 1, aborted = (abm >> bk) & 1, n1 = (n1m >> bk) & 1;
                uint32_t o0 = pc0, o1 = pc1;
                bool store = true;
                if (coded && !aborted) {
                    if (intra_r && n1) {
                        uint32_t d = (uint32_t)resC[0]; d |= d << 8; d |= d << 16;
                        o0 = o1 = d;
                    } else {
                        o0 = pin4(pc0, resC[0], resC[1], resC[2], resC[3]);
                        o1 = pin4(pc1, resC[4], resC[5], resC[6], resC[7]);
                    }
                } else if (intra_r) store = false;
                if (store) *(uint2*)(cur + tile + 256 + cplane * 64 + crow * 8) = make_uint2(o0, o1);
            }

            __syncwarp();
        }
    }
}

// host-side launch helper ------------------------------------------------------------------------
size_t ef_decode_smem_bytes() { return (size_t)kTableBytes + (size_t)kWarpsPerCta * kWarpBytes; }
int ef_decode_threads() { return kWarpsPerCta * 32; }

cudaError_t ef_decode_configure()
{
    return cudaFuncSetAttribute(ef_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ef_decode_smem_bytes());
}

cudaError_t ef_launch_decode(const EfDev* dev, int pic, int ctas, cudaStream_t stream)
{
    ef_decode_kernel<<<ctas, kWarpsPerCta * 32, ef_decode_smem_bytes(), stream>>>(dev, pic);
    return cudaGetLastError();
}
