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
//     macroblock per iteration, straight out of HBM/L2 (MSB-first 32-bit windows, one funnel
//     shift per peek, CLZ-indexed shared-memory tables: one load per symbol).
//   * the parsed macroblock (<= 6 x 64 quantised levels + header) is left in shared memory; then
//     the whole WARP reconstructs the 32 macroblocks one after another: dequantise + the
//     reference's integer AAN IDCT with one lane per block column/row (4 luma blocks = 32 lanes),
//     half-pel motion compensation and the clamped add with one lane per 8-pixel row segment,
//     8-byte coalescing-friendly stores into the striped frame store.
//   * lanes that finish a slice pull the next one from a global cursor, so lane occupancy stays
//     high until the picture's work list is empty (persistent CTAs, one per SM).
// Bit-exactness notes: Q1 clamp [0,248]; Q2 oddification maps 0 -> +1; Q3 chroma
// vector = floor(luma position / 2); Q4 matrices indexed in raster order (done at index time);
// Q5 single-coefficient blocks bypass the IDCT with floor; Q6 first macroblock of a slice lands in
// column 0; Q7 intra DC-only blocks are replicated unclamped.
#include "ef_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRecStride = 804;                 // bytes per lane record; 201 words -> conflict-free lane-strided access
constexpr int kRecCoef = 0;                     // int16 [6][64] quantised levels, zig-zag order, stored as 2*level+1 (0 = none)
constexpr int kRecDc = 768;                     // int32 [6] intra DC (pixel scale)
constexpr int kRecInfo = 792;                   // bit0 valid, bit1 intra, 2-7 coded blocks, 8-13 n==1 mask, 14-19 abort mask (bit b = block b), 20-24 qscale
constexpr int kRecPos = 796;                    // mb_addr | skip_before << 16
constexpr int kRecMv = 800;                     // (int16 h) | (int16 v) << 16, half-pel units
constexpr int kScratchWords = 4 * 72;           // IDCT transpose scratch: 4 blocks x (64 + 8 pad) ints
constexpr int kWarpBytes = 32 * kRecStride + kScratchWords * 4;

struct SharedTables {
    uint16_t dct[12 * 32];
    uint16_t mba[8 * 32];
    uint16_t mv[7 * 32];
    uint16_t cbp[512];
    uint8_t ptype[64];
};

// ---------------------------------------------------------------------------------------------
// bit reader (FILL_BITS/peek_bits/get_bits, player.cpp:348-352, 495-514): MSB-first. `hi` holds
// the current 32-bit word, `lo` the next one, `nx` the one after (prefetched), pos = bits of `hi`
// already consumed. peek() is a single funnel shift.
// ---------------------------------------------------------------------------------------------
struct BitReader {
    const uint32_t* p;
    const uint32_t* end;
    uint32_t hi, lo, nx;
    int pos;

    __device__ __forceinline__ uint32_t fetch()
    {
        uint32_t v = 0;
        if (p < end) v = __byte_perm(__ldg(p), 0, 0x0123);
        p++;
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* base, const uint8_t* stop)
    {
        uintptr_t a = (uintptr_t)base;
        p = (const uint32_t*)(a & ~(uintptr_t)3);
        end = (const uint32_t*)(((uintptr_t)stop + 3) & ~(uintptr_t)3);
        pos = (int)(a & 3) * 8;
        hi = fetch(); lo = fetch(); nx = fetch();
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, pos); }
    __device__ __forceinline__ void skip(int n)
    {
        pos += n;
        if (pos >= 32) { pos -= 32; hi = lo; lo = nx; nx = fetch(); }
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
    uint32_t stream;
    uint8_t* cur;
    const uint8_t* ref;
    const EfSeq* seq;
    int mbw, mbn;            // macroblocks per row / per picture
    int mb_addr;             // linear address of the last macroblock handled
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
    int m0 = v[0];
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
        for (int i = 0; i < 8; i++) v[i] = (v[i] + 128) >> 8;
    }
}

// dequantise one stored level (block(), player.cpp:1110-1121). s = 2*level+1, never 0 here.
__device__ __forceinline__ int dequant(int s, bool intra, int qscale, int q, int prescale)
{
    int v = s - 1;                                  // level << 1
    if (!intra) v += v < 0 ? -1 : 1;
    v = (v * qscale * q) / 16;                      // C division: truncates toward zero
    if ((v & 1) == 0) v -= v > 0 ? 1 : -1;          // Q2: 0 becomes +1
    v = max(-2048, min(2047, v));
    return v * prescale;
}

__device__ __forceinline__ uint32_t pin4(uint32_t pred, int r0, int r1, int r2, int r3)
{
    // PIN(b + s) for four pixels (add_block, player.cpp:1189; _pin clamps to [0,248], Q1)
    int p0 = min(248, max(0, (int)(pred & 0xFF) + r0));
    int p1 = min(248, max(0, (int)((pred >> 8) & 0xFF) + r1));
    int p2 = min(248, max(0, (int)((pred >> 16) & 0xFF) + r2));
    int p3 = min(248, max(0, (int)(pred >> 24) + r3));
    return (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16) | ((uint32_t)p3 << 24);
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

// Eight predicted pixels starting at byte offset `off` of the reference frame with half-pel
// flags; `off2` is the offset of the row below (used when yh). Plain byte semantics of mocomp().
__device__ __forceinline__ void predict8(const uint8_t* ref, int off, int off2, int xh, int yh, uint32_t& o0, uint32_t& o1)
{
    off = max(0, min(EF_FRAME - 4, off));
    off2 = max(0, min(EF_FRAME - 4, off2));
    const uint32_t* a = (const uint32_t*)(ref + (off & ~3));
    int sh = (off & 3) * 8;
    uint32_t w0 = a[0], w1 = a[1], w2 = a[2];
    uint32_t p0 = __funnelshift_r(w0, w1, sh), p1 = __funnelshift_r(w1, w2, sh);
    if (xh) {
        uint32_t q0 = __funnelshift_rc(w0, w1, sh + 8), q1 = __funnelshift_rc(w1, w2, sh + 8);
        if (yh) {
            const uint32_t* b = (const uint32_t*)(ref + (off2 & ~3));
            int sh2 = (off2 & 3) * 8;
            uint32_t v0 = b[0], v1 = b[1], v2 = b[2];
            o0 = avg4x4(p0, q0, __funnelshift_r(v0, v1, sh2), __funnelshift_rc(v0, v1, sh2 + 8));
            o1 = avg4x4(p1, q1, __funnelshift_r(v1, v2, sh2), __funnelshift_rc(v1, v2, sh2 + 8));
        } else {
            o0 = avg2x4(p0, q0);
            o1 = avg2x4(p1, q1);
        }
    } else if (yh) {
        const uint32_t* b = (const uint32_t*)(ref + (off2 & ~3));
        int sh2 = (off2 & 3) * 8;
        uint32_t v0 = b[0], v1 = b[1], v2 = b[2];
        o0 = avg2x4(p0, __funnelshift_r(v0, v1, sh2));
        o1 = avg2x4(p1, __funnelshift_r(v1, v2, sh2));
    } else {
        o0 = p0; o1 = p1;
    }
}

// chroma row address inside a frame (Frame::get_cr / get_cb, player.cpp:38-46); plane 0 = block 4
__device__ __forceinline__ int chroma_row_off(int plane, int y) { return (y >> 3) * 8448 + ((y & 7) + plane * 8) * 528 + 352; }

// ---------------------------------------------------------------------------------------------
// parse one macroblock of this lane's slice into its record. Returns false when the slice ended.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool parse_macroblock(SliceState& s, uint8_t* rec, const SharedTables& T)
{
    BitReader& br = s.br;
    uint32_t bits = br.peek();
    if ((bits >> 9) == 0) return false;             // slice_done(): next 23 bits are zero (player.cpp:1238)

    // macroblock_address_increment (player.cpp:1267-1275)
    int increment = 0;
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
    if (s.first) { s.first = 0; s.mb_addr += 1; }   // inc_mb ignores its argument for the first macroblock (Q6)
    else {
        if (increment > 1) { s.dc_y = s.dc_cr = s.dc_cb = 128; s.mv_h = s.mv_v = 0; skip_before = increment - 1; }
        s.mb_addr += increment;
    }
    if (s.mb_addr >= s.mbn) return false;           // the reference would write past the frame here
    const int mb_addr_here = s.mb_addr;

    // macroblock_type (player.cpp:1292)
    int mb_type;
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
    int intra = mb_type & 1;
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

    cbp = (int)(__brev((unsigned)cbp) >> 26);       // from here on bit b = block b (the VLC value has block 0 in bit 5)
    int n1mask = 0, abortmask = 0;
    bool derailed = false;
    int16_t* coef = (int16_t*)(rec + kRecCoef);
    int* dcs = (int*)(rec + kRecDc);
    for (int blk = 0; blk < 6; blk++) {
        if (!((cbp >> blk) & 1)) continue;
        int16_t* c = coef + blk * 64;
        int n = 0;
        if (intra) {                                 // dct_dc_size + differential (player.cpp:1010-1068)
            bits = br.peek();
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
            dcs[blk] = dc;
            n = 1;
        }
        for (;;) {                                   // AC coefficients (player.cpp:1070-1122)
            bits = br.peek();
            int run, level, used;
            if (bits >> 31) {
                if (n) {
                    if (!((bits >> 30) & 1)) { br.skip(2); break; }     // '10' end of block
                    used = 3; level = ((bits >> 29) & 1) ? -1 : 1;       // '11s'
                } else { used = 2; level = ((bits >> 30) & 1) ? -1 : 1; }   // '1s' first coefficient
                run = 0;
            } else {
                int lz = __clz(bits);
                if (lz == 5) {                       // escape: 6-bit run, 8- or 16-bit level (player.cpp:1092)
                    run = (int)((bits >> 20) & 63);
                    int b = (int)((bits >> 12) & 255);
                    if (b == 0) { level = (int)((bits >> 4) & 255); used = 28; }
                    else if (b == 128) { level = (int)((bits >> 4) & 255) - 256; used = 28; }
                    else { level = (int)(int8_t)b; used = 20; }
                } else {
                    if (lz > 11) { derailed = true; break; }             // not a code: the reference derails here
                    uint32_t e = T.dct[lz * 32 + ((bits << (lz + 1)) >> 27)];
                    int len = e & 31;
                    if (!len) { derailed = true; break; }
                    run = (int)((e >> 5) & 31);
                    level = (int)(e >> 10);
                    if ((bits >> (31 - len)) & 1) level = -level;
                    used = len + 1;
                }
            }
            br.skip(used);
            n += run;
            if (n >= 64) { abortmask |= 1 << blk; break; }              // block() returns -1: nothing is stored
            c[n++] = (int16_t)(2 * level + 1);
        }
        if (derailed) {                              // give up on this and the remaining blocks, end the slice
            abortmask |= 0x3F & ~((1 << blk) - 1);
            s.mb_addr = s.mbn;
            break;
        }
        if (n == 1) n1mask |= 1 << blk;
    }
    uint32_t info = 1u | ((uint32_t)intra << 1) | ((uint32_t)cbp << 2) | ((uint32_t)n1mask << 8) |
                    ((uint32_t)abortmask << 14) | ((uint32_t)(s.qscale & 31) << 20);
    *(uint32_t*)(rec + kRecInfo) = info;
    *(uint32_t*)(rec + kRecPos) = (uint32_t)mb_addr_here | ((uint32_t)skip_before << 16);
    *(uint32_t*)(rec + kRecMv) = ((uint32_t)mvh & 0xFFFFu) | ((uint32_t)mvv << 16);
    return true;   // a derailed macroblock is still emitted; the slice then ends at the next call (mb_addr == mbn)
}

}  // namespace

__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
ef_decode_kernel(const EfDev* __restrict__ Dp, int pic)
{
    extern __shared__ __align__(16) uint8_t smem[];
    SharedTables& T = *(SharedTables*)smem;
    const EfDev& D = *Dp;
    {   // stage the VLC tables (the first sizeof(SharedTables) bytes of EfTables have the same layout)
        const uint32_t* src = (const uint32_t*)D.tables;
        uint32_t* dst = (uint32_t*)smem;
        for (int i = threadIdx.x; i < (int)(sizeof(SharedTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* wbase = smem + ((sizeof(SharedTables) + 15) & ~15) + (size_t)warp * kWarpBytes;
    uint8_t* rec = wbase + lane * kRecStride;
    int* scratch = (int*)(wbase + 32 * kRecStride);
    for (int i = 0; i < kRecStride / 4; i++) ((uint32_t*)rec)[i] = 0;
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
    uint8_t izz[8], psc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) { izz[r] = D.tables->izz[r * 8 + col]; psc[r] = D.tables->prescale[r * 8 + col]; }

    SliceState s;
    s.first = 0; s.stream = 0; s.cur = nullptr; s.ref = nullptr; s.seq = nullptr;
    bool active = false, exhausted = false;

    for (;;) {
        // ---- refill idle lanes with new slices -------------------------------------------------
        unsigned need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            uint32_t base = 0;
            int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1));
                if (idx >= total) exhausted = true;
                else {
                    EfWork w = work[idx];
                    const int code = w.info & 255;
                    s.stream = w.stream;
                    s.ptype = (w.info >> 8) & 7; s.full_pel = (w.info >> 11) & 1; s.r_size = (w.info >> 12) & 7;
                    s.seq = D.seq + (size_t)w.stream * (D.max_seq + 1) + (w.info >> 16);
                    s.mbw = min((int)s.seq->mb_width, EF_MBW_MAX);
                    int mbh = min((int)s.seq->mb_height, EF_MBH_MAX);
                    s.mbn = s.mbw * mbh;
                    const uint32_t fb = (D.base_pics[w.stream] + (uint32_t)pic + 1u) & 1u;     // flush_picture(), player.cpp:692
                    s.cur = D.frames + ef_frame_offset((int)w.stream, (int)fb);
                    s.ref = D.frames + ef_frame_offset((int)w.stream, (int)(fb ^ 1u));
                    const uint8_t* es = D.es + D.es_off[w.stream];
                    const uint8_t* stop = D.es + D.es_off[w.stream + 1];
                    s.br.init(es + w.es_off, stop);
                    s.mb_addr = (code - 1) * s.mbw - 1;      // slice(): row = code-1, first increment lands on column 0
                    s.first = 1;
                    s.dc_y = s.dc_cr = s.dc_cb = 128; s.mv_h = s.mv_v = 0;
                    active = code >= 1 && code <= mbh && s.mbw > 0 && s.seq->valid;
                    if (active) {
                        s.qscale = (int)s.br.get(5);
                        while (s.br.get(1)) s.br.skip(8);    // extra_information_slice
                    }
                }
            }
        }
        if (__all_sync(0xFFFFFFFFu, !active)) break;

        // ---- phase 1: every lane parses one macroblock of its own slice -------------------------
        bool have = false;
        if (active) {
            have = parse_macroblock(s, rec, T);
            if (!have) active = false;
        }
        unsigned todo = __ballot_sync(0xFFFFFFFFu, have);
        __syncwarp();

        // ---- phase 2: the warp reconstructs those macroblocks one by one -----------------------
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            uint8_t* R = wbase + r * kRecStride;
            const uint32_t info = *(const uint32_t*)(R + kRecInfo);
            const uint32_t posw = *(const uint32_t*)(R + kRecPos);
            const uint32_t mvw = *(const uint32_t*)(R + kRecMv);
            uint8_t* cur = (uint8_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)s.cur, r);
            const uint8_t* ref = (const uint8_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)s.ref, r);
            const EfSeq* seq = (const EfSeq*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)s.seq, r);
            const int mbw = __shfl_sync(0xFFFFFFFFu, s.mbw, r);
            const bool intra = (info >> 1) & 1;
            const int cbp = (info >> 2) & 63, n1mask = (info >> 8) & 63, abortmask = (info >> 14) & 63;
            const int qscale = (info >> 20) & 31;
            const int mb_addr = posw & 0xFFFF, skip_before = posw >> 16;

            // skipped macroblocks: predict_zero() copies them from the reference frame (player.cpp:1283-1288)
            for (int k = skip_before; k > 0; k--) {
                const int a = mb_addr - k;
                const int mx = a % mbw, my = a / mbw;
                const int yo = (my * 16 + prow) * EF_STRIDE + mx * 16 + phalf * 8;
                *(uint2*)(cur + yo) = *(const uint2*)(ref + yo);
                if (lane < 16) {
                    const int co = (my * 16 + lane) * EF_STRIDE + EF_W + mx * 8;    // strip rows 0-7 block 4, 8-15 block 5
                    *(uint2*)(cur + co) = *(const uint2*)(ref + co);
                }
            }

            const int mx = mb_addr % mbw, my = mb_addr / mbw;
            const int mvh = (int)(int16_t)(mvw & 0xFFFF), mvv = (int)(int16_t)(mvw >> 16);

            // ---- prediction: 8 luma pixels per lane, 8 chroma pixels for lanes 0..15 ------------
            uint32_t py0 = 0, py1 = 0, pc0 = 0, pc1 = 0;
            const int crow = lane & 7, cplane = (lane >> 3) & 1;
            if (!intra) {
                const int hx = mx * 32 + mvh, hy = my * 32 + mvv;
                const int yo = ((hy >> 1) + prow) * EF_STRIDE + (hx >> 1) + phalf * 8;
                predict8(ref, yo, yo + EF_STRIDE, hx & 1, hy & 1, py0, py1);
                if (lane < 16) {
                    const int cx = hx >> 1, cy = hy >> 1;                           // Q3: floor
                    const int y0 = (cy >> 1) + crow;
                    const int o1 = chroma_row_off(cplane, y0) + (cx >> 1);
                    const int o2 = chroma_row_off(cplane, y0 + 1) + (cx >> 1);
                    predict8(ref, o1, o2, cx & 1, cy & 1, pc0, pc1);
                }
            }

            // ---- residual: luma set (blocks 0-3, cbp bits 5..2), then chroma set (blocks 4,5) ---
            const int16_t* coef = (const int16_t*)(R + kRecCoef);
            const int* dcs = (const int*)(R + kRecDc);
            int resY[8], resC[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { resY[i] = 0; resC[i] = 0; }

#pragma unroll
            for (int set = 0; set < 2; set++) {
                const int setmask = set == 0 ? 0x0F : 0x30;
                if (!(cbp & setmask & ~abortmask)) continue;            // warp-uniform
                const int blk = set == 0 ? cblk : 4 + (cblk & 1);
                const bool lane_on = set == 0 || lane < 16;
                const bool coded = lane_on && ((cbp & ~abortmask) >> blk) & 1;
                const bool full = coded && !((n1mask >> blk) & 1);
                int v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = 0;
                if (full) {
                    const uint2 qv = *(const uint2*)((intra ? seq->intra_qT : seq->inter_qT) + col * 8);
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) {
                        const int sv = coef[blk * 64 + izz[rr]];
                        const int q = (int)(((rr < 4 ? qv.x : qv.y) >> ((rr & 3) * 8)) & 255);
                        if (sv) v[rr] = dequant(sv, intra, qscale, q, psc[rr]);
                    }
                    if (intra && col == 0) v[0] = (int)((uint32_t)dcs[blk] << 8);   // b[0] <<= 8, player.cpp:1065
                    idct8<false>(v);
                }
                int* Tb = scratch + cblk * 72;
                if (lane_on) {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) Tb[rr * 8 + col] = v[rr];
                }
                __syncwarp();
                // row pass: luma lanes own (rblk, rrow); chroma lanes 0..15 own (4 + lane/8, lane%8)
                const int ob = set == 0 ? rblk : (lane >> 3) & 1;          // scratch slot of the row this lane owns
                const int orow = set == 0 ? rrow : (lane & 7);
                const int oblk = set == 0 ? rblk : 4 + ((lane >> 3) & 1);
                const bool ocoded = lane_on && ((cbp & ~abortmask) >> oblk) & 1;
                const bool ofull = ocoded && !((n1mask >> oblk) & 1);
                int w[8];
                {
                    const int4 a = *(const int4*)(scratch + ob * 72 + orow * 8);
                    const int4 b = *(const int4*)(scratch + ob * 72 + orow * 8 + 4);
                    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                }
                __syncwarp();
                if (ofull) idct8<true>(w);
                else if (ocoded) {                                       // n == 1: dc = b[0] >> 8 (Q5)
                    int dc;
                    if (intra) dc = dcs[oblk];
                    else {
                        const int sv = coef[oblk * 64];
                        const int q = (int)seq->inter_qT[0];
                        dc = dequant(sv, false, qscale, q, 32) >> 8;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = dc;
                }
                if (set == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) resY[i] = w[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) resC[i] = w[i];
                }
            }

            // ---- combine + store (copy_block / copy_block_dc / add_block / add_block_dc) --------
            {
                const int blk = rblk;
                const bool coded = (cbp >> blk) & 1, aborted = (abortmask >> blk) & 1, n1 = (n1mask >> blk) & 1;
                uint32_t o0 = py0, o1 = py1;
                bool store = true;
                if (coded && !aborted) {
                    if (intra && n1) {                                   // copy_block_dc: replicated, not clamped (Q7)
                        uint32_t d = (uint32_t)resY[0]; d |= d << 8; d |= d << 16;
                        o0 = o1 = d;
                    } else {
                        o0 = pin4(py0, resY[0], resY[1], resY[2], resY[3]);
                        o1 = pin4(py1, resY[4], resY[5], resY[6], resY[7]);
                    }
                } else if (intra) store = false;                          // aborted intra block: destination untouched
                if (store) *(uint2*)(cur + (my * 16 + prow) * EF_STRIDE + mx * 16 + phalf * 8) = make_uint2(o0, o1);
            }
            if (lane < 16) {
                const int blk = 4 + cplane;
                const bool coded = (cbp >> blk) & 1, aborted = (abortmask >> blk) & 1, n1 = (n1mask >> blk) & 1;
                uint32_t o0 = pc0, o1 = pc1;
                bool store = true;
                if (coded && !aborted) {
                    if (intra && n1) {
                        uint32_t d = (uint32_t)resC[0]; d |= d << 8; d |= d << 16;
                        o0 = o1 = d;
                    } else {
                        o0 = pin4(pc0, resC[0], resC[1], resC[2], resC[3]);
                        o1 = pin4(pc1, resC[4], resC[5], resC[6], resC[7]);
                    }
                } else if (intra) store = false;
                if (store) *(uint2*)(cur + (my * 16 + cplane * 8 + crow) * EF_STRIDE + EF_W + mx * 8) = make_uint2(o0, o1);
            }

            // ---- clear the coefficient slots this macroblock used (parser invariant: all zero) --
#pragma unroll
            for (int b = 0; b < 6; b++)
                if ((cbp >> b) & 1) ((uint32_t*)(R + kRecCoef))[b * 32 + lane] = 0;
            __syncwarp();
        }
    }
}

// host-side launch helper ------------------------------------------------------------------------
size_t ef_decode_smem_bytes() { return ((sizeof(SharedTables) + 15) & ~(size_t)15) + (size_t)kWarpsPerCta * kWarpBytes; }
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
