// espflix_b200/csrc/ef_audio.cu — the audio half of a transport-stream program, batched over streams.
//
// Replaces, for every stream at once (paths under /root/reference):
//   MpegDecoder::demux() for PID 0x101 / 0x102 -> push_audio()      src/player.cpp:381-432, src/video.cpp:1007
//   decode_audio() -> sbc_decoder(): get_samples(), bit_allocation(), IQUANT(), synthesize8()
//                                                                    src/video.cpp:964-986, src/sbc_decoder.cpp:70-373
//   write_pcm_16() -> pdm_second_order()                             espflix.ino:73-136
// The reference decodes one 128-sample frame at a time through a 170-word ring shared by 16 sliding windows. The
// ring is only a delay line: V_t[i] = (sum_j matrix[i][j] * S_t[j]) >> 15 for block t, and
//   pcm_t[i] = clip((sum_{d=0..9} window[d][i] * V_{t-d}[d even ? i : (i + 8) & 15]) >> 15),
// so every frame - and every block - is independent once the V rows are in HBM:
//   ef_sbc_probe_kernel   one thread per stream: the frame size decode_audio() learns from the first 64 bytes
//   ef_sbc_matrix_kernel  one warp per frame: header, scale factors, bit allocation (12.6.3), 128 samples cut out of
//                         the bit field at computed offsets (no serial bit reader), IQUANT, matrixing -> V[16][16]
//   ef_sbc_window_kernel  one thread per PCM sample: 10 taps over the V rows of this and the nine blocks before
//   ef_pdm_kernel         one thread per stream: the second-order delta-sigma modulator is a serial non-linear
//                         recurrence (32 one-bit samples per PCM sample); streams run side by side
//   ef_audio_ts_*         TS -> audio bytes on the device (one thread per packet; the "PES without PTS mutes the
//                         stream" state of demux() is a running maximum over PES starts)
// Quirks kept: decode_audio() decodes the first frame twice (once to learn the frame size, video.cpp:971), so the
// filter memory already holds it when the real decode starts; a frame it rejects (bad sync byte, joint stereo,
// 4 subbands) re-synthesises the previous frame's samples; 32-bit wrap-around of the accumulators. Domain: mono,
// 8 subbands, 16 blocks (write_pcm_16(mono,128,1) passes 128 samples whatever the header says); a frame size that
// does not divide the reference's 4 KB ring makes it read past the ring on the straddling frame (undefined there;
// here the stream is simply linear).
#include "ef_common.cuh"
#include "ef_sbc_tables.h"

namespace {

__constant__ int c_matrix[16][8];
__constant__ int c_window[10][8];
__constant__ signed char c_offset8[4][8];

// header + scale factors + bit allocation of one frame (get_samples / bit_allocation, sbc_decoder.cpp:141-305).
// Returns the sum of bits over the 8 subbands, or -1 for a frame the reference rejects, -2 outside its domain.
__device__ int sbc_frame_bits(const uint8_t* d, uint64_t avail, int* bits, int* sf)
{
    if (avail < 4 || d[0] != 0x9C) return -1;
    const uint32_t b1 = d[1];
    const int frequency = (b1 >> 6) & 3, blocks = 4 * (((b1 >> 4) & 3) + 1), mode = (b1 >> 2) & 3;
    const int allocation = (b1 >> 1) & 1, subbands = (b1 & 1) ? 8 : 4, bitpool = avail > 2 ? d[2] : 0;
    if (mode == 3 || subbands == 4) return -1;
    if (mode != 0 || blocks != 16) return -2;
    int bitneed[8], max_bitneed = 0;
#pragma unroll
    for (int sb = 0; sb < 8; sb++) {
        const uint32_t byte = 4 + (sb >> 1) < (int)avail ? d[4 + (sb >> 1)] : 0u;
        const int s = (sb & 1) ? (byte & 15) : (byte >> 4);
        sf[sb] = s;
        int need;
        if (allocation) need = s;
        else if (s == 0) need = -5;
        else { need = s - c_offset8[frequency][sb]; if (need > 0) need /= 2; }
        bitneed[sb] = need;
        max_bitneed = max(max_bitneed, need);
    }
    int bitcount = 0, slicecount = 0, bitslice = max_bitneed + 1;
    do {
        bitslice--;
        bitcount += slicecount;
        slicecount = 0;
#pragma unroll
        for (int sb = 0; sb < 8; sb++) {
            if (bitneed[sb] > bitslice + 1 && bitneed[sb] < bitslice + 16) slicecount++;
            else if (bitneed[sb] == bitslice + 1) slicecount += 2;
        }
    } while (bitcount + slicecount < bitpool);
    if (bitcount + slicecount == bitpool) { bitcount += slicecount; bitslice--; }
    int total = 0;
#pragma unroll
    for (int sb = 0; sb < 8; sb++) bits[sb] = bitneed[sb] < bitslice + 2 ? 0 : min(bitneed[sb] - bitslice, 16);
#pragma unroll
    for (int sb = 0; sb < 8; sb++) {
        if (bitcount < bitpool) {
            if (bits[sb] >= 2 && bits[sb] < 16) { bits[sb]++; bitcount++; }
            else if (bitneed[sb] == bitslice + 1 && bitpool > bitcount + 1) { bits[sb] = 2; bitcount += 2; }
        }
    }
#pragma unroll
    for (int sb = 0; sb < 8; sb++) {
        if (bitcount < bitpool && bits[sb] < 16) { bits[sb]++; bitcount++; }
        total += bits[sb];
    }
    return total;
}

}  // namespace

cudaError_t ef_audio_upload_constants()
{
    cudaError_t e = cudaMemcpyToSymbol(c_matrix, ef_sbc_matrix, sizeof(ef_sbc_matrix));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_window, ef_sbc_window, sizeof(ef_sbc_window));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_offset8, ef_sbc_offset8, sizeof(ef_sbc_offset8));
}

// info[s]: frame size in bytes (0: empty stream, -1: first frame rejected, -2: outside the domain)
__global__ void ef_sbc_probe_kernel(const uint8_t* __restrict__ es, const uint64_t* __restrict__ off, int n_streams, int* __restrict__ frame_size)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const uint64_t len = off[s + 1] - off[s];
    int fs = 0;
    if (len) {
        int bits[8], sf[8];
        const int total = sbc_frame_bits(es + off[s], len < 64 ? len : 64, bits, sf);     // decode_audio() hands sbc_decoder 64 bytes (video.cpp:971)
        fs = total < 0 ? total : 8 + 2 * total;            // the lazy byte loader has consumed header + scale factors + 16 * total bits
    }
    frame_size[s] = fs;
}

// One warp per frame slot f = 0 .. n_frames of a stream: slot 0 is the probe decode of frame 0 (video.cpp:971), slot
// f >= 1 is frame f - 1. V rows go to vrows[(row_base[s] + 16 f + blk) * 16 + i].
__global__ void __launch_bounds__(128)
ef_sbc_matrix_kernel(const uint8_t* __restrict__ es, const uint64_t* __restrict__ off, const int* __restrict__ frame_size,
                     const uint64_t* __restrict__ slot_base /* [n_streams + 1] frame slots before stream s */, int n_streams, int32_t* __restrict__ vrows)
{
    __shared__ int32_t sb_s[4][16][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t slot = (uint64_t)blockIdx.x * 4 + w;
    const uint64_t total_slots = slot_base[n_streams];
    if (slot < total_slots) {                              // (no early return: the warp reaches __syncwarp below either way)
        int lo = 0, hi = n_streams;                        // stream of this slot
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (slot_base[mid] <= slot) lo = mid; else hi = mid; }
        const int s = lo;
        const int fs = frame_size[s];
        const uint64_t base = off[s], len = off[s + 1] - off[s];
        long f = (long)(slot - slot_base[s]) - 1;          // -1 = the probe
        const bool probe = f < 0;
        if (probe) f = 0;
        // a rejected frame re-synthesises the samples of the last accepted one (sb_sample[] is simply left as it was)
        int bits[8], sf[8], total = -1;
        const uint8_t* d = nullptr;
        uint64_t avail = 0;
        for (long g = f; g >= 0 && total < 0; g--) {
            d = es + base + (uint64_t)g * (uint64_t)fs;
            avail = len - (uint64_t)g * (uint64_t)fs;     // the bit loader may run on into the next frame; bytes behind the stream read as 0
            total = sbc_frame_bits(d, avail, bits, sf);
            if (probe) break;
        }
        // lane -> block lane/2, subbands 4 (lane & 1) .. + 3
        const int blk = lane >> 1, sb0 = (lane & 1) * 4;
        int pre = 0;
#pragma unroll
        for (int sb = 0; sb < 8; sb++) if (sb < sb0) pre += bits[sb];
        uint32_t bitpos = (uint32_t)(blk * max(total, 0) + pre);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int sb = sb0 + k, level = total < 0 ? 0 : bits[sb];
            int32_t sample = 0;
            if (level) {
                const uint64_t byte = 8 + (bitpos >> 3);
                uint32_t win = 0;                          // 24 bits starting at that byte cover <= 7 + 16 bits
#pragma unroll
                for (int q = 0; q < 3; q++) win = (win << 8) | (byte + q < avail ? d[byte + q] : 0u);
                const uint32_t raw = (win >> (24 - (bitpos & 7) - level)) & ((1u << level) - 1);
                sample = (int32_t)((uint32_t)((raw << 1) | 1) << sf[sb]) / (int32_t)((1u << level) - 1);   // IQUANT, sbc_decoder.cpp:262
                sample -= 1 << sf[sb];
                bitpos += (uint32_t)level;
            }
            sb_s[w][blk][sb] = sample;
        }
    }
    __syncwarp();
    if (slot < total_slots) {
        // matrixing: lane -> block lane/2, outputs 8 (lane & 1) .. + 7
        const int blk = lane >> 1, i0 = (lane & 1) * 8;
        int32_t src[8];
#pragma unroll
        for (int j = 0; j < 8; j++) src[j] = sb_s[w][blk][j];
        int32_t* out = vrows + (slot * 16 + blk) * 16 + i0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) acc += (uint32_t)c_matrix[i0 + i][j] * (uint32_t)src[j];
            out[i] = (int32_t)acc >> 15;
        }
    }
}

// one thread per PCM sample; t = block index inside the stream's slot rows (16 probe blocks first)
__global__ void ef_sbc_window_kernel(const int32_t* __restrict__ vrows, const uint64_t* __restrict__ slot_base, const uint64_t* __restrict__ pcm_off,
                                     int n_streams, int16_t* __restrict__ pcm)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pcm_off[n_streams]) return;
    int lo = 0, hi = n_streams;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pcm_off[mid] <= k) lo = mid; else hi = mid; }
    const int s = lo;
    const uint64_t local = k - pcm_off[s];
    const int i = (int)(local & 7);
    const uint64_t t = 16 + (local >> 3);                  // block index incl. the 16 probe blocks
    const int32_t* rows = vrows + slot_base[s] * 16 * 16;
    uint32_t acc = 0;
#pragma unroll
    for (int d = 0; d < 10; d++) {
        const int32_t v = t >= (uint64_t)d ? rows[(t - d) * 16 + ((d & 1) ? ((i + 8) & 15) : i)] : 0;
        acc += (uint32_t)v * (uint32_t)c_window[d][i];
    }
    int32_t v = (int32_t)acc >> 15;
    v = max(-0x7FFF, min(0x7FFF, v));
    pcm[k] = (int16_t)v;
}

// pdm_second_order (espflix.ino:73-107) over the whole PCM of a stream, modulator state starting at zero
__global__ void ef_pdm_kernel(const int16_t* __restrict__ pcm, const uint64_t* __restrict__ pcm_off, int n_streams, uint16_t* __restrict__ pdm)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const int32_t a1 = (int32_t)(0x7FFF * 1.18940), a2 = (int32_t)(0x7FFF * 2.12340);
    int32_t i0 = 0, i1 = 0, i2 = 0;
    uint32_t b = 0;
    const uint64_t p0 = pcm_off[s], p1 = pcm_off[s + 1];
    for (uint64_t k = p0; k < p1; k++) {
        const int32_t smp = (int32_t)pcm[k] * 2;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            i0 = (i0 + smp) >> 1;                          // low pass
#pragma unroll 4
            for (int j = 0; j < 16; j++) {
                b <<= 1;
                if (i2 >= 0) { i1 += i0 - a1 - (i2 >> 7); i2 += i1 - a2; b |= 1; }
                else { i1 += i0 + a1 - (i2 >> 7); i2 += i1 + a2; }
            }
            pdm[2 * k + h] = (uint16_t)b;
        }
    }
}

// ---- TS -> audio bytes --------------------------------------------------------------------------------------
// per packet: payload start (0 = none) and length of PID 0x101 / 0x102, and for PES starts whether the PTS parses
__global__ void ef_audio_ts_packet_kernel(const uint8_t* __restrict__ ts, uint64_t n_packets, uint8_t* __restrict__ start, uint8_t* __restrict__ kind)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_packets) return;
    const uint8_t* d = ts + k * 188;
    uint8_t st = 0, kd = 0;                                // kind: 0 continuation, 1 PES start with PTS, 2 PES start without
    const int pid = ((d[1] << 8) | d[2]) & 0x1FFF;
    if (d[0] == 0x47 && (pid == 0x101 || pid == 0x102) && (d[3] & 0x10)) {
        int o = 4;
        if (d[3] & 0x20) o = 5 + d[4];
        if (d[1] & 0x40) {
            kd = 2;
            if (o + 9 <= 188) {
                const int flags = (d[o + 6] << 8) | d[o + 7];
                const int p = o + 9;
                if ((flags & 0x80) && p < 188 && (d[p] & 0xF0) == ((flags >> 2) & 0x30)) kd = 1;   // parse_pts(), player.cpp:299
                o = o + 9 + d[o + 8];
            } else o = 188;
        }
        if (o < 188) st = (uint8_t)o;
    }
    start[k] = st;
    kind[k] = kd;                                          // a packet without the sync byte is skipped (more(), player.cpp:476-479)
}

// one thread per file walks its packets in order (a few thousand): byte offsets of the payloads that reach push_audio()
__global__ void ef_audio_ts_scan_kernel(const uint64_t* __restrict__ pkt_off, int n_files, const uint8_t* __restrict__ start, const uint8_t* __restrict__ kind,
                                        uint32_t* __restrict__ out_pos /* per packet, 0xFFFFFFFF = dropped */, uint64_t* __restrict__ es_len)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_files) return;
    bool open = false;                                     // _audio_pts != -1
    uint64_t out = 0;
    for (uint64_t k = pkt_off[f]; k < pkt_off[f + 1]; k++) {
        const uint32_t kd = kind[k];
        if (kd == 1) open = true; else if (kd == 2) open = false;
        const uint32_t st = start[k];
        if (open && st) { out_pos[k] = (uint32_t)out; out += 188 - st; }
        else out_pos[k] = 0xFFFFFFFFu;
    }
    es_len[f] = out;
}

__global__ void ef_audio_ts_copy_kernel(const uint8_t* __restrict__ ts, const uint64_t* __restrict__ pkt_off, int n_files, uint64_t n_packets,
                                        const uint8_t* __restrict__ start, const uint32_t* __restrict__ out_pos, const uint64_t* __restrict__ es_off, uint8_t* __restrict__ es)
{
    const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_packets || out_pos[k] == 0xFFFFFFFFu) return;
    int lo = 0, hi = n_files;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pkt_off[mid] <= k) lo = mid; else hi = mid; }
    const uint8_t* d = ts + k * 188;
    const int st = start[k];
    uint8_t* dst = es + es_off[lo] + out_pos[k];
    for (int i = st + lane; i < 188; i += 32) dst[i - st] = d[i];
}
