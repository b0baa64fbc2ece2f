// espflix_b200/host/ef_indexer.h — host-side mirror of the reference's index builder interface
// (indexer/indexer.cpp:22-36, 77-88, 90, 209, 230): same type and function names, same arguments; the scans
// run on the GPU through the C-ABI (ef_tsidx_scan / ef_tsidx_samples). ffmpeg orchestration (the rest of the
// reference tool) is out of scope.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

// Layout-compatible with the reference's records (indexer.cpp:22-36, 77-88): video.idx is these structs written raw.
struct idx_rec {
    int64_t first_pts, last_pts;                        // pts of the first sequence header / of the last video PES start
    uint32_t bin_size, trick_speed, sample_count;       // ticks per sample, playback speed of the stream, samples that follow
};

struct idx_hdr {
    uint32_t sig, len;                                  // 'I','D','X',0 and the number of records (3)
    idx_rec video, fwd, rev;                            // main stream, fast-forward and rewind trick streams
};

struct seq {                                            // one sequence header:
    int64_t pts;                                        //   PES pts of the packet it starts in
    uint32_t pos188;                                    //   that packet's number in the file
};

struct idx {                                            // everything known about one transport stream
    std::vector<seq> seqs;
    int64_t first_pts, last_pts;
    std::vector<uint32_t> samples;                      // filled by pts2seq()
};

// scan one transport stream for its sequence headers (indexer.cpp:90); appends one idx to idxs. Throws std::runtime_error
// when the file cannot be read or the GPU call fails (the reference would crash on fopen failure).
void make_index(const std::string& src, std::vector<idx>& idxs);
// indexer.cpp:209: fills id.samples, returns the header record
idx_rec pts2seq(idx& id, int speedx, int bin_size);
// indexer.cpp:230: all = {video, fwd, rev}; writes <path>/video.idx (struct padding zeroed)
void merge_index(std::vector<idx>& all, const std::string& path);
