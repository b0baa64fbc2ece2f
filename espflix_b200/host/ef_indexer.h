// espflix_b200/host/ef_indexer.h — host-side mirror of the reference's index builder interface
// (indexer/indexer.cpp:22-36, 77-88, 90, 209, 230): same type and function names, same arguments; the scans
// run on the GPU through the C-ABI (ef_tsidx_scan / ef_tsidx_samples). ffmpeg orchestration (the rest of the
// reference tool) is out of scope.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

typedef struct {            // indexer.cpp:22
    int64_t first_pts;
    int64_t last_pts;
    uint32_t bin_size;
    uint32_t trick_speed;
    uint32_t sample_count;
} idx_rec;

typedef struct {            // indexer.cpp:30
    uint32_t sig;
    uint32_t len;           // 3
    idx_rec video;
    idx_rec fwd;
    idx_rec rev;
} idx_hdr;

typedef struct {            // indexer.cpp:78: one sequence header: its PES pts and the TS packet it starts in
    int64_t pts;
    uint32_t pos188;
} seq;

typedef struct {            // indexer.cpp:83
    std::vector<seq> seqs;
    int64_t first_pts;
    int64_t last_pts;
    std::vector<uint32_t> samples;
} idx;

// scan one transport stream for its sequence headers (indexer.cpp:90); appends one idx to idxs. Throws std::runtime_error
// when the file cannot be read or the GPU call fails (the reference would crash on fopen failure).
void make_index(const std::string& src, std::vector<idx>& idxs);
// indexer.cpp:209: fills id.samples, returns the header record
idx_rec pts2seq(idx& id, int speedx, int bin_size);
// indexer.cpp:230: all = {video, fwd, rev}; writes <path>/video.idx (struct padding zeroed)
void merge_index(std::vector<idx>& all, const std::string& path);
