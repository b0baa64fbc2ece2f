// oracle/ref_indexer_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference index builder (/root/reference/indexer/indexer.cpp: make_index :90-187,
// pts2seq :209-228, merge_index :230-253) into oracle/_ref/libefref_idx.so. The reference file is a
// command-line tool; it is compiled where it lies through the include below with its main() renamed, so its
// static helpers (parse, parse_pts, pts2pos) are exercised exactly as shipped. SURVEY.md §8f-4.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define main efref_indexer_main
#include "indexer.cpp"          // -I$(REF)/indexer
#undef main

extern "C" {

// make_index() of one transport stream file: the sequence-header table (pts, TS packet number), first/last
// video pts. Returns the number of sequence headers (copies at most `cap`).
int efref_make_index(const char* path, int64_t* pts, uint32_t* pos188, int cap, int64_t* first_pts, int64_t* last_pts)
{
    vector<idx> all;
    make_index(path, all);
    const idx& id = all[0];
    int n = (int)id.seqs.size();
    for (int i = 0; i < n && i < cap; i++) { pts[i] = id.seqs[i].pts; pos188[i] = id.seqs[i].pos188; }
    *first_pts = id.first_pts; *last_pts = id.last_pts;
    return n;
}

// the whole tool minus ffmpeg: three streams (main, fast-forward, rewind) -> <out_dir>/video.idx
int efref_build_idx(const char* video_ts, const char* fwd_ts, const char* rev_ts, const char* out_dir)
{
    vector<idx> all;
    make_index(video_ts, all);
    make_index(fwd_ts, all);
    make_index(rev_ts, all);
    merge_index(all, out_dir);
    return 0;
}

}  // extern "C"
