// ef_indexer_cli video.ts fwd.ts rev.ts out_dir — the index half of the reference tool's main()
// (indexer/indexer.cpp:318-333) without the ffmpeg calls: three make_index() + merge_index().
#include <stdio.h>

#include <exception>

#include "ef_indexer.h"

int main(int argc, char** argv)
{
    if (argc != 5) { fprintf(stderr, "usage: %s video.ts fwd.ts rev.ts out_dir\n", argv[0]); return 2; }
    try {
        std::vector<idx> all;
        make_index(argv[1], all);
        make_index(argv[2], all);
        make_index(argv[3], all);
        merge_index(all, argv[4]);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
