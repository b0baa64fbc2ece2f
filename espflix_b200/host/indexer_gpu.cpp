// espflix_b200/host/indexer_gpu.cpp — make_index / pts2seq / merge_index with the reference's signatures
// (indexer/indexer.cpp:90-253) on top of the C-ABI. See ef_indexer.h.
#include "ef_indexer.h"

#include <stdio.h>
#include <string.h>

#include <stdexcept>

#include "espflix_b200.h"

static_assert(sizeof(idx_rec) == 32 && sizeof(idx_hdr) == 104, "video.idx layout (indexer.cpp:22-36)");

static int g_index_device = 0;
void ef_indexer_set_device(int device) { g_index_device = device; }

void make_index(const std::string& src, std::vector<idx>& idxs)
{
    FILE* f = fopen(src.c_str(), "rb");
    if (!f) throw std::runtime_error("make_index: cannot open " + src);
    std::vector<uint8_t> ts;
    uint8_t buf[1024 * 188];                       // the reference reads in blocks of 1,024 packets too
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) ts.insert(ts.end(), buf, buf + n);
    fclose(f);
    ts.resize(ts.size() / 188 * 188);              // a trailing partial packet carries nothing
    printf(">%s index\n", src.c_str());
    const uint64_t off[2] = { 0, ts.size() };
    const size_t np = ts.size() / 188;
    std::vector<int64_t> pts(np ? np : 1);
    std::vector<uint32_t> pos(np ? np : 1);
    ef_tsidx_info info;
    uint8_t none = 0;
    if (ef_tsidx_scan(g_index_device, ts.empty() ? &none : ts.data(), off, 1, 90000 / 12, &info, pts.data(), pos.data()) != EF_OK)
        throw std::runtime_error(std::string("make_index: ") + ef_last_error());
    idx id;
    for (uint32_t i = 0; i < info.n_seq; i++) id.seqs.push_back({ pts[i], pos[i] });
    id.first_pts = info.first_pts;
    id.last_pts = info.last_pts;
    idxs.push_back(id);
}

idx_rec pts2seq(idx& id, int speedx, int bin_size)
{
    std::vector<int64_t> pts(id.seqs.size());
    std::vector<uint32_t> pos(id.seqs.size());
    for (size_t i = 0; i < id.seqs.size(); i++) { pts[i] = id.seqs[i].pts; pos[i] = id.seqs[i].pos188; }
    uint32_t n = 0;
    if (ef_tsidx_samples(g_index_device, pts.data(), pos.data(), (int)pts.size(), id.first_pts, id.last_pts, (uint32_t)bin_size, nullptr, 0, &n) != EF_OK && n == 0)
        throw std::runtime_error(std::string("pts2seq: ") + ef_last_error());
    id.samples.assign(n, 0);
    if (n && ef_tsidx_samples(g_index_device, pts.data(), pos.data(), (int)pts.size(), id.first_pts, id.last_pts, (uint32_t)bin_size, id.samples.data(), n, &n) != EF_OK)
        throw std::runtime_error(std::string("pts2seq: ") + ef_last_error());
    printf("%d->%d, produced %d from %d\n", (int)id.first_pts, (int)id.last_pts, (int)id.samples.size(), (int)id.seqs.size());
    idx_rec r;
    memset(&r, 0, sizeof(r));
    r.first_pts = id.first_pts;
    r.last_pts = id.last_pts;
    r.sample_count = (uint32_t)id.samples.size();
    r.bin_size = (uint32_t)bin_size;
    r.trick_speed = (uint32_t)speedx;
    return r;
}

void merge_index(std::vector<idx>& all, const std::string& path)
{
    const int speedx = 15;
    idx_hdr hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.sig = ('I' << 0) | ('D' << 8) | ('X' << 16);
    hdr.len = 3;
    hdr.video = pts2seq(all[0], 1, 90000 / 12);
    hdr.fwd = pts2seq(all[1], speedx, 90000 / 12);
    hdr.rev = pts2seq(all[2], speedx, 90000 / 12);
    const std::string p = path + "/video.idx";
    FILE* f = fopen(p.c_str(), "wb");
    if (!f) throw std::runtime_error("merge_index: cannot write " + p);
    fwrite(&hdr, 1, sizeof(hdr), f);
    for (int k = 0; k < 3; k++)
        if (!all[k].samples.empty()) fwrite(all[k].samples.data(), 1, 4 * all[k].samples.size(), f);
    fclose(f);
}
