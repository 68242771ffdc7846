// host_capi.cpp -- C entry points over the host layer (block cutter, .arc container) so that the
// CPU test-suite can drive it through ctypes without a GPU.  libseqarc_host.so; not used by the CLI.
#include <cstring>
#include <stdexcept>

#include "arc_file.h"
#include "fastq_block.h"

using namespace sqhost;

extern "C" {

// Cut the input into blocks exactly as the reference's reader does; per block: nreads, src bytes of
// file 1 / 2.  Returns the number of blocks (<= cap) or -1.
int sqh_cut_blocks(const char *p1, const char *p2, uint64_t block_bytes, uint32_t *nreads, uint64_t *src1, uint64_t *src2, int cap) {
    try {
        BlockReader rd(p1, p2 ? p2 : "", block_bytes);
        if (!rd.ok()) return -1;
        MemBuf mb;
        int n = 0;
        while (rd.next(mb)) {
            if (n >= cap) return -1;
            nreads[n] = mb.nreads(); src1[n] = mb.src_bytes[0]; src2[n] = mb.src_bytes[1];
            n++;
        }
        return n;
    } catch (const std::exception &) {
        return -2;
    }
}

// Write an archive from ready-made block records (concatenated in `recs`, lengths in `lens`).
int sqh_write_arc(const char *path, int pe, int usemodel, int lossy, int md5, const char *name1, const char *name2,
                  const uint8_t *idbin, const uint8_t *model, uint64_t modellen, const uint8_t *recs, const uint64_t *lens,
                  const uint64_t *src1, const uint64_t *src2, int nblocks) {
    ArcWriter w;
    if (!w.open(path)) return -1;
    uint64_t o = 0, so[2] = {0, 0};
    for (int i = 0; i < nblocks; i++) {
        uint64_t sb[2] = {src1[i], src2 ? src2[i] : 0};
        w.add_block(recs + o, lens[i], sb, so);
        o += lens[i];
        so[0] += sb[0]; so[1] += sb[1];
    }
    ArcParams p;
    p.pe = pe; p.usemodel = usemodel; p.lossy = lossy; p.md5 = md5;
    p.name1 = name1 ? name1 : ""; p.name2 = name2 ? name2 : "";
    if (idbin) memcpy(p.idbin, idbin, 512);
    if (model) p.model.assign(model, model + modellen);
    return w.finish(p) ? 0 : -1;
}

// Parse an archive; returns the number of blocks and fills flags = pe | usemodel<<1 | lossy<<2 | md5<<3 | idrule<<4.
int sqh_read_arc(const char *path, int *flags, uint64_t *arc_off, uint32_t *len, int cap) {
    ArcReader r;
    if (!r.load(path)) return -1;
    *flags = (int)r.p.pe | (r.p.usemodel << 1) | (r.p.lossy << 2) | (r.p.md5 << 3) | ((r.p.idbin[0] != 0) << 4);
    int n = (int)r.idx.size();
    for (int i = 0; i < n && i < cap; i++) { arc_off[i] = r.idx[i].arc_off; len[i] = r.idx[i].len; }
    return n;
}
}
