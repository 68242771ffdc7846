// host_capi.cpp -- C entry points over the host layer (block cutter, .arc container) so that the
// CPU test-suite can drive it through ctypes without a GPU.  libseqarc_host.so; not used by the CLI.
#include <cstring>
#include <stdexcept>

#include "arc_file.h"
#include "fastq_block.h"
#include "id_process.h"

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
                  const uint64_t *src1, const uint64_t *src2, int nblocks, const uint8_t *ref_md5, int comment, const uint8_t *longflags) {
    ArcWriter w;
    if (!w.open(path)) return -1;
    uint64_t o = 0, so[2] = {0, 0};
    for (int i = 0; i < nblocks; i++) {
        uint64_t sb[2] = {src1[i], src2 ? src2[i] : 0};
        w.add_block(recs + o, lens[i], sb, so, longflags && longflags[i]);
        o += lens[i];
        so[0] += sb[0]; so[1] += sb[1];
    }
    ArcParams p;
    p.pe = pe; p.usemodel = usemodel; p.lossy = lossy; p.md5 = md5;
    p.name1 = name1 ? name1 : ""; p.name2 = name2 ? name2 : "";
    if (idbin) memcpy(p.idbin, idbin, 512);
    if (model) p.model.assign(model, model + modellen);
    if (ref_md5) { p.has_ref = true; memcpy(p.ref_md5, ref_md5, 16); }     // archive written against a reference index
    p.comment = comment != 0;                                              // '+' lines repeat the ID
    return w.finish(p) ? 0 : -1;
}

// SeqArcPreProcess's ID-bin analysis over the head sample blocks of the input (sample blocks: whole records until
// BlockSize (SE) / BlockSize/2 per mate (PE) bytes, at most five).  Fills idbin[512]; returns idbin[0].
int sqh_analyse_ids(const char *p1, const char *p2, uint64_t block_bytes, uint8_t *idbin) {
    try {
        bool pe = p2 && *p2;
        InFile f1, f2;
        if (!f1.open(p1) || (pe && !f2.open(p2))) return -1;
        uint64_t quota = block_bytes / (pe ? 2 : 1);
        auto record = [&](InFile &f, std::vector<char> &t) {
            for (int k = 0; k < 4; k++) if (!f.getline(t)) return false;
            return true;
        };
        memset(idbin, 0, 512);
        bool ok = true;
        for (int b = 0; b < 5 && ok; b++) {
            std::vector<char> t1, t2;
            while (t1.size() < quota) { if (!record(f1, t1)) break; if (pe && !record(f2, t2)) break; }
            if (t1.empty()) break;
            MemBuf mb;
            parse_block(t1, t2, pe, mb);
            ok = analyse_id_bins(mb, pe, idbin);
        }
        return idbin[0];
    } catch (const std::exception &) {
        return -2;
    }
}

// IDProcess::decodeIDS on one tag-5 payload; returns the number of ID bytes written (<= cap) or -1.
long sqh_decode_ids(const uint8_t *payload, uint64_t paylen, const uint8_t *idbin, const int32_t *seqlen, uint32_t nreads,
                    uint8_t *ids, uint64_t cap, uint16_t *idlen) {
    std::vector<uint8_t> o;
    std::vector<uint16_t> l;
    if (!decode_ids(payload, paylen, idbin, seqlen, nreads, o, l) || o.size() > cap || l.size() != nreads) return -1;
    memcpy(ids, o.data(), o.size());
    memcpy(idlen, l.data(), 2 * l.size());
    return (long)o.size();
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
