#include "arc_file.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

namespace sqhost {

static void put_vint(std::vector<uint8_t> &o, uint64_t v, int n) {      // value with an explicit width
    uint64_t x = v | ((uint64_t)1 << (7 * n));
    for (int i = n - 1; i >= 0; i--) o.push_back((uint8_t)(x >> (8 * i)));
}
static void put_id(std::vector<uint8_t> &o, uint64_t id) {              // Encap::setID: smallest width
    int n = 1;
    while (n < 8 && id > (((uint64_t)1 << (7 * n)) - 2)) n++;
    put_vint(o, id, n);
}
static void put_param(std::vector<uint8_t> &o, int id, const void *v, size_t n) {
    put_id(o, id);
    put_vint(o, n, n < 127 ? 1 : 2);
    o.insert(o.end(), (const uint8_t *)v, (const uint8_t *)v + n);
}
static void le(std::vector<uint8_t> &o, uint64_t v, int n) { for (int i = 0; i < n; i++) o.push_back((uint8_t)(v >> (8 * i))); }

bool ArcWriter::open(const std::string &path) {
    f_ = fopen(path.c_str(), "wb");
    if (!f_) return false;
    const uint8_t hdr[16] = {'.', 'a', 'r', 'c', 1, 6, 0, 0x82, 1, 0, 0, 0, 0, 0, 0, 0};   // TAG 2 size patched in finish()
    fwrite(hdr, 1, 16, f_);
    off_ = 16;
    return true;
}

void ArcWriter::add_block(const uint8_t *rec, uint64_t len, const uint64_t sb[2], const uint64_t so[2], bool longflag) {
    BlockIndex b;
    b.len = (uint32_t)len;
    b.longflag = longflag;
    b.src_bytes[0] = (uint32_t)sb[0]; b.src_bytes[1] = (uint32_t)sb[1];
    b.arc_off = off_;
    b.src_off[0] = so[0]; b.src_off[1] = so[1];
    idx_.push_back(b);
    fwrite(rec, 1, len, f_);
    off_ += len;
}

bool ArcWriter::finish(const ArcParams &p) {
    std::vector<uint8_t> par;
    auto b1 = [&](int id, bool v) { uint8_t x = v; put_param(par, id, &x, 1); };
    auto u16 = [&](int id, uint16_t v) { uint8_t x[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; put_param(par, id, x, 2); };
    auto u32 = [&](int id, uint32_t v) { uint8_t x[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; put_param(par, id, x, 4); };
    b1(1, !p.has_ref);                   // 1 = no reference index
    b1(2, !p.comment);                   // 1 = bare '+' lines
    b1(3, !p.pe);                        // single-end
    b1(4, p.gz); b1(5, true);
    b1(6, p.usemodel);
    b1(7, false);
    u16(8, 1); u16(9, 6);                // format version 1.6
    u16(10, 0);                          // max insert (aligned mode)
    u32(11, (uint32_t)idx_.size());      // number of blocks
    u32(12, 0);
    put_param(par, 13, p.name1.data(), p.name1.size());
    if (p.pe) put_param(par, 14, p.name2.data(), p.name2.size());
    put_param(par, 15, p.idbin, 512);
    b1(16, p.lossy); b1(17, p.md5); b1(18, false);

    std::vector<uint8_t> tr;
    put_id(tr, 1); put_vint(tr, par.size(), 2);
    tr.insert(tr.end(), par.begin(), par.end());
    if (p.usemodel) {
        put_id(tr, 6); put_vint(tr, p.model.size() + 4, 4);
        const uint8_t h[4] = {0x81, 0x82, 0x01, 0x00};
        tr.insert(tr.end(), h, h + 4);
        tr.insert(tr.end(), p.model.begin(), p.model.end());
    }
    if (p.has_ref) {
        put_id(tr, 8); put_vint(tr, 16, 1);
        tr.insert(tr.end(), p.ref_md5, p.ref_md5 + 16);
    }
    std::vector<uint8_t> ix;
    for (auto &b : idx_) {
        le(ix, (uint64_t)b.len * 2 + (b.longflag ? 1 : 0), 4);    // low bit: long-read flag
        if (p.pe) { le(ix, b.src_bytes[0], 4); le(ix, b.src_bytes[1], 4); }
        else { le(ix, p.comment ? 0 : 1, 4); le(ix, b.src_bytes[0], 4); }
        le(ix, 0, 4);
        le(ix, b.arc_off, 8);
        le(ix, b.src_off[0], 8);
        if (p.pe) le(ix, b.src_off[1], 8);
    }
    put_id(tr, 7); put_vint(tr, ix.size(), 4);
    tr.insert(tr.end(), ix.begin(), ix.end());

    std::vector<uint8_t> out;
    put_id(out, 3); put_vint(out, tr.size(), 4);
    out.insert(out.end(), tr.begin(), tr.end());
    fwrite(out.data(), 1, out.size(), f_);
    // patch TAG 2's 8-byte size = bytes of all block records
    std::vector<uint8_t> sz;
    put_vint(sz, off_ - 16, 8);
    fseek(f_, 8, SEEK_SET);
    fwrite(sz.data(), 1, 8, f_);
    off_ += out.size();
    bool ok = fclose(f_) == 0;
    f_ = nullptr;
    return ok;
}

static int get_vint(const uint8_t *p, const uint8_t *end, uint64_t *v) {
    if (p >= end) return 0;
    int n = 1;
    uint8_t m = 0x80;
    while (n <= 8 && !(p[0] & m)) { n++; m >>= 1; }
    if (n > 8 || p + n > end) return 0;
    uint64_t x = p[0] & (m - 1);
    for (int i = 1; i < n; i++) x = (x << 8) | p[i];
    *v = x;
    return n;
}

ArcReader::~ArcReader() {
    if (base && size) munmap((void *)base, size);
}

bool ArcReader::load(const std::string &path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); err = "cannot stat " + path; return false; }
    long n = (long)st.st_size;
    if (n > 0) {
        void *m = mmap(nullptr, (size_t)n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); err = "cannot map " + path; return false; }
        base = (const uint8_t *)m;
        size = (size_t)n;
    }
    ::close(fd);
    const uint8_t *d = base, *end = d + size;
    if (n < 16 || memcmp(d, ".arc", 4) != 0 || d[4] != 1 || d[5] != 6) { err = "not a SeqArc v1.6 archive"; return false; }
    uint64_t tag, sz;
    const uint8_t *q = d + 7;
    int k = get_vint(q, end, &tag); if (!k || tag != 2) { err = "bad container"; return false; } q += k;
    k = get_vint(q, end, &sz); if (!k || q + k + sz > end) { err = "bad container size"; return false; } q += k;
    q += sz;
    k = get_vint(q, end, &tag); if (!k || tag != 3) { err = "no trailer"; return false; } q += k;
    k = get_vint(q, end, &sz); if (!k || q + k + sz > end) { err = "bad trailer size"; return false; } q += k;
    const uint8_t *te = q + sz;
    std::vector<uint8_t> ix;
    while (q < te) {
        k = get_vint(q, te, &tag); if (!k) break; q += k;
        k = get_vint(q, te, &sz); if (!k || q + k + sz > te) { err = "bad trailer entry"; return false; } q += k;
        const uint8_t *pay = q;
        q += sz;
        if (tag == 1) {
            const uint8_t *a = pay, *ae = pay + sz;
            while (a < ae) {
                uint64_t id, ps;
                k = get_vint(a, ae, &id); if (!k) break; a += k;
                k = get_vint(a, ae, &ps); if (!k || a + k + ps > ae) break; a += k;
                if (id == 1 && ps == 1) p.has_ref = !a[0];
                else if (id == 2 && ps == 1) p.comment = !a[0];
                else if (id == 3 && ps == 1) p.pe = !a[0];
                else if (id == 4 && ps == 1) p.gz = a[0];
                else if (id == 6 && ps == 1) p.usemodel = a[0];
                else if (id == 13) p.name1.assign((const char *)a, ps);
                else if (id == 14) p.name2.assign((const char *)a, ps);
                else if (id == 15 && ps == 512) memcpy(p.idbin, a, 512);
                else if (id == 16 && ps == 1) p.lossy = a[0];
                else if (id == 17 && ps == 1) p.md5 = a[0];
                a += ps;
            }
        } else if (tag == 6) {
            if (sz < 4) { err = "bad model entry"; return false; }
            p.model.assign(pay + 4, pay + sz);
        } else if (tag == 8 && sz == 16) {
            memcpy(p.ref_md5, pay, 16);
        } else if (tag == 7) {
            ix.assign(pay, pay + sz);
        }
    }
    size_t rec = p.pe ? 40 : 32;
    auto rd = [&](const uint8_t *x, int nb) { uint64_t v = 0; for (int i = 0; i < nb; i++) v |= (uint64_t)x[i] << (8 * i); return v; };
    for (size_t o = 0; o + rec <= ix.size(); o += rec) {
        const uint8_t *x = ix.data() + o;
        BlockIndex b;
        b.len = (uint32_t)(rd(x, 4) >> 1);
        b.longflag = rd(x, 4) & 1;
        if (p.pe) { b.src_bytes[0] = (uint32_t)rd(x + 4, 4); b.src_bytes[1] = (uint32_t)rd(x + 8, 4); }
        else b.src_bytes[0] = (uint32_t)rd(x + 8, 4);
        b.arc_off = rd(x + 16, 8);
        b.src_off[0] = rd(x + 24, 8);
        if (p.pe) b.src_off[1] = rd(x + 32, 8);
        if (b.arc_off + b.len > size) { err = "block index points outside the file"; return false; }
        idx.push_back(b);
    }
    return true;
}

}  // namespace sqhost
