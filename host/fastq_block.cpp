#include "fastq_block.h"

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>

#include <cstring>
#include <stdexcept>

namespace sqhost {

bool InFile::open(const std::string &path) {
    close();
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    unsigned char magic[2] = {0, 0};
    ssize_t k = ::pread(fd, magic, 2, 0);
    if (k == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {       // gzip member
        ::close(fd);
        gzFile g = gzopen(path.c_str(), "rb");
        if (g) gzbuffer(g, 1 << 20);
        f_ = g;
        return f_ != nullptr;
    }
    fd_ = fd;                                                    // plain: read(2) straight into the caller's buffer
    return true;
}
void InFile::close() {
    if (f_) gzclose((gzFile)f_);
    if (fd_ >= 0) ::close(fd_);
    f_ = nullptr;
    fd_ = -1;
    gbuf_.clear(); gpos_ = 0;
}
size_t InFile::read(void *dst, size_t n) {
    size_t got = 0;
    if (fd_ >= 0) {
        if (gpos_ < gbuf_.size()) {                              // bytes getline() read ahead
            got = std::min(n, gbuf_.size() - gpos_);
            memcpy(dst, gbuf_.data() + gpos_, got);
            gpos_ += got;
        }
        while (got < n) {
            ssize_t k = ::read(fd_, (char *)dst + got, n - got);
            if (k <= 0) break;
            got += (size_t)k;
        }
        return got;
    }
    while (got < n) {
        unsigned chunk = (unsigned)std::min<size_t>(n - got, 1u << 30);
        int r = gzread((gzFile)f_, (char *)dst + got, chunk);
        if (r <= 0) break;
        got += (size_t)r;
    }
    return got;
}
bool InFile::getline(std::vector<char> &out) {
    if (fd_ >= 0) {
        bool any = false;
        for (;;) {
            if (gpos_ >= gbuf_.size()) {
                gbuf_.resize(1 << 20);
                ssize_t k = ::read(fd_, gbuf_.data(), gbuf_.size());
                gbuf_.resize(k > 0 ? (size_t)k : 0);
                gpos_ = 0;
                if (gbuf_.empty()) return any;
            }
            const char *p = gbuf_.data() + gpos_, *e = gbuf_.data() + gbuf_.size();
            const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *stop = q ? q + 1 : e;
            out.insert(out.end(), p, stop);
            gpos_ += (size_t)(stop - p);
            any = true;
            if (q) return true;
        }
    }
    char buf[1 << 16];
    bool any = false;
    while (gzgets((gzFile)f_, buf, sizeof buf) != nullptr) {
        size_t n = strlen(buf);
        out.insert(out.end(), buf, buf + n);
        any = true;
        if (n && buf[n - 1] == '\n') break;
    }
    return any;
}

BlockReader::BlockReader(const std::string &p1, const std::string &p2, uint64_t block_bytes) : block_bytes_(block_bytes) {
    pe_ = !p2.empty();
    f_[0].open(p1);
    if (pe_) f_[1].open(p2);
}
BlockReader::~BlockReader() {}

// getEndPos @0x4320c0: scan backwards from index `len` for buf[p] == '\n' && buf[p+1] == '@' such that
// buf[p+1+k] == firstline[k] holds for 6 CONSECUTIVE k somewhere in [0, |firstline|).  Bytes at or
// beyond `valid` are stale memory in the reference; they are treated as non-matching here.
int BlockReader::get_end_pos(int len, const char *buf, int valid, const std::string &first) {
    for (int p = len; p > 0; p--) {
        if (buf[p] != '\n' || p + 1 >= valid || buf[p + 1] != '@') continue;
        int run = 0;
        for (size_t k = 0; k < first.size(); k++) {
            int q = p + 1 + (int)k;
            if (q < valid && buf[q] == first[k]) { if (++run > 5) return p; }
            else run = 0;
        }
    }
    return 0;
}

static void first_line(const char *buf, size_t n, std::string &out) {     // getFirstLine @0x431eb0
    size_t i = 0;
    while (i < n && buf[i] != '\n') i++;
    out.assign(buf, i < n ? i + 1 : n);
}

bool BlockReader::next_text(std::vector<char> &t1, std::vector<char> &t2) {
    t1.clear(); t2.clear();
    if (eof_) return false;
    const int nf = pe_ ? 2 : 1;
    uint64_t quota = pe_ ? block_bytes_ / 2 : block_bytes_;            // @0x432dd2: shr $1 for PE
    std::vector<char> *t[2] = {&t1, &t2};
    bool shortread[2] = {false, false};
    for (int k = 0; k < nf; k++) {
        std::vector<char> &b = *t[k];
        size_t L = left_[k].size();
        size_t want = quota > L ? quota - L : 0;
        // leftover first, then the file: the bytes are appended in chunks so that no byte is written twice
        // (a resize to the quota would zero-fill 25 MiB per mate and block on the single reader thread)
        b.reserve(L + want + 1);
        b.assign(left_[k].begin(), left_[k].end());
        left_[k].clear();
        size_t got = 0;
        while (got < want) {
            const size_t chunk = std::min<size_t>(want - got, 4u << 20);
            const size_t at = b.size();
            b.resize(at + chunk);                     // 4 MiB at a time stays in cache between the fill and the read
            size_t k2 = f_[k].read(b.data() + at, chunk);
            b.resize(at + k2);
            got += k2;
            if (k2 < chunk) break;
        }
        if (first_[k].empty()) first_line(b.data() + L, got, first_[k]);
        if (got < want) shortread[k] = true;
    }
    // doReadJob @0x432a80 ends on a short read; doReadPEJob ends only when BOTH files came up short in the same
    // round (@0x432ec2-0x432ede: got[0] >= want[0] || got[1] >= want[1] keeps cutting).  Mates of unequal byte
    // size (trimmed reads, R1 / R2 of different length) exhaust one file first: its leftover keeps feeding blocks
    // whose size the other file's line count decides.
    if (pe_ ? (shortread[0] && shortread[1]) : shortread[0]) {
        eof_ = true;                      // the rest goes out as the last block: cultPEbuf(.., false) keeps
        return !t1.empty();               // min(#lines) lines of both buffers -- equal for valid paired input
    }
    if (!pe_) {                           // cultbuf @0x432530
        int total = (int)t1.size();
        t1.push_back(0);
        int n = get_end_pos(total - (int)first_[0].size(), t1.data(), total, first_[0]);
        int cut = n + 1;
        left_[0].assign(t1.data() + cut, total - cut);
        t1.resize(cut);
        return true;
    }
    // cultPEbuf @0x432180 (cut = true): newline positions of both buffers, back off two lines from the
    // shorter one, find the record start in file 1, cut file 2 after the same number of lines
    std::vector<int> nl[2];
    for (int k = 0; k < 2; k++) {                                    // this scan is on the single reader thread: 16 bytes per compare
        const char *b0 = t[k]->data();
        const size_t n = t[k]->size();
        nl[k].resize(n / 32 + 64);
        size_t cnt = 0, i = 0;
#if defined(__SSE2__)
        const __m128i nlv = _mm_set1_epi8('\n');
        for (; i + 16 <= n; i += 16) {
            unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)(b0 + i)), nlv));
            if (!m) continue;
            if (cnt + 16 > nl[k].size()) nl[k].resize(nl[k].size() * 2);
            while (m) { nl[k][cnt++] = (int)(i + (unsigned)__builtin_ctz(m)); m &= m - 1; }
        }
#endif
        for (; i < n; i++)
            if (b0[i] == '\n') { if (cnt + 1 > nl[k].size()) nl[k].resize(nl[k].size() * 2); nl[k][cnt++] = (int)i; }
        nl[k].resize(cnt);
    }
    int nmin = (int)std::min(nl[0].size(), nl[1].size());
    if (nmin < 3) throw std::runtime_error("block too small to cut a paired-end buffer");
    int e = nmin - 2;
    int total1 = (int)t1.size();
    t1.push_back(0);
    int n = get_end_pos(nl[0][e], t1.data(), total1, first_[0]);
    t1.pop_back();
    while (e > 0 && nl[0][e] != n) e--;
    if (nl[0][e] != n) throw std::runtime_error("paired-end cut failed (no record start found)");
    int cut1 = n + 1, cut2 = nl[1][e] + 1;
    left_[0].assign(t1.data() + cut1, t1.size() - cut1);
    left_[1].assign(t2.data() + cut2, t2.size() - cut2);
    t1.resize(cut1);
    t2.resize(cut2);
    return true;
}

static void parse_one(const std::vector<char> &t, std::vector<const char *> &line, std::vector<int> &len) {
    const char *p = t.data(), *end = p + t.size();
    while (p < end) {
        const char *q = (const char *)memchr(p, '\n', end - p);
        if (!q) q = end;
        line.push_back(p);
        len.push_back((int)(q - p));
        p = q + 1;
    }
}

void parse_block(const std::vector<char> &t1, const std::vector<char> &t2, bool pe, MemBuf &mb) {
    mb.ids.clear(); mb.seq.clear(); mb.qual.clear(); mb.idlen.clear(); mb.seqlen.clear();
    std::vector<const char *> ln[2];
    std::vector<int> ll[2];
    parse_one(t1, ln[0], ll[0]);
    if (pe) parse_one(t2, ln[1], ll[1]);
    size_t nrec = ln[0].size() / 4;
    if (ln[0].size() % 4) throw std::runtime_error("FASTQ block is not a multiple of four lines");
    if (pe && ln[1].size() != ln[0].size()) throw std::runtime_error("paired files disagree in record count");
    mb.ids.reserve(t1.size() / 4 + t2.size() / 4);
    mb.seq.reserve((t1.size() + t2.size()) / 2);
    mb.qual.reserve((t1.size() + t2.size()) / 2);
    mb.comment = nrec && ll[0][2] > 1;
    mb.longflag = false;
    for (size_t r = 0; r < nrec; r++) {
        for (int k = 0; k < (pe ? 2 : 1); k++) {
            const char *id = ln[k][4 * r], *sq = ln[k][4 * r + 1], *pl = ln[k][4 * r + 2], *ql = ln[k][4 * r + 3];
            int il = ll[k][4 * r], sl = ll[k][4 * r + 1], pll = ll[k][4 * r + 2], qll = ll[k][4 * r + 3];
            if (il < 1 || id[0] != '@' || pll < 1 || pl[0] != '+') throw std::runtime_error("malformed FASTQ record");
            // the archive only records WHETHER the '+' lines repeat the ID (the reference's own round trip fails on a mix)
            if (mb.comment ? (pll != il || memcmp(pl + 1, id + 1, (size_t)il - 1) != 0) : pll != 1)
                throw std::runtime_error("'+' lines must be all bare or all a copy of the ID");
            if (sl != qll) throw std::runtime_error("sequence and quality lengths differ");
            if (il - 1 > 255) throw std::runtime_error("read ID longer than 255 bytes (limit of the reference's name model)");
            mb.ids.insert(mb.ids.end(), id + 1, id + il);
            mb.seq.insert(mb.seq.end(), sq, sq + sl);
            mb.qual.insert(mb.qual.end(), ql, ql + qll);
            mb.idlen.push_back((uint16_t)(il - 1));
            mb.seqlen.push_back(sl);
            if (sl > 65535) mb.longflag = true;
        }
    }
}

bool BlockReader::next_raw(Raw &r) {
    if (!next_text(r.t1, r.t2)) return false;
    r.src_off[0] = off_[0]; r.src_off[1] = off_[1];
    off_[0] += r.t1.size(); off_[1] += r.t2.size();
    return true;
}
void BlockReader::parse_raw(const Raw &r, bool pe, MemBuf &mb) {
    mb.src_off[0] = r.src_off[0]; mb.src_off[1] = r.src_off[1];
    mb.src_bytes[0] = r.t1.size(); mb.src_bytes[1] = r.t2.size();
    parse_block(r.t1, r.t2, pe, mb);
}
bool BlockReader::next(MemBuf &mb) {
    Raw r;
    if (!next_raw(r)) return false;
    parse_raw(r, pe_, mb);
    return true;
}

void render_raw(const uint8_t *ids, const uint16_t *idlen, const uint8_t *seq, const uint8_t *qual, const int32_t *seqlen,
                uint32_t nreads, bool comment, bool pe, std::string &o1, std::string &o2) {
    o1.clear(); o2.clear();
    size_t tot[2] = {0, 0};
    for (uint32_t r = 0; r < nreads; r++) tot[(pe && (r & 1)) ? 1 : 0] += (size_t)idlen[r] * (comment ? 2 : 1) + 2 * (size_t)seqlen[r] + 6;
    o1.reserve(tot[0]); o2.reserve(tot[1]);
    size_t io = 0, so = 0;
    for (uint32_t r = 0; r < nreads; r++) {
        std::string &o = (pe && (r & 1)) ? o2 : o1;
        o.push_back('@');
        o.append((const char *)ids + io, idlen[r]);
        o.push_back('\n');
        o.append((const char *)seq + so, seqlen[r]);
        o.append("\n+");
        if (comment) o.append((const char *)ids + io, idlen[r]);
        o.push_back('\n');
        o.append((const char *)qual + so, seqlen[r]);
        o.push_back('\n');
        io += idlen[r];
        so += seqlen[r];
    }
}

void render_block(const MemBuf &mb, bool pe, std::string &o1, std::string &o2) {
    render_raw(mb.ids.data(), mb.idlen.data(), mb.seq.data(), mb.qual.data(), mb.seqlen.data(), mb.nreads(), mb.comment, pe, o1, o2);
}

}  // namespace sqhost
