// fastq_block.h -- host side of the path: FASTQ reader / block cutter and block assembly.
//
// Restates, for plain (optionally gzip-free) 4-line FASTQ, what the reference's reader thread and
// block parser do, so that every block holds exactly the reads the reference would put in it:
//   SeqArcRead::doReadJob    @0x432a80   leftover + read(BlockSize - leftover), cut, hand over
//   SeqArcRead::doReadPEJob  @0x432d10   the same per mate file with BlockSize / 2
//   SeqArcRead::getFirstLine @0x431eb0   first line of the file INCLUDING its '\n'
//   SeqArcRead::getEndPos    @0x4320c0   last "\n@" whose following bytes agree with the first line
//                                        in 6 consecutive positions
//   SeqArcRead::cultbuf      @0x432530   SE cut;  SeqArcRead::cultPEbuf @0x432180  PE cut by line index
//   AlignEncodeSEJob::getBlockRead @0x411b60 / AlignEncodePEJob::getBlockReadPE @0x412920
//                                        FASTQ text -> id / seq / qual buffers (PE interleaved R1,R2,...)
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace sqhost {

// One block as SeqArcMemBuf holds it at the doFqzEncode seam.
struct MemBuf {
    std::vector<uint8_t> ids, seq, qual;
    std::vector<uint16_t> idlen;
    std::vector<int32_t> seqlen;
    uint64_t src_bytes[2] = {0, 0};     // FASTQ bytes this block took from file 1 / file 2
    uint64_t src_off[2] = {0, 0};       // their offsets in the source files
    bool longflag = false;              // membuf+0x2: a read longer than 65,535 bases (Len stream with four length bytes)
    bool comment = false;               // the '+' lines repeat the ID (trailer param 2 = 0); decided by the block's first record
    uint32_t nreads() const { return (uint32_t)seqlen.size(); }
};

// FASTQ input, plain or gzip (the reference reads both; zlib's gz layer passes plain files through).
class InFile {
public:
    InFile() = default;
    InFile(const InFile &) = delete;
    InFile &operator=(const InFile &) = delete;
    ~InFile() { close(); }
    bool open(const std::string &path);
    void close();
    bool ok() const { return f_ != nullptr || fd_ >= 0; }
    size_t read(void *dst, size_t n);                 // like fread(dst, 1, n)
    bool getline(std::vector<char> &append_to);       // appends one line including its '\n'; false at end of input
    static bool is_gz_name(const std::string &path) { return path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0; }

private:
    void *f_ = nullptr;                               // gzFile (gzip input)
    int fd_ = -1;                                     // plain input: read(2) straight from the page cache, no stdio copy
    std::vector<char> gbuf_;                          // read-ahead of getline()
    size_t gpos_ = 0;
};

class BlockReader {
public:
    // path2 empty = single-end.  block_bytes = SeqArcParam BlockSize (50 MiB by default).
    BlockReader(const std::string &path1, const std::string &path2, uint64_t block_bytes);
    ~BlockReader();
    bool ok() const { return f_[0].ok() && (!pe_ || f_[1].ok()); }
    bool pe() const { return pe_; }
    // next block of FASTQ text (cut exactly like the reference); false at end of input
    bool next_text(std::vector<char> &t1, std::vector<char> &t2);
    // next block parsed into a MemBuf; false at end of input; throws std::runtime_error on malformed FASTQ
    bool next(MemBuf &mb);
    // the same in two steps, so that callers can parse several blocks in parallel: next_raw cuts the text
    // (sequential), parse_raw fills the MemBuf (independent per block)
    struct Raw { std::vector<char> t1, t2; uint64_t src_off[2] = {0, 0}; };
    bool next_raw(Raw &r);
    static void parse_raw(const Raw &r, bool pe, MemBuf &mb);
    static int get_end_pos(int len, const char *buf, int valid, const std::string &firstline);

private:
    InFile f_[2];
    bool pe_ = false, eof_ = false;
    uint64_t block_bytes_;
    std::string first_[2], left_[2];
    uint64_t off_[2] = {0, 0};
};

// FASTQ text -> MemBuf (SE: t2 empty).  '+' lines are either all bare or all "+<ID>" (MemBuf::comment).
void parse_block(const std::vector<char> &t1, const std::vector<char> &t2, bool pe, MemBuf &mb);
// inverse (recoverDataNoComment): MemBuf -> FASTQ text, PE de-interleaved
void render_block(const MemBuf &mb, bool pe, std::string &o1, std::string &o2);
void render_raw(const uint8_t *ids, const uint16_t *idlen, const uint8_t *seq, const uint8_t *qual, const int32_t *seqlen,
                uint32_t nreads, bool comment, bool pe, std::string &o1, std::string &o2);

}  // namespace sqhost
