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
    uint32_t nreads() const { return (uint32_t)seqlen.size(); }
};

class BlockReader {
public:
    // path2 empty = single-end.  block_bytes = SeqArcParam BlockSize (50 MiB by default).
    BlockReader(const std::string &path1, const std::string &path2, uint64_t block_bytes);
    ~BlockReader();
    bool ok() const { return f_[0] != nullptr && (!pe_ || f_[1] != nullptr); }
    bool pe() const { return pe_; }
    // next block of FASTQ text (cut exactly like the reference); false at end of input
    bool next_text(std::vector<char> &t1, std::vector<char> &t2);
    // next block parsed into a MemBuf; false at end of input; throws std::runtime_error on malformed FASTQ
    bool next(MemBuf &mb);
    static int get_end_pos(int len, const char *buf, int valid, const std::string &firstline);

private:
    FILE *f_[2] = {nullptr, nullptr};
    bool pe_ = false, eof_ = false;
    uint64_t block_bytes_;
    std::string first_[2], left_[2];
    uint64_t off_[2] = {0, 0};
};

// FASTQ text -> MemBuf (SE: t2 empty).  Only plain "+" separator lines are supported.
void parse_block(const std::vector<char> &t1, const std::vector<char> &t2, bool pe, MemBuf &mb);
// inverse (recoverDataNoComment): MemBuf -> FASTQ text, PE de-interleaved
void render_block(const MemBuf &mb, bool pe, std::string &o1, std::string &o2);

}  // namespace sqhost
