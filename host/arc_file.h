// arc_file.h -- the .arc container, format version 1.6, writer and reader.
//   SeqArcFile::SeqArcFile     @0x415c20  ".arc" 01 06 00, TAG 2 with an 8-byte size patched at the end
//   SeqArcFile::writeFileInfo  @0x4171b0  trailer: TAG 3 { TAG 1 params | [TAG 6 model] | TAG 7 block index }
//   SeqArcFile::writeParam     @0x416450  params 1..18
//   SeqArcFile::readFileInfo   @0x419660
//   Encap::setID / setSize / getValue @0x420680-0x4207bd  (EBML-style vints)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sqhost {

struct ArcParams {
    bool pe = false;               // param 3 is 1 for single-end input
    bool usemodel = false;         // param 6
    bool lossy = false;            // param 16
    bool md5 = true;               // param 17
    bool gz = false;               // param 4: the input was gzip-compressed
    bool comment = false;          // param 2 is 0 when the '+' lines repeat the ID (SE index records then carry 0, not 1)
    std::string name1, name2;      // params 13 / 14: base names of the inputs
    uint8_t idbin[512] = {0};      // param 15: ID-bin types (idbin[0] = 1: rule-based ID path)
    std::vector<uint8_t> model;    // trailer TAG 6 payload after its "TAG 1, u16 = 1" header
    bool has_ref = false;          // param 1 is 0 and trailer TAG 8 holds the MD5 of the reference FASTA
    uint8_t ref_md5[16] = {0};     // (SeqArcFile::writeMd5 @0x416ba0, after the model, before the block index)
};

struct BlockIndex {                // one record of trailer TAG 7 (32 B single-end, 40 B paired-end)
    uint32_t len = 0;              // bytes of the block record
    bool longflag = false;         // bit 0 of the index length field: the block holds a read > 65,535 bases
    uint32_t src_bytes[2] = {0, 0};
    uint64_t arc_off = 0;          // offset of the record in the .arc (first one: 16)
    uint64_t src_off[2] = {0, 0};
};

class ArcWriter {
public:
    bool open(const std::string &path);
    // append one block record (already "tag 1 | size(4) | streams"); records must come in block order
    void add_block(const uint8_t *rec, uint64_t len, const uint64_t src_bytes[2], const uint64_t src_off[2], bool longflag = false);
    bool finish(const ArcParams &p);
    uint64_t bytes() const { return off_; }

private:
    FILE *f_ = nullptr;
    uint64_t off_ = 0;
    std::vector<BlockIndex> idx_;
};

struct ArcReader {
    ArcParams p;
    std::vector<BlockIndex> idx;
    const uint8_t *base = nullptr; // the file, mapped read-only (block records are handed to the GPU from here)
    size_t size = 0;
    std::string err;
    bool load(const std::string &path);
    ArcReader() = default;
    ArcReader(const ArcReader &) = delete;
    ArcReader &operator=(const ArcReader &) = delete;
    ~ArcReader();
};

}  // namespace sqhost
