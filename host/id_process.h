// id_process.h -- the reference's rule-based read-ID path (IDProcess.cpp), host-side string logic.
//   IDProcess::analysisIDBinType @0x4310a0  decide per sample block whether every ID is predictable from the
//                                           block's first ID: fields (split at punctuation / space) are
//                                           constant (0), +1 per read or pair (1) or "the read length after a
//                                           'length' token" (3); paired IDs relate as identical (1), .../1 vs
//                                           .../2 (2) or equal up to "length" (3)          -> idbin[512]
//   IDProcess::analysisPEType    @0x430f50
//   IDProcess::encodeIDS         @0x430040  block payload = u16le length + first ID          (done in k_pack)
//   IDProcess::decodeIDS         @0x430610 / assembleNameS @0x4300a0 / strSplitWithDeli @0x430360
//   SeqArcPreProcess::doPreProcess @0x4153b0  runs the analysis over the (<= 5) head sample blocks and stops
//                                           at the first block that fails
// idbin layout (SeqArcParam+0x18a4, trailer param 15): [0] rule path on, [1] PE relation, [2 + j] type of field j.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fastq_block.h"

namespace sqhost {

// one call of analysisIDBinType on one sample block; returns idbin[0]
bool analyse_id_bins(const MemBuf &mb, bool pe, uint8_t idbin[512]);

// decodeIDS: regenerate every ID of a block from its tag-5 payload (u16le length + first ID).
// seqlen / nreads come from the decoded Len stream.  Appends to ids / idlen.
bool decode_ids(const uint8_t *payload, size_t paylen, const uint8_t idbin[512], const int32_t *seqlen, uint32_t nreads,
                std::vector<uint8_t> &ids, std::vector<uint16_t> &idlen);

void md5_digest(const uint8_t *data, size_t n, uint8_t out[16]);     // RFC 1321 (Md5.cpp @0x404de0)

}  // namespace sqhost
