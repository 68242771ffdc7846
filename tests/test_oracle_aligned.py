"""Oracle vs reference-written blocks that hold MAPPED reads, single- and paired-end.
  * the coders of the aligned-mode side streams: each stream is the bare range-coder output of a byte array through a fresh
    SIMPLE_MODEL<N> (oracle/seqarc_oracle.c: sqo_small_encode / sqo_small_decode), the counts travel in their own tags;
    decoding a stream and coding it again must give back the reference's bytes;
  * the decode side as a whole (doAlignDecode @0x42c8d0): with the reference sequence the oracle rebuilds every mapped read
    (getRealPos @0x42de30, AlignInfoToSeq @0x42e020) and gives back the input FASTQ, degenerate bases included;
  * the encode side given the aligner's arrays (doAlignEncode @0x42d4c0): the whole record, byte for byte.
(The CUDA path does not code these streams yet: DESIGN.md section 7.)"""
import hashlib
import os
import struct

import numpy as np
import pytest

from tests import common
from oracle import arcwalk, pyoracle as po

# tag -> (name, model size, which count)
SIDE = {8: ("Order", 5, "order"), 11: ("Pos", 2, "pos"), 15: ("Mis", 8, "align"), 10: ("Rev", 2, "align"),
        12: ("CigaL", 2, "cigal"), 13: ("CigaV", 4, "cigav")}


def _inputs(name):
    """(reference FASTA, (FASTQ, ...)) of an aligned golden, rebuilt by the seeded generator that made it."""
    from tools import make_golden as mg
    m = common.MANIFEST[name]
    return mg.aligned_case(dict(pe=m["pe"], aligned=m["aligned"]))


@pytest.mark.parametrize("name", common.ALIGNED)
def test_aligned_side_streams_decode_and_recode_to_the_reference_bytes(name):
    m = common.MANIFEST[name]
    data = open(os.path.join(common.GOLD, name + ".arc"), "rb").read()
    assert hashlib.md5(data).hexdigest() == m["arc_md5"]
    arc = arcwalk.parse_arc(data)
    assert arc["trailer"][8].hex() == m["ref_md5"]                      # MD5 of the reference FASTA (SeqArcFile::writeMd5 @0x416ba0)
    for blk in arc["blocks"]:
        st = blk["streams"]
        u32 = lambda t: struct.unpack("<I", st[t])[0]                   # noqa: E731
        # compressReadCount / OrderCount / AlignCount / PosCount / CigalCount / CigaVCount @0x422a60-0x422b8f
        cnt = {"reads": u32(1), "order": u32(27), "align": u32(17), "pos": u32(18), "cigal": u32(19), "cigav": u32(20)}
        assert cnt["order"] == cnt["reads"] and 0 < cnt["align"] <= cnt["reads"]
        sym = {}
        for tag, (label, n_model, which) in SIDE.items():
            sym[label] = po.small_decode(n_model, st[tag], cnt[which])
            assert po.small_encode(n_model, sym[label]) == st[tag], label
        # what the arrays say, as far as the coders' parity needs it: one Order symbol per read (0 = not mapped), one
        # mismatch count and one strand flag per mapped read, one CigaV symbol (which of the three other bases) per mismatch
        assert int((sym["Order"] == 0).sum()) == cnt["reads"] - cnt["align"]
        assert int(sym["Mis"].astype(np.int64).sum()) == cnt["cigav"]
        assert sym["CigaV"].max() <= 3                                  # 3 = the read has N where the reference has a base
        if not m["pe"]:
            assert cnt["pos"] % cnt["align"] == 0                       # single-end: the same number of position bits per mapped read
        else:                                                           # compressPERelation @0x422be0: one symbol per pair mapped with both mates
            n_rel = u32(22)
            rel = po.small_decode(4, st[9], n_rel)
            assert po.small_encode(4, rel) == st[9] and 0 < n_rel <= cnt["align"] // 2
        # the entropy-coded streams of the same block are still there: IDs, qualities, and the bases of the unmapped reads
        assert all(t in st for t in (4, 5, 7, 6))


@pytest.mark.parametrize("name", common.ALIGNED)
def test_oracle_decodes_the_aligned_block_back_to_the_input(name):
    m = common.MANIFEST[name]
    fasta, fastq = _inputs(name)
    assert [hashlib.md5(t).hexdigest() for t in fastq] == m["fastq_md5"] and hashlib.md5(fasta).hexdigest() == m["ref_md5"]
    arc = arcwalk.parse_arc(open(os.path.join(common.GOLD, name + ".arc"), "rb").read())
    want = po.block_from_fastq(*fastq)
    o = po.Oracle(md5=1)
    with pytest.raises(RuntimeError):                       # mapped reads cannot be rebuilt without the reference
        o.decode_block(arc["blocks"][0]["raw"])
    o.set_reference(fasta)
    got = o.decode_block(arc["blocks"][0]["raw"])
    assert np.array_equal(got.seqlen, want.seqlen) and np.array_equal(got.idlen, want.idlen)
    assert got.ids.tobytes() == want.ids.tobytes() and got.qual.tobytes() == want.qual.tobytes()
    assert got.seq.tobytes() == want.seq.tobytes()
    st = arc["blocks"][0]["streams"]
    assert st[6][:16] == po.md5(want.seq.tobytes())          # the block's base MD5 covers mapped and unmapped reads alike
    mapped = struct.unpack("<I", st[17])[0]
    assert 0 < mapped < want.seqlen.size                     # both kinds of read are in the block


@pytest.mark.parametrize("name", common.ALIGNED)
def test_oracle_encodes_the_aligned_block_from_the_aligners_arrays(name):
    """Encode side (doAlignEncode @0x42d4c0): the input block plus what the aligner found -- taken from the decoded golden,
    the aligner itself is outside the path -- must give the reference's whole record: side streams, counts, Dege streams over
    all reads, and a base stream that codes the unmapped reads only."""
    m = common.MANIFEST[name]
    fasta, fastq = _inputs(name)
    rec = arcwalk.parse_arc(open(os.path.join(common.GOLD, name + ".arc"), "rb").read())["blocks"][0]["raw"]
    d = po.Oracle(md5=1)
    d.set_reference(fasta)
    d.decode_block(rec)
    arrays = d.align_arrays()
    e = po.Oracle(md5=1, ref_mode=2 if m["pe"] else 1)
    e.set_align(arrays)
    blk = po.block_from_fastq(*fastq)
    assert e.encode_block(blk) == rec
    e.set_align(None)                                        # without them the same call writes the nothing-mapped framing
    plain = e.encode_block(blk)
    assert plain != rec and po.Oracle(md5=1).decode_block(plain).seq.tobytes() == blk.seq.tobytes()


def test_small_coder_round_trip_on_random_arrays():
    rng = np.random.default_rng(5)
    for n_model, n in ((2, 50000), (4, 3000), (5, 20000), (8, 1), (9, 70000), (2, 0)):
        p = rng.dirichlet(np.ones(n_model) * 0.3)
        a = rng.choice(n_model, n, p=p).astype(np.uint8)
        enc = po.small_encode(n_model, a)
        assert len(enc) >= 8
        assert np.array_equal(po.small_decode(n_model, enc, n), a)
