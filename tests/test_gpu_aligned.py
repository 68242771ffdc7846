"""GPU: the aligned-mode side-stream coders behind the C ABI (sq_side_encode / sq_side_decode) against the bytes the
reference wrote into blocks with mapped reads, single- and paired-end (the block framing and the rebuild of
mapped reads are oracle-only so far)."""
import os
import struct

import numpy as np
import pytest

from tests import common
from oracle import arcwalk, pyoracle as po

pytestmark = pytest.mark.gpu

# tag -> (model size, which count tag holds the number of symbols)
SIDE = {8: (5, 27), 11: (2, 18), 15: (8, 17), 10: (2, 17), 12: (2, 19), 13: (4, 20), 9: (4, 22)}


@pytest.mark.parametrize("name", common.ALIGNED)
def test_side_streams_match_the_reference_bytes(name):
    import fastqueeze_b200 as fq
    arc = arcwalk.parse_arc(open(os.path.join(common.GOLD, name + ".arc"), "rb").read())
    st = arc["blocks"][0]["streams"]
    tags = [t for t in SIDE if t in st and len(st[SIDE[t][1]]) == 4]
    assert len(tags) >= 6
    jobs = [(SIDE[t][0], st[t], struct.unpack("<I", st[SIDE[t][1]])[0]) for t in tags]
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    arrays = codec.decompressSideStreams(jobs)
    for (model, data, n), a in zip(jobs, arrays):
        assert np.array_equal(a, po.small_decode(model, data, n))          # what the oracle reads out of the reference's bytes
    again = codec.compressSideStreams([(j[0], a) for j, a in zip(jobs, arrays)])
    assert again == [j[1] for j in jobs]                                    # and back to the reference's bytes
    codec.close()


def test_side_stream_edge_cases():
    import fastqueeze_b200 as fq
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    rng = np.random.default_rng(3)
    arrays = [(2, np.zeros(0, np.uint8)), (9, rng.integers(0, 9, 70000).astype(np.uint8)), (5, np.full(40000, 4, np.uint8)),
              (4, rng.choice(4, 1, p=[.7, .1, .1, .1]).astype(np.uint8))]
    enc = codec.compressSideStreams(arrays)
    assert enc == [po.small_encode(m, a) for m, a in arrays]                # halving of a saturated model included (40000 x one symbol)
    back = codec.decompressSideStreams([(m, e, a.shape[0]) for (m, a), e in zip(arrays, enc)])
    assert all(np.array_equal(b, a) for b, (m, a) in zip(back, arrays))
    with pytest.raises(fq.SeqArcB200Error):                                 # a symbol the model does not hold
        codec.compressSideStreams([(4, np.array([0, 1, 7], np.uint8))])
    codec.close()


@pytest.mark.parametrize("name", common.ALIGNED)
def test_mapped_reads_are_rebuilt_from_the_reference(name):
    """sq_align_rebuild against the input FASTQ: the side arrays of a reference-written block (read out by the oracle) and the
    reference sequence give back the bases of every mapped read; unmapped reads are left alone."""
    import fastqueeze_b200 as fq
    from tools import make_golden as mg
    m = common.MANIFEST[name]
    fasta, fastq = mg.aligned_case(dict(pe=m["pe"], aligned=m["aligned"]))
    rec = arcwalk.parse_arc(open(os.path.join(common.GOLD, name + ".arc"), "rb").read())["blocks"][0]["raw"]
    o = po.Oracle(md5=1)
    o.set_reference(fasta)
    want = o.decode_block(rec)
    arrays = o.align_arrays()
    off = np.concatenate([[0], np.cumsum(want.seqlen)])
    given = want.seq.copy()
    mapped = np.zeros(given.shape[0], bool)
    for i in np.nonzero(arrays["order"])[0]:
        mapped[off[i]:off[i + 1]] = True
    given[mapped] = ord("?")                                  # what the call has to fill in
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    codec.set_reference(fasta)
    blk = dict(arrays, seqlen=want.seqlen, seq=given, pe=m["pe"])
    got = codec.alignRebuild([blk, blk])                      # two blocks in one call
    for g in got:
        assert np.array_equal(g[~mapped], want.seq[~mapped])
        diff = np.nonzero(g != want.seq)[0]
        # the degenerate-base overlay is not part of this call: a mapped read's N may come back as N (coded as a mismatch) only
        assert all(want.seq[k] == ord("N") for k in diff) and not (g == ord("?")).any()
    bad = dict(blk, pos=arrays["pos"][:100])                  # truncated position bits: reported, not a crash
    with pytest.raises(fq.SeqArcB200Error):
        codec.alignRebuild([bad])
    codec.close()
