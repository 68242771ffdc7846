"""CPU: the oracle (oracle/seqarc_oracle.c) against archives written by the reference binary.

This is what pins the oracle (the reference ships no tests or known-answer vectors):
tests/golden/*.arc were produced by the UNMODIFIED SeqArc-1.6 -t 1 (tools/make_golden.py).
"""
import hashlib

import numpy as np
import pytest

from tests import common
from oracle import arcwalk, pyoracle as po


def test_fixture_inputs_unchanged():
    for name in common.CASES:
        fq = common.case_fastq(name)
        assert [hashlib.md5(t).hexdigest() for t in fq] == common.MANIFEST[name]["fastq_md5"], name


@pytest.mark.parametrize("name", common.CASES)
def test_oracle_encode_matches_reference(name):
    kw, blocks = common.case_blocks(name)
    o = po.Oracle(**kw)
    for i, (blk, ref) in enumerate(blocks):
        rec = o.encode_block(blk)
        if rec != ref:
            a = {t: v for t, v in arcwalk.walk_children(rec, 5, len(rec))}
            raise AssertionError("%s block %d differs: %d vs %d bytes" % (name, i, len(rec), len(ref)))


@pytest.mark.parametrize("name", [c for c in common.CASES if "lossy" not in c])
def test_oracle_decode_reference_records(name):
    kw, blocks = common.case_blocks(name)
    o = po.Oracle(**kw)
    for blk, ref in blocks:
        d = o.decode_block(ref, longflag=blk.longflag)
        assert np.array_equal(d.seqlen, blk.seqlen)
        assert d.seq.tobytes() == blk.seq.tobytes()
        assert d.qual.tobytes() == blk.qual.tobytes()
        if kw["id_rule"]:
            # rule-based IDs (IDProcess::decodeIDS) are host string logic: regenerated and checked against the
            # original IDs in tests/test_host_layer.py; the oracle leaves the ID buffer empty
            assert d.ids.shape[0] == 0
        else:
            assert d.ids.tobytes() == blk.ids.tobytes()
            assert np.array_equal(d.idlen, blk.idlen)


def test_md5_known_answers():
    # RFC 1321 appendix A.5
    assert po.md5(b"").hex() == "d41d8cd98f00b204e9800998ecf8427e"
    assert po.md5(b"abc").hex() == "900150983cd24fb0d6963f7d28e17f72"
    assert po.md5(b"a" * 119).hex() == hashlib.md5(b"a" * 119).hexdigest()
    for n in (55, 56, 63, 64, 65, 1000):
        d = bytes(range(256)) * 4
        assert po.md5(d[:n]).hex() == hashlib.md5(d[:n]).hexdigest()
