"""GPU: the CUDA path (through the C ABI) against the reference's golden archives and the oracle.

Bit-exact is the bar (integer / byte work).  Every call goes through libseqarc_b200.so; the oracle
is only the checker.
"""
import numpy as np
import pytest

from tests import common
from oracle import arcwalk, pyoracle as po

pytestmark = pytest.mark.gpu

WS = 3 << 30


@pytest.fixture(scope="module")
def fq():
    import fastqueeze_b200
    return fastqueeze_b200


def _streams(rec):
    return {t: bytes(rec[s:e]) for t, s, e in arcwalk.walk_children(rec, 5, len(rec))}


def _explain(codec, o, blk, rec, ref, md5):
    """Which stream differs, and where its coding triples first diverge from the oracle's."""
    msg = ["record differs: %d vs %d bytes" % (len(rec), len(ref))]
    try:
        a, b = _streams(rec), _streams(ref)
    except Exception as e:   # malformed framing
        return "\n".join(msg + ["cannot walk CUDA record: %r" % e, rec[:64].hex(), ref[:64].hex()])
    for t in sorted(set(a) | set(b)):
        x, y = a.get(t), b.get(t)
        if x != y:
            first = next((i for i in range(min(len(x or b""), len(y or b""))) if x[i] != y[i]), -1)
            msg.append("tag %d: cuda %s B, ref %s B, first diff at %d" %
                       (t, None if x is None else len(x), None if y is None else len(y), first))
    for which, nm in ((0, "seq"), (1, "qual"), (2, "name")):
        try:
            got = codec.debug_triples(which)
        except Exception as e:
            msg.append("%s triples unavailable: %r" % (nm, e))
            continue
        if which == 2 and o.p.id_rule:
            continue
        exp = o.trace(which, blk)
        if got.shape != exp.shape:
            msg.append("%s: %d triples vs oracle %d" % (nm, got.shape[0], exp.shape[0]))
        n = min(got.shape[0], exp.shape[0])
        bad = np.nonzero(got[:n] != exp[:n])[0]
        if bad.size:
            i = int(bad[0])
            msg.append("%s: %d/%d triples differ, first at %d: cuda %x oracle %x" %
                       (nm, bad.size, n, i, int(got[i]), int(exp[i])))
        else:
            msg.append("%s: triples agree on the first %d" % (nm, n))
    return "\n".join(msg)


@pytest.mark.parametrize("name", common.CASES)
def test_encode_matches_reference_record(fq, name):
    kw, blocks = common.case_blocks(name)
    codec = fq.EncapFqzComp(common.to_param(kw), workspace_bytes=WS)
    o = po.Oracle(**kw)
    for blk, ref in blocks:     # one block per call so that the debug hooks see this block
        mb = common.to_membuf(blk)
        codec.doFqzEncode([mb])
        assert mb.status == 0
        if mb.out != ref:
            raise AssertionError(name + ": " + _explain(codec, o, blk, mb.out, ref, kw["md5"]))
    # and all blocks in one batch (wave) give the same records
    mbs = [common.to_membuf(b) for b, _ in blocks]
    codec.doFqzEncode(mbs)
    assert [m.out for m in mbs] == [r for _, r in blocks]
    codec.close()


@pytest.mark.parametrize("name", [c for c in common.CASES if "lossy" not in c])
def test_decode_reference_record(fq, name):
    kw, blocks = common.case_blocks(name)
    if kw["id_rule"]:
        pytest.skip("rule-based ID regeneration (IDProcess::decodeIDS) is a 'next' row")
    codec = fq.EncapFqzComp(common.to_param(kw), workspace_bytes=WS)
    out = codec.doDecode([ref for _, ref in blocks])
    for (blk, _), d in zip(blocks, out):
        assert d.status == 0
        assert np.array_equal(d.seqlen, blk.seqlen)
        assert d.qual.tobytes() == blk.qual.tobytes()
        assert d.seq.tobytes() == blk.seq.tobytes()
        assert d.ids.tobytes() == blk.ids.tobytes()
        assert np.array_equal(d.idlen, blk.idlen)
    codec.close()


def test_lossy_decode_matches_oracle(fq):
    kw, blocks = common.case_blocks("illumina_pe150_lossy")
    codec = fq.EncapFqzComp(common.to_param(kw), workspace_bytes=WS)
    o = po.Oracle(**kw)
    for blk, ref in blocks:
        d = codec.doDecode([ref])[0]
        e = o.decode_block(ref)
        assert d.qual.tobytes() == e.qual.tobytes() == po.rblock(blk.qual, kw["lossy_r"]).tobytes()
        assert d.seq.tobytes() == blk.seq.tobytes()
    codec.close()


def test_rblock_matches_oracle(fq):
    codec = fq.EncapFqzComp(workspace_bytes=1 << 28)
    rng = np.random.default_rng(5)
    for n in (1, 2, 77, 50000):
        q = (np.clip(np.rint(rng.normal(30, 8, n)), 0, 60).astype(np.uint8) + 33)
        for r in (1.05, 1.15, 1.5, 3.0):
            assert codec.rblock(q, r).tobytes() == po.rblock(q, r).tobytes()
    codec.close()


def test_md5_mismatch_is_reported(fq):
    kw, blocks = common.case_blocks("illumina_pe150")
    codec = fq.EncapFqzComp(common.to_param(kw), workspace_bytes=WS)
    ref = bytearray(blocks[0][1])
    st = {t: s for t, s, e in arcwalk.walk_children(ref, 5, len(ref))}
    ref[st[7] + 3] ^= 0xff            # corrupt the embedded quality MD5
    d = codec.doDecode([bytes(ref)], raise_on_error=False)[0]
    assert d.status == -10
    codec.close()


def test_train_model_matches_reference_snapshot(fq):
    """encodeForModel on the GPU reproduces the snapshot the reference trained (md5 of its 27 MB blob)."""
    import hashlib
    for name in ("se100_q8_usemodel", "illumina_pe150_usemodel"):
        _train_case(fq, name)


def _train_case(fq, name):
    import hashlib
    m = common.MANIFEST[name]
    t = common.case_fastq(name)
    sample = common.preprocess_sample(t[0], 1 << 20, fastq2=t[1] if m["pe"] else None)
    codec = fq.EncapFqzComp(fq.SeqArcParam(), workspace_bytes=WS)
    blob = codec.encodeForModel([common.to_membuf(b) for b in sample])
    assert len(blob) == m["model_bytes"]
    assert hashlib.md5(blob).hexdigest() == m["model_md5"]
    codec.close()
