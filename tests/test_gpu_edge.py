"""GPU: edge cases of the block coder against the oracle (ragged / empty / degenerate inputs, waves, errors)."""
import numpy as np
import pytest

from tests import common
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fq():
    import fastqueeze_b200
    return fastqueeze_b200


def ragged_block(seed, nreads, maxlen=260, n_rate=0.02, id_style=0, empty=False):
    """Variable read / ID lengths, IUPAC codes, '#' tails.  Zero-length reads and lower-case bases are only
    used for encode parity (empty=True): the reference itself cannot round-trip them (its decoder repeats
    the previous length for an empty read, SURVEY 8a, and restores every base in upper case)."""
    rng = np.random.default_rng(seed)
    recs = []
    alphabet = b"NMRYKSWHBVDn" if empty else b"NMRYKSWHBVD"
    for i in range(nreads):
        L = int(rng.integers(1, maxlen))
        if empty and rng.random() < 0.05:
            L = 0
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        qual = rng.integers(33, 33 + 42, L).astype(np.uint8)
        m = rng.random(L) < n_rate
        seq = np.where(m, rng.choice(np.frombuffer(alphabet, np.uint8), L), seq).astype(np.uint8)
        if L and rng.random() < 0.2:
            qual[L - int(rng.integers(1, L + 1)):] = ord("#")
        if L and rng.random() < 0.03:
            seq[:] = ord("N")                                   # a read of N only
        if id_style == 0:
            ident = b"r%d:%d %s" % (i, int(rng.integers(0, 10 ** int(rng.integers(1, 9)))), b"x" * int(rng.integers(0, 40)))
        else:
            ident = bytes(rng.integers(33, 127, int(rng.integers(0, 200))).astype(np.uint8))
        recs.append((ident, seq.tobytes(), qual.tobytes()))
    return po.block_from_records(recs)


def check(fq, blocks, roundtrip=True, **kw):
    params = dict(usemodel=0, lossy=0, lossy_r=0.0, md5=1, slevel=3, qlevel=2, id_rule=0)
    params.update(kw)
    codec = fq.EncapFqzComp(common.to_param(params), workspace_bytes=3 << 30)
    o = po.Oracle(**params)
    mbs = [common.to_membuf(b) for b in blocks]
    codec.doFqzEncode(mbs)
    for b, m in zip(blocks, mbs):
        assert m.status == 0
        assert m.out == o.encode_block(b)
    back = codec.doDecode([m.out for m in mbs], longflags=[b.longflag for b in blocks]) if roundtrip else []
    for b, d in zip(blocks, back):
        want_q = po.rblock(b.qual, params["lossy_r"]) if params["lossy"] else b.qual
        assert d.status == 0
        assert np.array_equal(d.seqlen, b.seqlen) and np.array_equal(d.idlen, b.idlen)
        assert d.seq.tobytes() == b.seq.tobytes() and d.qual.tobytes() == want_q.tobytes() and d.ids.tobytes() == b.ids.tobytes()
    codec.close()


def test_ragged_reads_with_degenerate_bases(fq):
    check(fq, [ragged_block(1, 3000), ragged_block(2, 1500, id_style=1), ragged_block(3, 1, maxlen=50)])


def low_complexity_block(seed, nreads=1500):
    """Homopolymers and short tandem repeats (periods 1..6), runs of one quality value: the base decoder's context word is
    then re-used inside its look-ahead window (the forwarding log), the quality decoder stays in one context for many
    symbols (the same-context path), and contexts saturate (the halving paths of both models)."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(nreads):
        L = int(rng.integers(20, 400))
        period = int(rng.integers(1, 7))
        unit = rng.choice(np.frombuffer(b"ACGT", np.uint8), period)
        seq = np.tile(unit, L // period + 1)[:L].copy()
        for _ in range(int(rng.integers(0, 4))):                # a few point changes break the period
            seq[int(rng.integers(0, L))] = rng.choice(np.frombuffer(b"ACGTN", np.uint8))
        qual = np.full(L, 33 + int(rng.integers(2, 41)), np.uint8)
        k = int(rng.integers(0, L))
        qual[k:] = 33 + int(rng.integers(2, 41))                # two runs per read
        recs.append((b"lc%d/%d" % (i, period), seq.tobytes(), qual.tobytes()))
    return po.block_from_records(recs)


def test_low_complexity_reads(fq):
    check(fq, [low_complexity_block(21), low_complexity_block(22, 4000)])
    check(fq, [low_complexity_block(23)], slevel=0)             # order-7 contexts repeat even sooner


def wide_quality_block(seed, nreads=2500, lo=0, hi=93, runs=False):
    """Qualities over the coder's whole alphabet ('!'..'~'): the decoder keeps the head of every quality model in the lanes
    it reads per symbol and the rest behind (front / tail split); symbols 48..93 start in the tail, bubble forward across
    the boundary when they are frequent, and models halve with entries on both sides."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(nreads):
        L = int(rng.integers(30, 300))
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        if runs:                                                # few contexts, tens of thousands of symbols each: halving
            vals = rng.integers(lo, hi + 1, 3)
            qual = vals[rng.integers(0, 3, L)]
            qual = np.where(rng.random(L) < 0.02, rng.integers(0, 94, L), qual)
        else:
            qual = rng.integers(lo, hi + 1, L)
        recs.append((b"wq%d" % i, seq.tobytes(), (qual + 33).astype(np.uint8).tobytes()))
    return po.block_from_records(recs)


def test_wide_quality_alphabet(fq):
    check(fq, [wide_quality_block(31), wide_quality_block(32, lo=50, hi=93), wide_quality_block(33, 6000, lo=40, hi=60, runs=True)])
    check(fq, [wide_quality_block(34, 1500), wide_quality_block(35, 4000, lo=44, hi=52, runs=True)], qlevel=1)
    check(fq, [wide_quality_block(36, 1500, lo=20, hi=80)], qlevel=3)


def test_encode_parity_on_inputs_the_reference_cannot_round_trip(fq):
    check(fq, [ragged_block(9, 2000, empty=True)], roundtrip=False)


def test_ragged_lossy_and_nomd5(fq):
    check(fq, [ragged_block(4, 2000, n_rate=0.0)], lossy=1, lossy_r=1.3, md5=0)


def test_other_levels(fq):
    check(fq, [ragged_block(5, 1500)], slevel=0, qlevel=1)
    check(fq, [ragged_block(6, 1500)], slevel=5, qlevel=3)


def test_degenerate_blocks(fq):
    one = po.block_from_records([(b"only", b"ACGTN", b"IIII#")])
    all_n = po.block_from_records([(b"n%d" % i, b"N" * 70, b"#" * 70) for i in range(200)])
    homopolymer = po.block_from_records([(b"h%d" % i, b"A" * 150, b"F" * 150) for i in range(3000)])
    check(fq, [one, all_n, homopolymer])
    empty_reads = po.block_from_records([(b"a", b"", b""), (b"b", b"", b""), (b"", b"ACGT", b"!!!!")])
    check(fq, [empty_reads], roundtrip=False)


def test_long_reads(fq):
    """Reads beyond 65,535 bases switch the block to compressLen_long (four length bytes); mixed with short, empty-tailed
    and N-carrying reads, next to an ordinary block in the same call."""
    rng = np.random.default_rng(77)
    recs = []
    for i, n in enumerate([70001, 12, 65536, 300, 131075]):
        s = np.frombuffer(b"ACGTN", np.uint8)[rng.choice(5, n, p=[.2495, .2495, .2495, .2495, .002])].tobytes()
        q = bytearray((np.clip(np.rint(rng.normal(25, 8, n)), 2, 40).astype(np.uint8) + 33).tobytes())
        if i == 0:
            q[-500:] = b"#" * 500
        recs.append((b"long.%d length=%d" % (i, n), s, bytes(q)))
    long_blk = po.block_from_records(recs)
    assert long_blk.longflag == 1 and ragged_block(5, 200).longflag == 0
    check(fq, [long_blk, ragged_block(5, 200)])


def test_waves_give_identical_records(fq):
    """A workspace too small for the batch forces several waves; records do not depend on the split."""
    blocks = [common.synth_block(kind="illumina_pe150", n_templates=1500, seed=50 + i, n_rate=0.002) for i in range(6)]
    big = fq.EncapFqzComp(workspace_bytes=2 << 30)
    small = fq.EncapFqzComp(workspace_bytes=60 << 20)
    a = [common.to_membuf(b) for b in blocks]
    b = [common.to_membuf(b) for b in blocks]
    d1, d2 = fq.DebugInfo(), fq.DebugInfo()
    big.doFqzEncode(a, d1)
    small.doFqzEncode(b, d2)
    assert d2.wave_blocks < len(blocks) <= d1.wave_blocks
    assert [m.out for m in a] == [m.out for m in b]
    o = po.Oracle()
    assert a[3].out == o.encode_block(blocks[3])
    back = big.doDecode([m.out for m in b])
    assert all(x.status == 0 for x in back) and back[5].qual.tobytes() == blocks[5].qual.tobytes()
    big.close(); small.close()


def test_input_the_reference_cannot_represent_is_rejected(fq):
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    good = ragged_block(7, 300)
    long_id = po.block_from_records([(b"x" * 300, b"ACGT", b"IIII")])
    bad_q = po.block_from_records([(b"q", b"ACGT", b"II\x1fI")])
    del_q = po.block_from_records([(b"q", b"ACGT", b"II\x7fI")])                 # 0x7f would become symbol 94, the read terminator
    hi_id = po.block_from_records([(b"id\xc3\xa9 1", b"ACGT", b"IIII")])           # the name model codes ch & 0x7f
    tilde = po.block_from_records([(b"q", b"ACGT", b"I~~I")])                     # '~' (symbol 93) is the largest quality the coder holds
    mbs = [common.to_membuf(b) for b in (good, long_id, bad_q, del_q, hi_id, tilde)]
    codec.doFqzEncode(mbs, raise_on_error=False)
    assert mbs[0].status == 0 and mbs[0].out == po.Oracle().encode_block(good)      # a bad block does not hurt its neighbours
    assert [m.status for m in mbs[1:5]] == [-5, -5, -5, -5]                          # SQ_ERR_INPUT
    assert mbs[5].status == 0 and mbs[5].out == po.Oracle().encode_block(tilde)
    codec.close()


def test_corrupt_record_is_reported_not_crashed(fq):
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    blk = ragged_block(8, 500)
    mb = common.to_membuf(blk)
    codec.doFqzEncode([mb])
    rec = bytearray(mb.out)
    rec[len(rec) // 2] ^= 0x55                       # payload corruption -> an MD5 mismatch
    truncated = bytes(mb.out[:40])
    out = codec.doDecode([bytes(rec), truncated, b"\x00\x01\x02"], raise_on_error=False)
    assert out[0].status in (-8, -9, -10, -6, -4)
    assert out[1].status == -6 and out[2].status == -6
    codec.close()


def test_block_with_mapped_reads_is_refused_not_misdecoded(fq):
    """Aligned-mode side streams are not coded by this path yet: a reference-written block that holds
    mapped reads must come back SQ_ERR_FORMAT, next to an ordinary block that decodes."""
    import os
    from oracle import arcwalk
    name = "aligned_se100"
    arc = arcwalk.parse_arc(open(os.path.join(common.GOLD, name + ".arc"), "rb").read())
    aligned = arc["blocks"][0]["raw"]
    codec = fq.EncapFqzComp(workspace_bytes=1 << 30)
    good = ragged_block(12, 300)
    mb = common.to_membuf(good)
    codec.doFqzEncode([mb])
    out = codec.doDecode([mb.out, aligned], raise_on_error=False)
    assert out[0].status == 0 and out[0].seq.tobytes() == good.seq.tobytes()
    assert out[1].status == -6
    codec.close()
