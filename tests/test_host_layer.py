"""CPU: the host C++ layer (host/: block cutter, .arc container) against archives the reference wrote."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from tests import common
from oracle import arcwalk

HOSTLIB = os.path.join(common.ROOT, "fastqueeze_b200", "libseqarc_host.so")


@pytest.fixture(scope="module")
def hl():
    import __graft_entry__
    if not os.path.exists(HOSTLIB):
        __graft_entry__.build()
    L = C.CDLL(HOSTLIB)
    L.sqh_cut_blocks.restype = C.c_int
    L.sqh_cut_blocks.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sqh_write_arc.restype = C.c_int
    L.sqh_write_arc.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p,
                                C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
    return L


def _write_inputs(name, d):
    fq = common.case_fastq(name)
    paths = []
    for k, t in enumerate(fq):
        p = os.path.join(d, "in_%d.fq" % (k + 1))
        open(p, "wb").write(t)
        paths.append(p)
    return paths


@pytest.mark.parametrize("name", ["se100_q8_3blocks", "illumina_pe150_3blocks"])
def test_block_cutter_reads_gzip_input(hl, name):
    """The reference cuts a .gz input exactly like the plain file (measured: same records, same block index, only
    trailer param 4 differs); so does the reader here."""
    import gzip
    m = common.MANIFEST[name]
    bs = int((m["config"] or {}).get("BlockSize(M)", 50)) << 20
    with tempfile.TemporaryDirectory() as d:
        paths = _write_inputs(name, d)
        gz = []
        for p in paths:
            with open(p + ".gz", "wb") as f:
                with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:
                    g.write(open(p, "rb").read())
            gz.append(p + ".gz")
        res = []
        for ps in (paths, gz):
            cap = 64
            nr, s1, s2 = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
            n = hl.sqh_cut_blocks(ps[0].encode(), ps[1].encode() if m["pe"] else None, bs, nr.ctypes.data, s1.ctypes.data, s2.ctypes.data, cap)
            res.append((n, list(nr[:n]), list(s1[:n]), list(s2[:n])))
    assert res[0][0] == 3 and res[0] == res[1]


@pytest.mark.parametrize("name", ["se100_q8_3blocks", "illumina_pe150_3blocks", "se100_q8_usemodel", "illumina_pe150", "err_pe1000"])
def test_block_cutter_matches_reference_index(hl, name):
    """Same reads per block and same source byte ranges as the reference's reader (getEndPos / cultbuf / cultPEbuf)."""
    m = common.MANIFEST[name]
    arc = arcwalk.parse_arc(common.case_archive(name))
    idx = arcwalk.block_index(arc, m["pe"])
    bs = int((m["config"] or {}).get("BlockSize(M)", 50)) << 20
    with tempfile.TemporaryDirectory() as d:
        paths = _write_inputs(name, d)
        cap = 64
        nr, s1, s2 = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        n = hl.sqh_cut_blocks(paths[0].encode(), paths[1].encode() if m["pe"] else None, bs,
                              nr.ctypes.data, s1.ctypes.data, s2.ctypes.data, cap)
    assert n == len(arc["blocks"])
    want_reads = [int.from_bytes(b["streams"][1], "little") for b in arc["blocks"]]
    assert list(nr[:n]) == want_reads
    if m["pe"]:
        assert [(int(a), int(b)) for a, b in zip(s1[:n], s2[:n])] == [(r["src1"], r["src2"]) for r in idx]
    else:
        assert [int(a) for a in s1[:n]] == [r["src2"] for r in idx]      # SE record: (len*2, 1, src bytes, 0)


@pytest.mark.parametrize("name", common.CASES)
def test_arc_writer_reproduces_reference_file(hl, name):
    """Re-packing the reference's own block records with ArcWriter gives the reference's file byte for byte."""
    m = common.MANIFEST[name]
    data = common.case_archive(name)
    arc = arcwalk.parse_arc(data)
    idx = arcwalk.block_index(arc, m["pe"])
    recs = b"".join(b["raw"] for b in arc["blocks"])
    lens = np.array([len(b["raw"]) for b in arc["blocks"]], np.uint64)
    if m["pe"]:
        s1 = np.array([r["src1"] for r in idx], np.uint64)
        s2 = np.array([r["src2"] for r in idx], np.uint64)
    else:
        s1 = np.array([r["src2"] for r in idx], np.uint64)
        s2 = None
    params = arc["trailer"]["params"]
    idbin = np.frombuffer(params[15], np.uint8).copy()
    model = arc["trailer"].get("model")
    mb = np.frombuffer(model[4:], np.uint8).copy() if model else None
    ra = np.frombuffer(recs, np.uint8)
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.arc")
        r = hl.sqh_write_arc(out.encode(), int(m["pe"]), int(params[6][0]), int(params[16][0]), int(params[17][0]),
                             params[13], params.get(14, b""), idbin.ctypes.data, mb.ctypes.data if mb is not None else None,
                             0 if mb is None else mb.shape[0], ra.ctypes.data, lens.ctypes.data, s1.ctypes.data,
                             s2.ctypes.data if s2 is not None else None, len(lens), arc["trailer"].get(8), int(params[2] == b"\x00"),
                             bytes(r["long"] for r in idx))
        assert r == 0
        got = open(out, "rb").read()
    assert got == data


def _hl_ids(hl):
    hl.sqh_analyse_ids.restype = C.c_int
    hl.sqh_analyse_ids.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_void_p]
    hl.sqh_decode_ids.restype = C.c_long
    hl.sqh_decode_ids.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    return hl


@pytest.mark.parametrize("name", common.CASES)
def test_id_bin_analysis_matches_reference_trailer(hl, name):
    """IDProcess::analysisIDBinType over the pre-process sample blocks gives the 512-byte array the reference stored."""
    _hl_ids(hl)
    m = common.MANIFEST[name]
    arc = arcwalk.parse_arc(common.case_archive(name))
    want = arc["trailer"]["params"][15]
    bs = int((m["config"] or {}).get("BlockSize(M)", 50)) << 20
    with tempfile.TemporaryDirectory() as d:
        paths = _write_inputs(name, d)
        idbin = np.zeros(512, np.uint8)
        r = hl.sqh_analyse_ids(paths[0].encode(), paths[1].encode() if m["pe"] else None, bs, idbin.ctypes.data)
    assert r == want[0]
    assert idbin.tobytes() == want


@pytest.mark.parametrize("name", ["err_pe1000", "err_pe1000_nomd5", "srr_pe100"])
def test_rule_based_ids_are_regenerated(hl, name):
    """IDProcess::decodeIDS: every ID of the block from the stored first ID + the ID-bin array."""
    _hl_ids(hl)
    kw, blocks = common.case_blocks(name)
    arc = arcwalk.parse_arc(common.case_archive(name))
    idbin = np.frombuffer(arc["trailer"]["params"][15], np.uint8).copy()
    assert kw["id_rule"] == 1
    for (blk, ref), ab in zip(blocks, arc["blocks"]):
        pay = ab["streams"][5][16 if kw["md5"] else 0:]
        pa = np.frombuffer(pay, np.uint8)
        ids = np.zeros(blk.ids.shape[0] + 64, np.uint8)
        idlen = np.zeros(blk.nreads, np.uint16)
        sl = np.ascontiguousarray(blk.seqlen)
        n = hl.sqh_decode_ids(pa.ctypes.data, len(pay), idbin.ctypes.data, sl.ctypes.data, blk.nreads, ids.ctypes.data,
                              ids.shape[0], idlen.ctypes.data)
        assert n == blk.ids.shape[0]
        assert ids[:n].tobytes() == blk.ids.tobytes()
        assert np.array_equal(idlen, blk.idlen)


def test_id_bin_analysis_on_other_id_styles(hl):
    """Styles the reference was observed on plus the edge rules read from the disassembly."""
    _hl_ids(hl)
    import numpy as np

    def run(ids1, ids2=None, L=50):
        with tempfile.TemporaryDirectory() as d:
            ps = []
            for k, ids in enumerate([ids1] + ([ids2] if ids2 else [])):
                p = os.path.join(d, "x_%d.fq" % k)
                with open(p, "wb") as f:
                    for s in ids:
                        f.write(b"@" + s + b"\n" + b"A" * L + b"\n+\n" + b"I" * L + b"\n")
                ps.append(p)
            idbin = np.zeros(512, np.uint8)
            r = hl.sqh_analyse_ids(ps[0].encode(), ps[1].encode() if ids2 else None, 1 << 20, idbin.ctypes.data)
            return r, bytes(idbin[:8])
    assert run([b"read%d" % i for i in range(1, 50)])[0] == 0                    # "read1": one token, not all digits
    r, b = run([b"r.%d" % i for i in range(1, 50)])
    assert r == 1 and b[:4] == bytes([1, 0, 0, 1])                               # constant + counter, single-end
    r, b = run([b"r.%d/1" % i for i in range(1, 50)], [b"r.%d/2" % i for i in range(1, 50)])
    assert r == 1 and b[:5] == bytes([1, 2, 0, 1, 0])                            # petype 2: .../1 vs .../2
    assert run([b"r.%d" % (2 * i) for i in range(1, 50)])[0] == 0                # step 2 is not a rule
    r, b = run([b"r.%d len=50" % i for i in range(1, 50)])
    assert r == 1 and b[:6] == bytes([1, 0, 0, 1, 0, 3])                         # "len" token + read length


def test_mixed_plus_lines_are_refused(hl):
    """The archive only records WHETHER the '+' lines repeat the ID; a mix breaks the reference's own round trip
    (measured) and is an input error here."""
    from tools.make_golden import plus_lines
    t = common.case_fastq("se100_q8")[0]
    lines = plus_lines(t).split(b"\n")
    lines[6] = b"+"                      # second record back to a bare '+'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "mixed.fq")
        open(p, "wb").write(b"\n".join(lines))
        cap = 8
        nr, s1, s2 = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        assert hl.sqh_cut_blocks(p.encode(), None, 50 << 20, nr.ctypes.data, s1.ctypes.data, s2.ctypes.data, cap) == -2
        q = os.path.join(d, "ok.fq")
        open(q, "wb").write(plus_lines(t))
        assert hl.sqh_cut_blocks(q.encode(), None, 50 << 20, nr.ctypes.data, s1.ctypes.data, s2.ctypes.data, cap) == 1
        assert int(nr[0]) == 4000 and int(s1[0]) == len(plus_lines(t))


def test_cli_argument_forms():
    """Both forms of the reference command line: -1 / -2 / -o, or positional [ref.fa] inputs prefix."""
    import subprocess
    cli = os.path.join(common.ROOT, "fastqueeze_b200", "seqarc_b200")

    def parse(args):
        out = subprocess.run([cli, "--parse-only"] + args.split(), capture_output=True, text=True, cwd=tempfile.gettempdir())
        assert out.returncode == 0, out.stderr
        return dict(kv.split("=", 1) for kv in out.stdout.split())

    assert parse("-c -1 a.fq -2 b.fq -o out") == dict(mode="c", ref="", in1="a.fq", in2="b.fq", arc="", out="out", lossy="0", md5="1")
    assert parse("-c -t 16 -f ref.fa a.fq b.fq out -l 1.15") == dict(mode="c", ref="ref.fa", in1="a.fq", in2="b.fq", arc="", out="out", lossy="1", md5="1")
    assert parse("-c a.fq.gz out -n") == dict(mode="c", ref="", in1="a.fq.gz", in2="", arc="", out="out", lossy="0", md5="0")
    assert parse("-c ref.fasta -1 a.fq -o out")["ref"] == "ref.fasta"
    assert parse("-d x.arc -o back") == dict(mode="d", ref="", in1="", in2="", arc="x.arc", out="back", lossy="0", md5="1")
    assert parse("-d -t 8 ref.fa x.arc back") == dict(mode="d", ref="ref.fa", in1="", in2="", arc="x.arc", out="back", lossy="0", md5="1")


@pytest.mark.parametrize("name", common.CASES[:6] + common.ALIGNED)
def test_arc_reader_finds_every_block_record(hl, name):
    """ArcReader (what `seqarc_b200 -d` maps the archive with) against the container walker of the test side: block count,
    record offsets and lengths, the flags of the trailer -- also for archives whose blocks hold mapped reads."""
    hl.sqh_read_arc.restype = C.c_int
    hl.sqh_read_arc.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int]
    m = common.MANIFEST[name]
    data = common.case_archive(name) if name not in common.ALIGNED else open(os.path.join(common.GOLD, name + ".arc"), "rb").read()
    arc = arcwalk.parse_arc(data)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "a.arc")
        open(p, "wb").write(data)
        flags = C.c_int(0)
        off = np.zeros(64, np.uint64)
        ln = np.zeros(64, np.uint32)
        n = hl.sqh_read_arc(p.encode(), C.byref(flags), off.ctypes.data, ln.ctypes.data, 64)
    assert n == len(arc["blocks"])
    for i, b in enumerate(arc["blocks"][:64]):
        assert data[int(off[i]):int(off[i]) + int(ln[i])] == b["raw"]
    assert bool(flags.value & 1) == bool(m["pe"])
    assert bool(flags.value & 8) == ("-n" not in m["opts"])
