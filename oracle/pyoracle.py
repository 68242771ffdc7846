"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/liboracle.so) plus small
FASTQ -> block helpers.  Only tests/, tools/, __graft_entry__.smoke() and bench.py's CPU legs
may import this; the product (fastqueeze_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "SeqArc-1.6")
NLGUARD = os.path.join(HERE, "_ref", "nlguard.so")
SIZEFAKE = os.path.join(HERE, "_ref", "sizefake.so")


class Params(C.Structure):
    _fields_ = [("usemodel", C.c_int32), ("lossy", C.c_int32), ("lossy_r", C.c_double),
                ("md5", C.c_int32), ("slevel", C.c_int32), ("qlevel", C.c_int32),
                ("id_rule", C.c_int32), ("idbin", C.c_void_p), ("model", C.c_void_p),
                ("modellen", C.c_uint64), ("ref_mode", C.c_int32)]


class CBlock(C.Structure):
    _fields_ = [("nreads", C.c_uint32), ("longflag", C.c_int32), ("ids", C.c_void_p),
                ("seq", C.c_void_p), ("qual", C.c_void_p), ("idlen", C.c_void_p),
                ("seqlen", C.c_void_p), ("idbytes", C.c_uint64), ("seqbytes", C.c_uint64)]


@dataclass
class Block:
    """One block as SeqArcMemBuf holds it at the doFqzEncode seam."""
    ids: np.ndarray      # uint8, concatenated, no '@'
    idlen: np.ndarray    # uint16
    seq: np.ndarray      # uint8
    qual: np.ndarray     # uint8
    seqlen: np.ndarray   # int32

    @property
    def nreads(self):
        return int(self.seqlen.shape[0])

    @property
    def longflag(self):
        """membuf+0x2: the block holds a read longer than 65,535 bases (measured: 65,536 sets it, 65,535 does not)."""
        return int(self.seqlen.shape[0] > 0 and int(self.seqlen.max()) > 65535)

    def cblock(self, longflag=None):
        b = CBlock()
        b.nreads = self.nreads
        b.longflag = self.longflag if longflag is None else int(longflag)
        self._keep = [np.ascontiguousarray(a) for a in (self.ids, self.seq, self.qual, self.idlen, self.seqlen)]
        b.ids, b.seq, b.qual, b.idlen, b.seqlen = [a.ctypes.data for a in self._keep]
        b.idbytes = int(self.ids.shape[0])
        b.seqbytes = int(self.seq.shape[0])
        return b


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in os.listdir(HERE)
            if f.endswith((".c", ".inc", ".h"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    elif os.path.isdir("/root/reference") and not os.path.exists(REF_BIN):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.sqo_create.restype = C.c_void_p
        L.sqo_create.argtypes = [C.POINTER(Params)]
        L.sqo_destroy.argtypes = [C.c_void_p]
        L.sqo_encode_block.restype = C.c_int64
        L.sqo_encode_block.argtypes = [C.c_void_p, C.POINTER(CBlock), C.c_void_p, C.c_uint64]
        L.sqo_decode_block.restype = C.c_int64
        L.sqo_decode_block.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(CBlock),
                                       C.c_uint64, C.c_uint64, C.c_uint32]
        for n in ("sqo_stream_seq", "sqo_stream_qual", "sqo_stream_name", "sqo_stream_len"):
            f = getattr(L, n)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, C.POINTER(CBlock), C.c_void_p]
        L.sqo_trace_stream.restype = C.c_int64
        L.sqo_trace_stream.argtypes = [C.c_void_p, C.POINTER(CBlock), C.c_int, C.c_void_p, C.c_void_p]
        for n in ("sqo_small_encode", "sqo_small_decode"):
            f = getattr(L, n)
            f.restype = C.c_int64
            f.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
        L.sqo_set_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                    C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.sqo_set_align.restype = None
        L.sqo_set_align_pe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.sqo_set_align_pe.restype = None
        L.sqo_get_align.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.sqo_get_align.restype = C.c_uint32
        L.sqo_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.sqo_set_reference.restype = None
        L.sqo_md5.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.sqo_rblock.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
        L.sqo_model_size.restype = C.c_uint64
        L.sqo_model_size.argtypes = [C.POINTER(Params)]
        L.sqo_train_block.argtypes = [C.c_void_p, C.POINTER(CBlock), C.c_int]
        L.sqo_save_model.restype = C.c_uint64
        L.sqo_save_model.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, usemodel=0, lossy=0, lossy_r=0.0, md5=1, slevel=3, qlevel=2, id_rule=0,
                 idbin=None, model=None, ref_mode=0):
        self.p = Params()
        self.p.ref_mode = ref_mode
        self.p.usemodel, self.p.lossy, self.p.lossy_r = usemodel, lossy, lossy_r
        self.p.md5, self.p.slevel, self.p.qlevel, self.p.id_rule = md5, slevel, qlevel, id_rule
        self._idbin = np.frombuffer(idbin, dtype=np.uint8).copy() if idbin is not None else None
        self._model = np.frombuffer(model, dtype=np.uint8).copy() if model is not None else None
        self.p.idbin = self._idbin.ctypes.data if self._idbin is not None else None
        self.p.model = self._model.ctypes.data if self._model is not None else None
        self.p.modellen = 0 if self._model is None else self._model.shape[0]
        self.h = lib().sqo_create(C.byref(self.p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().sqo_destroy(self.h)
            self.h = None

    def encode_block(self, blk: Block) -> bytes:
        cap = int(blk.ids.shape[0] + 2 * blk.seq.shape[0] + 8 * blk.nreads + (1 << 16))
        out = np.empty(cap, dtype=np.uint8)
        q = blk.qual.copy() if self.p.lossy else blk.qual   # rblock smooths in place
        b2 = Block(blk.ids, blk.idlen, blk.seq, q, blk.seqlen)
        cb = b2.cblock()
        n = lib().sqo_encode_block(self.h, C.byref(cb), out.ctypes.data, cap)
        assert 0 < n <= cap
        return out[:n].tobytes()

    def stream(self, which, blk: Block) -> bytes:
        cap = int(blk.ids.shape[0] + 2 * blk.seq.shape[0] + 8 * blk.nreads + (1 << 16))
        out = np.empty(cap, dtype=np.uint8)
        q = blk.qual.copy()
        cb = Block(blk.ids, blk.idlen, blk.seq, q, blk.seqlen).cblock()
        n = getattr(lib(), "sqo_stream_" + which)(self.h, C.byref(cb), out.ctypes.data)
        return out[:n].tobytes()

    def trace(self, which, blk: Block) -> np.ndarray:
        """Coding triples (cum | freq<<16 | tot<<32) of stream 0 seq / 1 qual / 2 name, in stream order."""
        nmax = int(blk.ids.shape[0] + blk.seq.shape[0] + 4 * blk.nreads + 16)
        tr = np.empty(nmax, dtype=np.uint64)
        scratch = np.empty(2 * nmax + 64, dtype=np.uint8)
        q = blk.qual.copy()
        cb = Block(blk.ids, blk.idlen, blk.seq, q, blk.seqlen).cblock()
        n = lib().sqo_trace_stream(self.h, C.byref(cb), which, tr.ctypes.data, scratch.ctypes.data)
        return tr[:n].copy()

    def train(self, blocks) -> bytes:
        """encodeForModel @0x4228e0 over the sample blocks (no reset in between) + SaveModelToMem @0x420c00."""
        for j, blk in enumerate(blocks):
            cb = blk.cblock()
            lib().sqo_train_block(self.h, C.byref(cb), 1 if j == 0 else 0)
        n = int(lib().sqo_model_size(C.byref(self.p)))
        buf = np.zeros(n, np.uint8)
        got = lib().sqo_save_model(self.h, buf.ctypes.data)
        assert got == n
        return buf.tobytes()

    def set_reference(self, fasta: bytes):
        """Aligned mode (decode, single-end): the reference FASTA the archive was written against; header lines dropped, the
        sequences concatenated."""
        seq = b"".join(l for l in fasta.split(b"\n") if l and not l.startswith(b">"))
        self._ref = np.frombuffer(seq, np.uint8).copy()
        lib().sqo_set_reference(self.h, self._ref.ctypes.data, self._ref.shape[0])

    def align_arrays(self):
        """Side arrays of the aligned block decoded last: dict(order, pos, mis, rev, cigal, cigav) of uint8 arrays."""
        out = {}
        for k, name in enumerate(("order", "pos", "mis", "rev", "cigal", "cigav", "perel")):
            ptr = C.c_void_p()
            n = lib().sqo_get_align(self.h, k, C.byref(ptr))
            out[name] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (n,)).copy() if n else np.zeros(0, np.uint8)
        out["insbits"] = int(lib().sqo_get_align(self.h, 7, C.byref(C.c_void_p())))      # paired-end: PEInsrNum
        return out

    def set_align(self, a):
        """The aligner's arrays for the next encode_block call (ref_mode = 1, single-end); None clears them."""
        self._al = {k: np.ascontiguousarray(v, np.uint8) for k, v in (a or {}).items() if k != "insbits"}
        pe = self._al.get("perel")
        lib().sqo_set_align_pe(self.h, pe.ctypes.data if pe is not None and pe.shape[0] else None,
                               int(pe.shape[0]) if pe is not None else 0, int((a or {}).get("insbits", 0)))
        g = lambda k: self._al[k].ctypes.data if a and self._al[k].shape[0] else None      # noqa: E731
        n = lambda k: int(self._al[k].shape[0]) if a else 0                               # noqa: E731
        lib().sqo_set_align(self.h, g("order"), g("pos"), n("pos"), g("mis"), g("rev"), n("mis"), g("cigal"), n("cigal"),
                            g("cigav"), n("cigav"))

    def decode_block(self, rec: bytes, idcap=None, seqcap=None, readcap=None, longflag=0) -> Block:
        rec_a = np.frombuffer(rec, dtype=np.uint8)
        seqcap = seqcap or (64 << 20)
        idcap = idcap or (64 << 20)
        readcap = readcap or (1 << 22)
        blk = Block(np.zeros(idcap, np.uint8), np.zeros(readcap, np.uint16), np.zeros(seqcap, np.uint8),
                    np.zeros(seqcap, np.uint8), np.zeros(readcap, np.int32))
        cb = blk.cblock(longflag=longflag)        # the flag travels in the archive's block index, not in the record
        n = lib().sqo_decode_block(self.h, rec_a.ctypes.data, len(rec), C.byref(cb), idcap, seqcap, readcap)
        if n < 0:
            raise RuntimeError("oracle decode failed: %d" % n)
        nr = cb.nreads
        return Block(blk.ids[:cb.idbytes].copy(), blk.idlen[:nr].copy(), blk.seq[:cb.seqbytes].copy(),
                     blk.qual[:cb.seqbytes].copy(), blk.seqlen[:nr].copy())


def md5(data: bytes) -> bytes:
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(16, np.uint8)
    lib().sqo_md5(a.ctypes.data if len(data) else None, len(data), out.ctypes.data)
    return out.tobytes()


def small_encode(N: int, sym: np.ndarray) -> bytes:
    """Range-coder bytes of `sym` through a fresh SIMPLE_MODEL<N> (aligned-mode side streams, compressOrder @0x424b70 ...)."""
    sym = np.ascontiguousarray(sym, np.uint8)
    out = np.empty(2 * sym.shape[0] + 64, np.uint8)
    n = lib().sqo_small_encode(N, sym.ctypes.data if sym.shape[0] else None, sym.shape[0], out.ctypes.data)
    assert n >= 0
    return out[:n].tobytes()


def small_decode(N: int, data: bytes, n: int) -> np.ndarray:
    buf = np.frombuffer(data + bytes(16), np.uint8)
    sym = np.empty(n, np.uint8)
    r = lib().sqo_small_decode(N, buf.ctypes.data, n, sym.ctypes.data if n else None)
    assert r >= 0
    return sym


def rblock(q: np.ndarray, r: float) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.uint8).copy()
    lib().sqo_rblock(q.ctypes.data, q.shape[0], r)
    return q


# ------------------------------------------------------------------ FASTQ helpers (test side)
def parse_fastq(data: bytes):
    """4-line FASTQ -> list of (id_without_at, seq, qual) byte strings."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    assert len(lines) % 4 == 0, "not a 4-line FASTQ"
    recs = []
    for i in range(0, len(lines), 4):
        assert lines[i][:1] == b"@" and lines[i + 2][:1] == b"+"
        recs.append((lines[i][1:], lines[i + 1], lines[i + 3]))
    return recs


def block_from_records(recs) -> Block:
    ids = b"".join(r[0] for r in recs)
    seq = b"".join(r[1] for r in recs)
    qual = b"".join(r[2] for r in recs)
    return Block(np.frombuffer(ids, np.uint8).copy(), np.array([len(r[0]) for r in recs], np.uint16),
                 np.frombuffer(seq, np.uint8).copy(), np.frombuffer(qual, np.uint8).copy(),
                 np.array([len(r[1]) for r in recs], np.int32))


def block_from_fastq(data1: bytes, data2: bytes = None) -> Block:
    """Whole file(s) as ONE block; PE is interleaved R1,R2,R1,R2 (getBlockReadPE @0x412920)."""
    r1 = parse_fastq(data1)
    if data2 is None:
        return block_from_records(r1)
    r2 = parse_fastq(data2)
    assert len(r1) == len(r2)
    recs = [None] * (2 * len(r1))
    recs[0::2] = r1
    recs[1::2] = r2
    return block_from_records(recs)


def block_to_fastq(blk: Block, pe=False):
    """Inverse of block_from_fastq (recoverDataNoComment): returns bytes or (bytes, bytes)."""
    outs = ([], []) if pe else ([],)
    io = so = 0
    for i in range(blk.nreads):
        il, sl = int(blk.idlen[i]), int(blk.seqlen[i])
        rec = b"@" + blk.ids[io:io + il].tobytes() + b"\n" + blk.seq[so:so + sl].tobytes() + b"\n+\n" + \
              blk.qual[so:so + sl].tobytes() + b"\n"
        outs[i % 2 if pe else 0].append(rec)
        io += il
        so += sl
    res = tuple(b"".join(o) for o in outs)
    return res if pe else res[0]


# ------------------------------------------------------------------ reference binary runner
def have_ref():
    build()
    return os.path.exists(REF_BIN)


def run_ref(args, cwd, lossy_shim=False, config=None, timeout=3600, fake_size=None):
    """Run the unmodified reference executable in `cwd` (it reads ./seqarc.config from there)."""
    env = dict(os.environ)
    if lossy_shim:
        env["LD_PRELOAD"] = NLGUARD
    if fake_size:      # make the reference take its usemodel = 1 path on a small input (oracle/sizefake.c)
        env["LD_PRELOAD"] = SIZEFAKE
        env["SQ_FAKE_SIZE"] = str(int(fake_size))
    cfg = os.path.join(cwd, "seqarc.config")
    if config:
        with open(cfg, "w") as f:
            for k, v in config.items():
                f.write("%s:%s\n" % (k, v))
    elif os.path.exists(cfg):
        os.remove(cfg)
    r = subprocess.run([REF_BIN] + list(args), cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=timeout)
    return r.returncode, r.stdout.decode("latin1")
