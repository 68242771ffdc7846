"""Host-side mirror of the reference's coder seam, on top of the C ABI.

The reference has one `EncapFqzComp` per worker thread with
`doFqzEncode(SeqArcMemBuf*, DebugInfo*)` (@0x42d2d0 of /root/reference/SeqArc-1.6) and
`doDecode(SeqArcMemBuf*, DebugInfo*)` (@0x42eb60).  `EncapFqzComp` below keeps those names and
meanings, but takes a LIST of `SeqArcMemBuf` blocks per call (blocks are independent: every model
resets per block) and runs them on one B200 through libseqarc_b200.so.  Nothing here computes on
the CPU; arrays are numpy views handed to the library as plain pointers.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import SeqArcB200Error, SqBlockDec, SqBlockIn, SqBlockOut, SqParams


@dataclass
class SeqArcMemBuf:
    """One block as the reference's SeqArcMemBuf holds it: concatenated buffers, PE
    interleaved R1,R2,...; `out` receives the block record on encode."""
    ids: np.ndarray      # uint8, IDs without '@'          (membuf+0x80)
    idlen: np.ndarray    # uint16 per read                 (membuf+0xc0)
    seq: np.ndarray      # uint8                           (membuf+0x88)
    qual: np.ndarray     # uint8                           (membuf+0x90)
    seqlen: np.ndarray   # int32 per read                  (membuf+0xd0)
    out: bytes = b""     # block record "tag 1 | size | streams" (membuf+0x70)
    status: int = 0

    @property
    def nreads(self):
        return int(self.seqlen.shape[0])

    @property
    def longflag(self):
        """membuf+0x2: a read longer than 65,535 bases (the Len stream then carries four length bytes and the
        archive's block index records the flag in bit 0 of the length field)."""
        return int(self.seqlen.shape[0] > 0 and int(np.max(self.seqlen)) > 65535)


@dataclass
class SeqArcParam:
    """The SeqArcParam fields the hot path reads."""
    usemodel: int = 0
    lossy: int = 0
    lossy_r: float = 0.0
    md5: int = 1
    Slevel: int = 3
    Qlevel: int = 2
    id_rule: int = 0          # idbin[0]: 1 = rule-based ID path
    modelbuf: bytes = None    # snapshot when usemodel
    ref_mode: int = 0         # 1 (SE) / 2 (PE): reference index given, nothing mapped (doAlignEncode framing)


@dataclass
class DebugInfo:
    """Per-stage device milliseconds of the last call (DebugInfo @0x420080 holds wall seconds)."""
    stages: dict = field(default_factory=dict)
    launches: int = 0
    wave_blocks: int = 0


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


class EncapFqzComp:
    def __init__(self, param: SeqArcParam = None, device=0, workspace_bytes=0, stream=None):
        self.param = param or SeqArcParam()
        p = SqParams()
        p.device = device
        p.usemodel = self.param.usemodel
        p.lossy = self.param.lossy
        p.lossy_r = self.param.lossy_r
        p.md5 = self.param.md5
        p.slevel = self.param.Slevel
        p.qlevel = self.param.Qlevel
        p.id_rule = self.param.id_rule
        p.ref_mode = self.param.ref_mode
        p.workspace_bytes = workspace_bytes
        p.stream = stream
        self._L = _lib.lib()
        h = C.c_void_p()
        r = self._L.sq_create(C.byref(p), C.byref(h))
        if r != 0:
            raise SeqArcB200Error(r, self._L.sq_last_error(None).decode())
        self._h = h
        if self.param.usemodel:
            self.ReadModelFormMem(self.param.modelbuf)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sq_destroy(self._h)
            self._h = None

    __del__ = close

    def _err(self, r):
        return SeqArcB200Error(r, self._L.sq_last_error(self._h).decode())

    # ReadModelFormMem @0x420fc0
    def ReadModelFormMem(self, blob: bytes):
        if blob is None:
            raise ValueError("usemodel = 1 needs a model snapshot")
        a = np.frombuffer(blob, dtype=np.uint8)
        r = self._L.sq_load_model(self._h, a.ctypes.data, a.shape[0])
        if r != 0:
            raise self._err(r)

    def model_size(self):
        return int(self._L.sq_model_size(self._h))

    # encodeForModel @0x4228e0 + SaveModelToMem @0x420c00: train the snapshot on the sample blocks
    def encodeForModel(self, membufs) -> bytes:
        n = len(membufs)
        ins = (SqBlockIn * n)()
        keep = []
        for i, m in enumerate(membufs):
            ids, seq, qual = _c(m.ids, np.uint8), _c(m.seq, np.uint8), _c(m.qual, np.uint8)
            idlen, seqlen = _c(m.idlen, np.uint16), _c(m.seqlen, np.int32)
            keep.append((ids, seq, qual, idlen, seqlen))
            ins[i].nreads = seqlen.shape[0]
            ins[i].longflag = m.longflag
            ins[i].ids, ins[i].seq, ins[i].qual = ids.ctypes.data, seq.ctypes.data, qual.ctypes.data
            ins[i].idlen, ins[i].seqlen = idlen.ctypes.data, seqlen.ctypes.data
            ins[i].idbytes, ins[i].seqbytes = ids.shape[0], seq.shape[0]
        blob = np.empty(self.model_size(), np.uint8)
        r = self._L.sq_train_model(self._h, ins, n, blob.ctypes.data, blob.shape[0])
        if r != 0:
            raise self._err(r)
        return blob.tobytes()

    def _debug(self, dbg):
        if dbg is None:
            return
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = self._L.sq_last_timings(self._h, names, ms, 64)
        dbg.stages = {names[i].decode(): float(ms[i]) for i in range(n)}
        dbg.launches = int(self._L.sq_last_launches(self._h))
        dbg.wave_blocks = int(self._L.sq_last_wave_blocks(self._h))

    # doFqzEncode @0x42d2d0 over a batch of blocks; fills membuf.out, returns the record lengths
    def doFqzEncode(self, membufs, dbg: DebugInfo = None, raise_on_error=True):
        n = len(membufs)
        ins = (SqBlockIn * n)()
        outs = (SqBlockOut * n)()
        keep = []
        for i, m in enumerate(membufs):
            ids, seq, qual = _c(m.ids, np.uint8), _c(m.seq, np.uint8), _c(m.qual, np.uint8)
            idlen, seqlen = _c(m.idlen, np.uint16), _c(m.seqlen, np.int32)
            assert seq.shape[0] == qual.shape[0]
            cap = int(ids.shape[0] + 2 * seq.shape[0] + 8 * seqlen.shape[0] + 256)
            ob = np.empty(cap, dtype=np.uint8)
            keep.append((ids, seq, qual, idlen, seqlen, ob))
            ins[i].nreads = seqlen.shape[0]
            ins[i].longflag = m.longflag
            ins[i].ids, ins[i].seq, ins[i].qual = ids.ctypes.data, seq.ctypes.data, qual.ctypes.data
            ins[i].idlen, ins[i].seqlen = idlen.ctypes.data, seqlen.ctypes.data
            ins[i].idbytes, ins[i].seqbytes = ids.shape[0], seq.shape[0]
            outs[i].data, outs[i].cap = ob.ctypes.data, cap
        r = self._L.sq_encode_blocks(self._h, ins, n, outs)
        self._debug(dbg)
        lens = []
        for i, m in enumerate(membufs):
            m.status = int(outs[i].status)
            m.out = keep[i][5][:outs[i].len].tobytes() if outs[i].status == 0 else b""
            lens.append(int(outs[i].len))
        if r != 0 and raise_on_error:
            raise self._err(r)
        return lens

    # doDecode @0x42eb60 over a batch of block records; returns SeqArcMemBuf list
    def doDecode(self, records, dbg: DebugInfo = None, idcap=None, seqcap=None, readcap=None, raise_on_error=True,
                 longflags=None):
        n = len(records)
        recs = (C.c_void_p * n)()
        lens = (C.c_uint64 * n)()
        decs = (SqBlockDec * n)()
        keep = []
        for i, rec in enumerate(records):
            ra = np.frombuffer(rec, dtype=np.uint8)
            ic = idcap or (1 << 20) + 8 * len(rec)
            sc = seqcap or (1 << 20) + 8 * len(rec)
            rcap = readcap or (1 << 12) + len(rec)
            bufs = (np.empty(ic, np.uint8), np.empty(sc, np.uint8), np.empty(sc, np.uint8),
                    np.empty(rcap, np.uint16), np.empty(rcap, np.int32), ra)
            keep.append(bufs)
            recs[i] = ra.ctypes.data
            lens[i] = ra.shape[0]
            decs[i].ids, decs[i].idcap = bufs[0].ctypes.data, ic
            decs[i].seq, decs[i].qual, decs[i].seqcap = bufs[1].ctypes.data, bufs[2].ctypes.data, sc
            decs[i].idlen, decs[i].seqlen, decs[i].readcap = bufs[3].ctypes.data, bufs[4].ctypes.data, rcap
            decs[i].longflag = int(longflags[i]) if longflags is not None else 0
        r = self._L.sq_decode_blocks(self._h, recs, lens, n, decs)
        self._debug(dbg)
        res = []
        for i in range(n):
            d = decs[i]
            b = keep[i]
            if d.status == 0:
                res.append(SeqArcMemBuf(b[0][:d.idbytes].copy(), b[3][:d.nreads].copy(), b[1][:d.seqbytes].copy(),
                                        b[2][:d.seqbytes].copy(), b[4][:d.nreads].copy(), status=0))
            else:
                res.append(SeqArcMemBuf(np.zeros(0, np.uint8), np.zeros(0, np.uint16), np.zeros(0, np.uint8),
                                        np.zeros(0, np.uint8), np.zeros(0, np.int32), status=int(d.status)))
        if r != 0 and raise_on_error:
            raise self._err(r)
        return res

    # rblock @0x426c10
    def rblock(self, qual: np.ndarray, r: float) -> np.ndarray:
        q = np.ascontiguousarray(qual, dtype=np.uint8).copy()
        rc = self._L.sq_rblock(self._h, q.ctypes.data, q.shape[0], r)
        if rc != 0:
            raise self._err(rc)
        return q

    # Aligned-mode side streams (compressOrder @0x424b70, compressPERelation @0x422be0, compressAlignInfo_* @0x425d70-0x426c0f
    # and their decompress* twins): all streams of one call are coded in one launch.
    def compressSideStreams(self, streams):
        """streams: [(N, uint8 array)] -> [bytes], the bare range-coder output of each array through a fresh SIMPLE_MODEL<N>."""
        n = len(streams)
        ss = (_lib.SqSideStream * n)()
        keep = []
        for i, (model, sym) in enumerate(streams):
            a = np.ascontiguousarray(sym, dtype=np.uint8)
            out = np.empty(a.shape[0] + 16, np.uint8)
            keep.append((a, out))
            ss[i].model, ss[i].nsym, ss[i].sym = model, a.shape[0], a.ctypes.data if a.shape[0] else None
            ss[i].bytes, ss[i].cap = out.ctypes.data, out.shape[0]
        rc = self._L.sq_side_encode(self._h, ss, n)
        if rc != 0:
            raise self._err(rc)
        return [keep[i][1][:ss[i].len].tobytes() for i in range(n)]

    def decompressSideStreams(self, streams):
        """streams: [(N, bytes, nsym)] -> [uint8 array of nsym symbols]."""
        n = len(streams)
        ss = (_lib.SqSideStream * n)()
        keep = []
        for i, (model, data, nsym) in enumerate(streams):
            b = np.frombuffer(data, np.uint8).copy()
            out = np.empty(nsym, np.uint8)
            keep.append((b, out))
            ss[i].model, ss[i].nsym, ss[i].sym = model, nsym, out.ctypes.data if nsym else None
            ss[i].bytes, ss[i].cap = b.ctypes.data, b.shape[0]
        rc = self._L.sq_side_decode(self._h, ss, n)
        if rc != 0:
            raise self._err(rc)
        return [k[1] for k in keep]

    # Aligned mode, decode side: mapped reads rebuilt from the reference (getRealPos @0x42de30, AlignInfoToSeq @0x42e020)
    def set_reference(self, fasta: bytes):
        """The reference FASTA the archive was written against: header lines dropped, sequences concatenated, uploaded once."""
        seq = np.frombuffer(b"".join(l for l in fasta.split(b"\n") if l and not l.startswith(b">")), np.uint8).copy()
        rc = self._L.sq_set_reference(self._h, seq.ctypes.data, seq.shape[0])
        if rc != 0:
            raise self._err(rc)

    def alignRebuild(self, blocks):
        """blocks: [dict(seqlen, seq, pe, order, pos, mis, rev, cigal, cigav[, perel, insbits])] -> [seq array] with the mapped
        reads' bases filled in (Pos / CigaL are bit arrays, one bit per byte; the other reads' bytes are left as given)."""
        n = len(blocks)
        bl = (_lib.SqAlignBlock * n)()
        keep = []
        for i, b in enumerate(blocks):
            a = {k: np.ascontiguousarray(b.get(k, np.zeros(0, np.uint8)), np.uint8) for k in ("order", "pos", "mis", "rev", "cigal", "cigav", "perel")}
            a["seqlen"] = np.ascontiguousarray(b["seqlen"], np.int32)
            a["seq"] = np.ascontiguousarray(b["seq"], np.uint8).copy()
            keep.append(a)
            ptr = lambda k: a[k].ctypes.data if a[k].shape[0] else None      # noqa: E731
            bl[i].nreads, bl[i].pe = a["seqlen"].shape[0], int(bool(b.get("pe", 0)))
            bl[i].seqlen, bl[i].order, bl[i].pos, bl[i].mis, bl[i].rev = ptr("seqlen"), ptr("order"), ptr("pos"), ptr("mis"), ptr("rev")
            bl[i].cigal, bl[i].cigav, bl[i].perel = ptr("cigal"), ptr("cigav"), ptr("perel")
            bl[i].npos, bl[i].nalign, bl[i].ncigal = a["pos"].shape[0], a["mis"].shape[0], a["cigal"].shape[0]
            bl[i].ncigav, bl[i].nperel, bl[i].insbits = a["cigav"].shape[0], a["perel"].shape[0], int(b.get("insbits", 0))
            bl[i].seq, bl[i].seqbytes = ptr("seq"), a["seq"].shape[0]
        rc = self._L.sq_align_rebuild(self._h, bl, n)
        if rc != 0:
            raise self._err(rc)
        return [k["seq"] for k in keep]

    # test hook
    def debug_triples(self, which, block=0):
        cnt = np.zeros(3, np.uint64)
        got = C.c_uint64()
        r = self._L.sq_debug_fetch(self._h, 3, block, cnt.ctypes.data, 24, C.byref(got))
        if r != 0:
            raise self._err(r)
        out = np.empty(int(cnt[which]), np.uint64)
        r = self._L.sq_debug_fetch(self._h, which, block, out.ctypes.data, out.nbytes, C.byref(got))
        if r != 0:
            raise self._err(r)
        return out
