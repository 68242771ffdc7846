"""GPU: the device-resident entry points (sq_encode_blocks_dev / sq_decode_blocks_dev): inputs, records and outputs all
live in HBM (what bench.py's `value` leg times); records must equal the oracle's, the round trip must give back the
input, and a damaged record must be reported per block without hurting its neighbours."""
import ctypes as C

import numpy as np
import pytest

from tests import common
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).cuda()


def test_device_resident_round_trip_and_damaged_record():
    import torch
    from fastqueeze_b200 import _lib
    L = _lib.lib()
    blocks = [common.synth_block(kind="illumina_pe150", n_templates=1200, seed=70 + i, n_rate=0.003) for i in range(3)]
    p = _lib.SqParams()
    p.device, p.md5, p.slevel, p.qlevel, p.id_rule = 0, 1, 3, 2, 0
    p.workspace_bytes = 2 << 30
    p.stream = torch.cuda.current_stream().cuda_stream
    h = C.c_void_p()
    assert L.sq_create(C.byref(p), C.byref(h)) == 0, L.sq_last_error(None)
    try:
        n = len(blocks)
        keep = []
        ins, outs = (_lib.SqBlockIn * n)(), (_lib.SqBlockOut * n)()
        for i, b in enumerate(blocks):
            t = {k: _dev(torch, getattr(b, k)) for k in ("ids", "seq", "qual", "idlen", "seqlen")}
            rec = torch.zeros(int(b.ids.size + 2 * b.seq.size) + 4096, dtype=torch.uint8, device="cuda")
            keep.append((t, rec))
            ins[i].nreads, ins[i].longflag = int(b.seqlen.size), int(b.longflag)
            ins[i].ids, ins[i].seq, ins[i].qual = t["ids"].data_ptr(), t["seq"].data_ptr(), t["qual"].data_ptr()
            ins[i].idlen, ins[i].seqlen = t["idlen"].data_ptr(), t["seqlen"].data_ptr()
            ins[i].idbytes, ins[i].seqbytes = int(b.ids.size), int(b.seq.size)
            outs[i].data, outs[i].cap = rec.data_ptr(), rec.numel()
        assert L.sq_encode_blocks_dev(h, ins, n, outs) == 0, L.sq_last_error(h)
        torch.cuda.synchronize()
        o = po.Oracle()
        for i, b in enumerate(blocks):
            assert outs[i].status == 0
            assert keep[i][1][:outs[i].len].cpu().numpy().tobytes() == o.encode_block(b)

        def decode(recs):
            decs = (_lib.SqBlockDec * n)()
            ptrs, lens = (C.c_void_p * n)(), (C.c_uint64 * n)()
            outbuf = []
            for i, b in enumerate(blocks):
                d = {k: torch.zeros_like(keep[i][0][k]) for k in ("ids", "seq", "qual", "idlen", "seqlen")}
                outbuf.append(d)
                ptrs[i], lens[i] = recs[i][0].data_ptr(), recs[i][1]
                decs[i].ids, decs[i].idcap = d["ids"].data_ptr(), d["ids"].numel()
                decs[i].seq, decs[i].qual, decs[i].seqcap = d["seq"].data_ptr(), d["qual"].data_ptr(), d["seq"].numel()
                decs[i].idlen, decs[i].seqlen, decs[i].readcap = d["idlen"].data_ptr(), d["seqlen"].data_ptr(), int(b.seqlen.size)
            r = L.sq_decode_blocks_dev(h, ptrs, lens, n, decs)
            torch.cuda.synchronize()
            return r, decs, outbuf

        good = [(keep[i][1], int(outs[i].len)) for i in range(n)]
        r, decs, outbuf = decode(good)
        assert r == 0, L.sq_last_error(h)
        for i, b in enumerate(blocks):
            assert decs[i].status == 0 and decs[i].nreads == b.seqlen.size
            for k in ("ids", "seq", "qual", "idlen", "seqlen"):
                assert torch.equal(outbuf[i][k], keep[i][0][k]), (i, k)
        # block 1: a record cut short (the tag walk runs on the device); block 2: payload damage (an MD5 mismatch)
        cut = keep[1][1][:40].clone()
        bad = keep[2][1].clone()
        bad[int(outs[2].len) // 2] ^= 0x55
        r, decs, outbuf = decode([good[0], (cut, 40), (bad, good[2][1])])
        assert r != 0
        assert decs[0].status == 0 and torch.equal(outbuf[0]["qual"], keep[0][0]["qual"])
        assert decs[1].status == -6
        assert decs[2].status in (-8, -9, -10, -6, -4)
    finally:
        L.sq_destroy(h)
