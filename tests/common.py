"""Shared helpers of the test-suite: golden cases -> blocks + reference records."""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import arcwalk, pyoracle as po  # noqa: E402
from tools import synth_fastq as sf  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
ALIGNED = sorted(k for k in MANIFEST if MANIFEST[k].get("aligned"))      # blocks with mapped reads: side-stream coders only
CASES = sorted(k for k in MANIFEST if k not in ALIGNED)


def case_fastq(name):
    m = MANIFEST[name]
    if m["src"] == "err":
        return tuple(gzip.open(os.path.join(GOLD, "err_pe1000_%d.fq.gz" % k), "rb").read() for k in (1, 2))
    blk = sf.make_block(**m["src"])
    t = sf.to_fastq(blk)
    t = t if m["pe"] else (t,)
    if m.get("plus"):       # '+' lines that repeat the ID
        from tools.make_golden import plus_lines
        t = tuple(plus_lines(x) for x in t)
    return t


def preprocess_sample(fastq, blocksize: int, max_blocks: int = 5, fastq2=None):
    """The reads SeqArcPreProcess::doPreProcess @0x4153b0 trains the model snapshot on: up to five head
    blocks, each taking whole records until it holds at least `blocksize` bytes (BlockSize / 2 of mate 1
    for paired-end, the same number of records from mate 2).  Measured against the reference with the
    size-faking shim: 5 x 4298 reads of 244 B (SE) / 5 x 1461 pairs of 359 B per mate (PE) at BlockSize 1 MiB,
    i.e. one more record per block than the encoder's blocks hold."""
    recs = po.parse_fastq(fastq)
    recs2 = po.parse_fastq(fastq2) if fastq2 is not None else None
    quota = blocksize // 2 if recs2 is not None else blocksize
    out, i = [], 0
    for _ in range(max_blocks):
        if i >= len(recs):
            break
        size, j = 0, i
        while j < len(recs) and size < quota:
            size += len(recs[j][0]) + len(recs[j][1]) + len(recs[j][2]) + 6
            j += 1
        if recs2 is None:
            out.append(po.block_from_records(recs[i:j]))
        else:
            inter = [None] * (2 * (j - i))
            inter[0::2], inter[1::2] = recs[i:j], recs2[i:j]
            out.append(po.block_from_records(inter))
        i = j
    return out


_MODEL_CACHE = {}


def case_model(name):
    """Model snapshot of a usemodel golden case, rebuilt by the oracle's training twins and checked against
    the md5 of the blob the reference wrote (the 27 MB blob itself is not committed)."""
    import hashlib
    if name not in _MODEL_CACHE:
        m = MANIFEST[name]
        bs = int((m["config"] or {}).get("BlockSize(M)", 50)) << 20
        fq = case_fastq(name)
        o = po.Oracle(usemodel=0, id_rule=0)
        blob = o.train(preprocess_sample(fq[0], bs, fastq2=fq[1] if m["pe"] else None))
        assert hashlib.md5(blob).hexdigest() == m["model_md5"], "oracle-trained snapshot differs from the reference's"
        _MODEL_CACHE[name] = blob
    return _MODEL_CACHE[name]


def case_archive(name):
    m = MANIFEST[name]
    if not m.get("usemodel"):
        return open(os.path.join(GOLD, name + ".arc"), "rb").read()
    stripped = open(os.path.join(GOLD, name + ".arc_nomodel"), "rb").read()
    blob = bytes([0x81, 0x82, 0x01, 0x00]) + case_model(name)
    n = len(blob) | (1 << 28)
    cut = m["model_cut"]
    return stripped[:cut] + bytes([0x86]) + n.to_bytes(4, "big") + blob + stripped[cut:]


def case_params(name):
    """Coder options of a golden case as kwargs of pyoracle.Oracle / fastqueeze_b200.SeqArcParam."""
    import hashlib
    m = MANIFEST[name]
    data = case_archive(name)
    assert hashlib.md5(data).hexdigest() == m["arc_md5"]
    arc = arcwalk.parse_arc(data)
    cfg = m["config"] or {}
    idbin = arc["trailer"]["params"][15]
    kw = dict(usemodel=int(m.get("usemodel", 0)), lossy=0, lossy_r=0.0, md5=0 if "-n" in m["opts"] else 1,
              slevel=int(cfg.get("Slevel", 3)), qlevel=int(cfg.get("Qlevel", 2)), id_rule=int(idbin[0]))
    if kw["usemodel"]:
        kw["model"] = arc["trailer"]["model"][4:]
    if m.get("ref"):        # reference index given, nothing mapped: doAlignEncode framing
        kw["ref_mode"] = 2 if m["pe"] else 1
        assert arc["trailer"][8].hex() == m["ref_md5"] and arc["trailer"]["params"][1] == b"\x00"
    if "-l" in m["opts"]:
        kw["lossy"] = 1
        kw["lossy_r"] = float(m["opts"][m["opts"].index("-l") + 1])
    return kw, arc


def case_blocks(name):
    """[(pyoracle.Block, reference record bytes)] for every block of the golden archive."""
    m = MANIFEST[name]
    kw, arc = case_params(name)
    fq = case_fastq(name)
    idx = arcwalk.block_index(arc, m["pe"])
    out = []
    nb = len(arc["blocks"])
    for i, b in enumerate(arc["blocks"]):
        if nb == 1:
            parts = fq
        elif m["pe"]:   # cut both source texts at the offsets the reference recorded in its block index
            s1, s2 = idx[i]["offs"][1], idx[i]["offs"][2]
            parts = (fq[0][s1:s1 + idx[i]["src1"]], fq[1][s2:s2 + idx[i]["src2"]])
        else:
            s = idx[i]["offs"][1]
            e = idx[i + 1]["offs"][1] if i + 1 < nb else len(fq[0])
            parts = (fq[0][s:e],)
        blk = po.block_from_fastq(*parts)
        assert idx[i]["long"] == blk.longflag          # bit 0 of the index length field: a read > 65,535 bases
        out.append((blk, b["raw"]))
    return kw, out


def to_membuf(blk):
    from fastqueeze_b200 import SeqArcMemBuf
    return SeqArcMemBuf(blk.ids, blk.idlen, blk.seq, blk.qual, blk.seqlen)


def to_param(kw, model=None):
    from fastqueeze_b200 import SeqArcParam
    return SeqArcParam(usemodel=kw["usemodel"], lossy=kw["lossy"], lossy_r=kw["lossy_r"], md5=kw["md5"],
                       Slevel=kw["slevel"], Qlevel=kw["qlevel"], id_rule=kw["id_rule"],
                       modelbuf=model if model is not None else kw.get("model"), ref_mode=kw.get("ref_mode", 0))


def synth_block(**kw):
    d = sf.make_block(**kw)
    return po.Block(d["ids"], d["idlen"], d["seq"], d["qual"], d["seqlen"])
