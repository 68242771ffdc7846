"""Profiling driver: one encode (and optionally decode) of a few full-size synthetic blocks through
the C ABI.  Meant to be run under ncu on the GPU box (see profiles/README.md)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastqueeze_b200 as fq  # noqa: E402
from tools import synth_fastq as sf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=1)
ap.add_argument("--pairs", type=int, default=73020)
ap.add_argument("--kind", default="illumina_pe150")
ap.add_argument("--decode", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--ws-gb", type=float, default=0)
a = ap.parse_args()

blks = []
for i in range(a.blocks):
    if i < 4:
        d = sf.make_block(a.kind, n_templates=a.pairs, seed=100 + i, first=i * a.pairs)
    blks.append(fq.SeqArcMemBuf(d["ids"], d["idlen"], d["seq"], d["qual"], d["seqlen"]))   # blocks >= 4 repeat block 3 (timing only)
codec = fq.EncapFqzComp(fq.SeqArcParam(), workspace_bytes=int(a.ws_gb * (1 << 30)) if a.ws_gb else int(1.3 * (1 << 30)) * a.blocks + (1 << 30))
dbg = fq.DebugInfo()
for _ in range(a.reps):
    t0 = time.time()
    codec.doFqzEncode(blks, dbg)
    print("encode %.3fs" % (time.time() - t0), {k: round(v, 2) for k, v in dbg.stages.items()})
if a.decode:
    t0 = time.time()
    out = codec.doDecode([b.out for b in blks], dbg)
    print("decode %.3fs" % (time.time() - t0), {k: round(v, 2) for k, v in dbg.stages.items()})
    assert out[0].qual.tobytes() == blks[0].qual.tobytes() and out[0].seq.tobytes() == blks[0].seq.tobytes()
