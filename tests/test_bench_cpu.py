"""CPU: the pieces of bench.py that decide WHAT is measured -- the synthetic workload is valid FASTQ of the size the metric's
numerator counts, the named configurations exist, the oracle codes a generated block and decodes it back, and the committed
traffic capture is the one the headline line looks up."""
import json
import os
import sys

import numpy as np
import pytest

from tests.common import ROOT
from oracle import pyoracle as po

sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("kind", ["pe150", "se100_q8"])
def test_generated_block_is_fastq_of_the_counted_size(kind):
    import torch
    shape = bench.Shape(kind)
    units = 64
    blk = bench.to_host(torch, bench.gen_block_device(torch, torch.device("cpu"), shape, 3, units, 7))
    texts = [bench.fastq_text(blk, shape, k).tobytes() for k in range(shape.reads_per_unit)]
    assert sum(len(t) for t in texts) == shape.block_fastq_bytes(units)           # the bytes `value` divides by the time
    recs = [po.parse_fastq(t) for t in texts]
    assert all(len(r) == units for r in recs)
    assert all(len(s) == shape.L == len(q) for r in recs for (_, s, q) in r)
    # the block handed to the coder is those records (paired-end: interleaved), and the oracle round-trips it
    want = po.block_from_fastq(*texts)
    assert want.seq.tobytes() == blk["seq"].tobytes() and want.qual.tobytes() == blk["qual"].tobytes()
    assert want.ids.tobytes() == blk["ids"].tobytes()
    o = po.Oracle()
    back = o.decode_block(o.encode_block(want))
    assert back.seq.tobytes() == want.seq.tobytes() and back.qual.tobytes() == want.qual.tobytes()


def test_named_configurations_and_block_geometry():
    assert set(bench.CONFIGS) == {"pe150", "pe150_lossy", "se100_q8", "pe150_ref"}
    pe = bench.Shape("pe150")
    assert pe.units_per_block == 73020                      # pairs the reference's block cutter puts in a 50-MiB block of this shape
    assert -(-bench.CONFIGS["pe150"]["units"] // pe.units_per_block) == 274
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "configs" in base and len(base["configs"]) >= 5


def test_committed_traffic_capture_is_the_one_the_headline_line_quotes():
    traffic, src = bench.measured_traffic("dec.decode", "pe150", 274, 1, 1)
    assert traffic and traffic > 35.8e9                      # DRAM bytes per launch, above the algorithmic bytes by construction
    assert os.path.exists(os.path.join(ROOT, src.split(" ")[0]))
    assert bench.measured_traffic("dec.decode", "pe150", 274, 2, 1) == (None, None)     # no capture for other shapes: reported as null
