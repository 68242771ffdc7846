#!/usr/bin/env python
"""bench.py -- entropy-stage compress + decompress throughput of the B200 SeqArc path.

A "step" is one pass of the hot path over the job's blocks: every block is compressed (doFqzEncode seam) and the
records are decompressed again (doDecode seam).  The job is one of BASELINE.json's configs (--config):

  pe150        configs[1]  synthetic Illumina 150 bp PE, 20 M pairs, no reference, lossless      (default)
  pe150_lossy  configs[2]  the same with R-Block lossy qualities, -l 1.15
  se100_q8     configs[3]  synthetic 100 bp SE, 100 M reads, 8-level binned qualities
  pe150_ref    configs[4]  configs[1] written against a reference index no read maps to (doAlignEncode framing)

  value   whole-job MB/s of input FASTQ, inputs already resident in HBM (device pointers through
          sq_encode_blocks_dev / sq_decode_blocks_dev), CUDA events on the launching stream
  e2e     the same through the host-buffer C-ABI calls (sq_encode_blocks / sq_decode_blocks):
          pinned host buffers in, host records out, host<->device copies inside the timed region
  --impl reference   the UNMODIFIED reference binary (oracle/_ref/SeqArc-1.6) on the host cores,
          -t <cores>, on a bounded sample of the same workload

Multi-GPU (torchrun, one rank per GPU): the job is FIXED and its blocks are sharded over the ranks in contiguous
ranges ("scaling": "strong") -- blocks are independent units (models reset per block), so there is no data-path
collective; torch.distributed carries the barrier, the max-over-ranks time and the host-side ordered gather of the
record lengths that fixes every record's offset in the archive (fastqueeze_b200/shard.py).  Because the coder
chains of one block are serial, a GPU's time hardly depends on how many blocks it holds: the weak-scaling figure
(every rank codes a whole job of its own) is measured as well and reported as `replica_MBps`.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK_BYTES = 50 << 20            # SeqArc default BlockSize = 50 MiB of FASTQ per block (25 MiB per mate file)
ID_PREFIX = b"A00123:45:HXXXXDSXX:1:"
QLV8 = [2, 6, 15, 22, 27, 33, 37, 40]
QP8 = [.01, .02, .03, .04, .05, .10, .35, .40]


class Shape:
    """Geometry of one synthetic FASTQ shape: fixed-width records, so a block is a handful of matrix ops."""

    def __init__(self, kind):
        self.kind = kind
        self.pe = kind == "pe150"
        self.L = 150 if self.pe else 100
        self.id_w = len(ID_PREFIX) + 4 + 1 + 5 + 1 + 5 + (len(b" 1:N:0:ACGTACGT") if self.pe else 0)
        self.rec_bytes = (1 + self.id_w + 1) + (self.L + 1) + 2 + (self.L + 1)       # FASTQ text bytes per read
        # units = pairs (PE: BlockSize / 2 per mate file) or reads (SE) the block cutter puts in one block
        self.units_per_block = BLOCK_BYTES // (2 * self.rec_bytes) if self.pe else BLOCK_BYTES // self.rec_bytes
        self.reads_per_unit = 2 if self.pe else 1

    def block_fastq_bytes(self, units):
        return units * self.reads_per_unit * self.rec_bytes


CONFIGS = {
    "pe150": dict(shape="pe150", units=20_000_000, lossy=0, ref_mode=0, baseline="configs[1]",
                  text="synthetic Illumina 150 bp PE, no reference, lossless"),
    "pe150_lossy": dict(shape="pe150", units=20_000_000, lossy=1, ref_mode=0, baseline="configs[2]",
                        text="synthetic Illumina 150 bp PE, no reference, lossy quality R-Block -l 1.15"),
    "se100_q8": dict(shape="se100_q8", units=100_000_000, lossy=0, ref_mode=0, baseline="configs[3]",
                     text="synthetic 100 bp SE, binned Q-scores (8-level), no reference, lossless"),
    "pe150_ref": dict(shape="pe150", units=20_000_000, lossy=0, ref_mode=2, baseline="configs[4]",
                      text="synthetic Illumina 150 bp PE written against a random reference index (no read maps: unmapped-read + quality streams)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="pe150", choices=sorted(CONFIGS))
    ap.add_argument("--blocks", type=int, default=0, help="blocks of the job (0 = the full config)")
    ap.add_argument("--cpu-sample-blocks", type=int, default=0, help="blocks of the CPU baseline sample (0 = max(16, 2 x host cores))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-replica", action="store_true", help="N > 1: skip the weak-scaling (one whole job per rank) measurement")
    ap.add_argument("--verbose", action="store_true")
    return ap.parse_args()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic blocks, generated on the device (SURVEY 8d inputs 2 and 4)
def gen_block_device(torch, dev, shape, block_index, units, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1_000_003 + block_index)
    rpu = shape.reads_per_unit
    n = rpu * units
    L = shape.L
    first = block_index * units
    idx = torch.arange(first, first + units, device=dev, dtype=torch.int64)
    tile = 1101 + (idx // 4096) % 8000
    x = torch.randint(1000, 32000, (units,), device=dev, generator=g)
    y = torch.randint(1000, 32000, (units,), device=dev, generator=g)

    def digits(v, w):
        p = torch.tensor([10 ** k for k in range(w - 1, -1, -1)], device=dev, dtype=torch.int64)
        return ((v[:, None] // p[None, :]) % 10 + 48).to(torch.uint8)

    def const(s):
        return torch.tensor(list(s), device=dev, dtype=torch.uint8)[None, :].expand(units, len(s))

    def ids(mate):
        cols = [const(ID_PREFIX), digits(tile, 4), const(b":"), digits(x, 5), const(b":"), digits(y, 5)]
        if mate:
            cols.append(const(b" %d:N:0:ACGTACGT" % mate))
        return torch.cat(cols, dim=1)

    if shape.pe:
        m1, m2 = ids(1), ids(2)
        idm = torch.stack([m1, m2], dim=1).reshape(n, shape.id_w).contiguous()
    else:
        idm = ids(0).contiguous()
    assert idm.shape[1] == shape.id_w
    acgt = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    seq = acgt[torch.randint(0, 4, (n, L), device=dev, generator=g)].contiguous()
    if shape.kind == "se100_q8":      # 8-level binned qualities {2,6,15,22,27,33,37,40} with p = {.01,.02,.03,.04,.05,.10,.35,.40}
        cdf = torch.tensor(np.cumsum(QP8)[:-1], device=dev, dtype=torch.float32)
        u = torch.rand((n, L), device=dev, generator=g)
        lv = torch.tensor(QLV8, device=dev, dtype=torch.uint8)
        qual = (lv[torch.bucketize(u, cdf)] + 33).contiguous()
    else:                             # clipped N(36, 5)
        q = torch.empty((n, L), device=dev, dtype=torch.float32).normal_(36.0, 5.0, generator=g)
        qual = (q.round().clamp_(2, 40) + 33).to(torch.uint8).contiguous()
    return dict(ids=idm.reshape(-1), seq=seq.reshape(-1), qual=qual.reshape(-1),
                idlen=torch.full((n,), shape.id_w, device=dev, dtype=torch.int16),
                seqlen=torch.full((n,), L, device=dev, dtype=torch.int32), nreads=n)


def fastq_text(blk_host, shape, mate):
    """Render one mate (PE) / the reads (SE) of a (host, numpy) block as FASTQ text, vectorised (fixed-width records)."""
    rpu, L, W = shape.reads_per_unit, shape.L, shape.id_w
    nr = blk_host["nreads"]
    sel = slice(mate, None, rpu)
    ids = blk_host["ids"].reshape(nr, W)[sel]
    seq = blk_host["seq"].reshape(nr, L)[sel]
    qual = blk_host["qual"].reshape(nr, L)[sel]
    n = ids.shape[0]
    rec = np.empty((n, shape.rec_bytes), dtype=np.uint8)
    o = 0
    rec[:, o] = ord("@"); o += 1
    rec[:, o:o + W] = ids; o += W
    rec[:, o] = 10; o += 1
    rec[:, o:o + L] = seq; o += L
    rec[:, o] = 10; rec[:, o + 1] = ord("+"); rec[:, o + 2] = 10; o += 3
    rec[:, o:o + L] = qual; o += L
    rec[:, o] = 10
    return rec.reshape(-1)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}



# ------------------------------------------------------------------------------------------------
def reference_run(sample_host, shape, cfg, cores, steps, warmup, verbose=False):
    """Time the unmodified reference binary (-c then -d, -t cores) on FASTQ text of the sample blocks.
    Returns dict(value MB/s over compress+decompress wall, compress/decompress MB/s, entropy-stage MB/s)."""
    from oracle import pyoracle as po
    if not po.have_ref():
        return None
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="sqref_", dir=base)
    try:
        names = [os.path.join(tmp, "s_%d.fq" % (k + 1)) for k in range(shape.reads_per_unit)]
        files = [open(p, "wb") for p in names]
        for blk in sample_host:
            for k, f in enumerate(files):
                fastq_text(blk, shape, k).tofile(f)
        for f in files:
            f.close()
        nbytes = sum(os.path.getsize(p) for p in names)
        cargs = ["-c", "-t", str(cores), "-f"]
        dargs = ["-d", "-t", str(cores), "-f"]
        kw = {}
        if cfg["lossy"]:
            cargs += ["-l", "1.15"]
            kw["lossy_shim"] = True           # the reference overruns its quality buffer without it
        if cfg["ref_mode"]:                   # a random reference no read maps to: SeqArc -i, then -c ref.fa ...
            from tools.make_golden import random_reference
            open(os.path.join(tmp, "ref.fa"), "wb").write(random_reference(2000000, 5))
            rc, out = po.run_ref(["-i", "ref.fa"], tmp)
            assert rc == 0, out[-1500:]
            cargs += ["ref.fa"]
            dargs += ["ref.fa"]
        cargs += ["-1", names[0]] + (["-2", names[1]] if shape.pe else []) + ["-o", "s"]
        dargs += ["s.arc", "-o", "back"]
        tc, td, ent = [], [], []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            rc, out_c = po.run_ref(cargs, tmp, **kw)
            t1 = time.perf_counter()
            assert rc == 0, out_c[-1500:]
            rc, out_d = po.run_ref(dargs, tmp)
            t2 = time.perf_counter()
            assert rc == 0, out_d[-1500:]
            if it >= warmup:
                tc.append(t1 - t0)
                td.append(t2 - t1)
                ent.append((_stage_seconds(out_c, "encodeTime"), _stage_seconds(out_d, "decodeTime")))
            if verbose:
                log("reference step %d: compress %.2fs decompress %.2fs" % (it, t1 - t0, t2 - t1))
        arc = os.path.getsize(os.path.join(tmp, "s.arc"))
        c, d = float(np.mean(tc)), float(np.mean(td))
        res = dict(value=nbytes / 1e6 / (c + d), compress=nbytes / 1e6 / c, decompress=nbytes / 1e6 / d,
                   fastq_bytes=nbytes, arc_bytes=arc, ms_per_step=(c + d) * 1e3)
        es = [(a, b) for a, b in ent if a and b]
        if es:   # the binary's own DebugInfo timers are thread-seconds summed over workers
            a = float(np.mean([x[0] for x in es])) / cores
            b = float(np.mean([x[1] for x in es])) / cores
            res["entropy_stage_value"] = nbytes / 1e6 / (a + b)
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _stage_seconds(text, col):
    """Pick one column of the reference's 'readTime getdataTime md5Time ... writeTime' table."""
    lines = text.splitlines()
    for i, ln in enumerate(lines):
        cols = ln.split()
        if col in cols and i + 1 < len(lines):
            vals = lines[i + 1].split()
            try:
                return float(vals[cols.index(col)])
            except Exception:
                return None
    return None


# ------------------------------------------------------------------------------------------------
class Workload:
    """The blocks one rank holds in HBM, bound to the C ABI: device-resident step + host-buffer (e2e) step."""

    def __init__(self, torch, L_, _lib, h, stream, dev, shape, cfg, block_ids, units):
        self.torch, self.L_, self._lib, self.h, self.stream, self.shape, self.cfg = torch, L_, _lib, h, stream, shape, cfg
        self.blocks = [gen_block_device(torch, dev, shape, i, units, 7) for i in block_ids]
        torch.cuda.synchronize()
        n = self.n = len(self.blocks)
        self.fastq_bytes = n * shape.block_fastq_bytes(units)
        self.in_bytes = sum(int(b["ids"].numel() + b["seq"].numel() + b["qual"].numel() + 6 * b["nreads"]) for b in self.blocks)
        b0 = self.blocks[0] if n else None
        self.cap = int(0.55 * (b0["ids"].numel() + 2 * b0["seq"].numel())) + 4096 if n else 4096
        self.recbuf = [torch.empty(self.cap, dtype=torch.uint8, device=dev) for _ in range(n)]
        self.dec = [dict(ids=torch.empty_like(b["ids"]), seq=torch.empty_like(b["seq"]), qual=torch.empty_like(b["qual"]),
                         idlen=torch.empty_like(b["idlen"]), seqlen=torch.empty_like(b["seqlen"])) for b in self.blocks]
        self.ins = (_lib.SqBlockIn * max(n, 1))()
        self.outs = (_lib.SqBlockOut * max(n, 1))()
        self.recs = (C.c_void_p * max(n, 1))()
        self.lens = (C.c_uint64 * max(n, 1))()
        self.decs = (_lib.SqBlockDec * max(n, 1))()
        for i, b in enumerate(self.blocks):
            self.ins[i].nreads = b["nreads"]
            self.ins[i].ids, self.ins[i].seq, self.ins[i].qual = b["ids"].data_ptr(), b["seq"].data_ptr(), b["qual"].data_ptr()
            self.ins[i].idlen, self.ins[i].seqlen = b["idlen"].data_ptr(), b["seqlen"].data_ptr()
            self.ins[i].idbytes, self.ins[i].seqbytes = b["ids"].numel(), b["seq"].numel()
            self.outs[i].data, self.outs[i].cap = self.recbuf[i].data_ptr(), self.cap
            d = self.dec[i]
            self.decs[i].ids, self.decs[i].idcap = d["ids"].data_ptr(), d["ids"].numel()
            self.decs[i].seq, self.decs[i].qual, self.decs[i].seqcap = d["seq"].data_ptr(), d["qual"].data_ptr(), d["seq"].numel()
            self.decs[i].idlen, self.decs[i].seqlen, self.decs[i].readcap = d["idlen"].data_ptr(), d["seqlen"].data_ptr(), b["nreads"]
        self.stage_ms, self.launches, self.split_ms = {}, 0, {"enc": [], "dec": []}

    def _timings(self, prefix):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        k = self.L_.sq_last_timings(self.h, names, ms, 64)
        for j in range(k):
            self.stage_ms.setdefault(prefix + names[j].decode(), []).append(float(ms[j]))
        self.launches += int(self.L_.sq_last_launches(self.h))

    def step_device(self, record=False):
        torch, L_, h, n = self.torch, self.L_, self.h, self.n
        if n == 0:
            return
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if record:
            ev[0].record(self.stream)
        r = L_.sq_encode_blocks_dev(h, self.ins, n, self.outs)
        if r:
            raise SystemExit("encode failed: %d %s" % (r, L_.sq_last_error(h)))
        if record:
            ev[1].record(self.stream)
            self._timings("enc.")
        for i in range(n):
            self.recs[i], self.lens[i] = self.recbuf[i].data_ptr(), self.outs[i].len
        r = L_.sq_decode_blocks_dev(h, self.recs, self.lens, n, self.decs)
        if r:
            raise SystemExit("decode failed: %d %s" % (r, L_.sq_last_error(h)))
        if record:
            ev[2].record(self.stream)
            self._timings("dec.")
            torch.cuda.synchronize()
            self.split_ms["enc"].append(ev[0].elapsed_time(ev[1]))
            self.split_ms["dec"].append(ev[1].elapsed_time(ev[2]))

    def check_roundtrip(self):
        """Size-independent property at full size: decode(encode(x)) == x for every block of this rank (bases, IDs and
        lengths always; qualities when lossless, rblock(x) for the first block when lossy)."""
        torch = self.torch
        for i in range(self.n):
            b, d = self.blocks[i], self.dec[i]
            assert torch.equal(d["seq"], b["seq"]) and torch.equal(d["ids"], b["ids"]) and torch.equal(d["seqlen"], b["seqlen"]), "block %d" % i
            if not self.cfg["lossy"]:
                assert torch.equal(d["qual"], b["qual"]), "block %d qualities" % i
        if self.cfg["lossy"] and self.n:      # what comes back is the R-Block smoothing of what went in (EncapFqzComp::rblock on its own)
            q = self.blocks[0]["qual"].cpu().numpy().copy()
            r = self.L_.sq_rblock(self.h, q.ctypes.data, q.shape[0], 1.15)
            assert r == 0 and np.array_equal(q, self.dec[0]["qual"].cpu().numpy()), "decoded qualities are not rblock(input)"

    def release(self):
        """Drop the device-resident copy of the job (inputs, records, decoded output)."""
        self.blocks, self.recbuf, self.dec = [], [], []

    def record_lens(self):
        return [int(self.outs[i].len) for i in range(self.n)]


def make_ctx(L_, _lib, local, stream, cfg, usemodel):
    p = _lib.SqParams()
    p.device, p.md5, p.slevel, p.qlevel, p.id_rule, p.stream = local, 1, 3, 2, 0, stream.cuda_stream
    p.usemodel = usemodel
    p.lossy, p.lossy_r = cfg["lossy"], 1.15 if cfg["lossy"] else 0.0
    p.ref_mode = cfg["ref_mode"]
    h = C.c_void_p()
    r = L_.sq_create(C.byref(p), C.byref(h))
    if r:
        raise SystemExit("sq_create failed: %d %s" % (r, L_.sq_last_error(None)))
    return h


def train_snapshot(torch, L_, _lib, h, dev, shape, units):
    """SeqArcPreProcess sample: 5 head blocks of one record (pair) more than the coder's blocks hold, i.e. the first
    5 * (units + 1) units of the input.  Every rank rebuilds them (the data is seeded) and trains its own copy."""
    npairs = 5 * (units + 1)
    head = [gen_block_device(torch, dev, shape, i, units, 7) for i in range(6)]
    rpu = shape.reads_per_unit
    cat = {k: torch.cat([b[k] for b in head]) for k in ("ids", "seq", "qual", "idlen", "seqlen")}
    nr = rpu * npairs
    samp = {"ids": cat["ids"][:nr * shape.id_w], "seq": cat["seq"][:nr * shape.L], "qual": cat["qual"][:nr * shape.L],
            "idlen": cat["idlen"][:nr], "seqlen": cat["seqlen"][:nr], "nreads": nr}
    hs = to_host(torch, samp)
    del cat, samp, head
    sin = (_lib.SqBlockIn * 1)()
    sin[0].nreads = hs["nreads"]
    sin[0].ids, sin[0].seq, sin[0].qual = hs["ids"].ctypes.data, hs["seq"].ctypes.data, hs["qual"].ctypes.data
    sin[0].idlen, sin[0].seqlen = hs["idlen"].ctypes.data, hs["seqlen"].ctypes.data
    sin[0].idbytes, sin[0].seqbytes = hs["ids"].shape[0], hs["seq"].shape[0]
    blob = np.empty(int(L_.sq_model_size(h)), np.uint8)
    r = L_.sq_train_model(h, sin, 1, blob.ctypes.data, blob.shape[0])
    if r:
        raise SystemExit("train failed: %d %s" % (r, L_.sq_last_error(h)))
    load_snapshot(L_, h, blob)
    return blob, npairs


def load_snapshot(L_, h, blob):
    r = L_.sq_load_model(h, blob.ctypes.data, blob.shape[0])
    if r:
        raise SystemExit("load model failed: %d %s" % (r, L_.sq_last_error(h)))


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[a.config]
    shape = Shape(cfg["shape"])
    units = shape.units_per_block
    full_blocks = -(-cfg["units"] // units)
    job_blocks = a.blocks or full_blocks
    workload = ("%s; %d blocks x %d %s, sharded over %d GPU(s) (%s, BASELINE %s)"
                % (cfg["text"], job_blocks, units, "pairs" if shape.pe else "reads", world,
                   "the full job" if job_blocks == full_blocks else "subset of the job", cfg["baseline"]))
    cores = len(os.sched_getaffinity(0))
    job_fastq = job_blocks * shape.block_fastq_bytes(units)
    # the reference turns usemodel on above 14,069,722,771 input bytes (doCheckSetEncodeOpt @0x408104): decided on the WHOLE job
    usemodel = int(job_fastq * 0.0023947358 > 33693269.0)

    if a.impl == "reference":
        if rank != 0:
            return
        import torch
        nb = a.cpu_sample_blocks or max(16, 2 * cores)          # enough blocks for every host thread at any core count
        nb = min(nb, job_blocks)
        dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        sample = [to_host(torch, gen_block_device(torch, dev, shape, i, units, 7)) for i in range(nb)]
        r = reference_run(sample, shape, cfg, cores, a.steps, a.warmup, a.verbose)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/SeqArc-1.6 is not staged"}))
            return
        sample_txt = "%d blocks (%.0f MB FASTQ) of the workload, -t %d, compress then decompress, files in /dev/shm" % (nb, r["fastq_bytes"] / 1e6, cores)
        print(json.dumps({
            "impl": "reference", "metric": "FASTQ MB/s compress+decompress (entropy stage)", "value": round(r["value"], 2),
            "unit": "MB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(r["ms_per_step"], 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16/u32 integer", "data": "synthetic",
            "config": {"workload": workload, "sample": sample_txt},
            "compress_MBps": round(r["compress"], 2), "decompress_MBps": round(r["decompress"], 2),
            "entropy_stage_MBps": round(r.get("entropy_stage_value", 0), 2),
            "cpu_baseline": {"value": round(r["value"], 2), "unit": "MB/s", "cores": cores, "kind": "reference", "sample": sample_txt},
            "e2e": {"value": round(r["value"], 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import fastqueeze_b200 as fq
    from fastqueeze_b200 import _lib
    from fastqueeze_b200.shard import block_range, ordered_offsets
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator is built: send fd 1 to
        # stderr until that has happened, so that stdout carries the one JSON line only
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    L_ = fq.lib()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    snapshot = [None]

    def measure(block_ids, steps, warmup, label):
        """Build the rank's share, run warmup + timed steps; returns (workload, max-over-ranks ms, clocks)."""
        t0 = time.perf_counter()
        # the workload's tensors first: the context then takes its workspace out of what is left of the HBM
        w = Workload(torch, L_, _lib, None, stream, dev, shape, cfg, block_ids, units)
        h = w.h = make_ctx(L_, _lib, local, stream, cfg, usemodel)
        if usemodel:
            snapshot[0], npairs = train_snapshot(torch, L_, _lib, h, dev, shape, units)
            torch.cuda.empty_cache()
            log("rank %d: trained the %d-byte model snapshot on %d %s (pre-process, untimed)" % (rank, snapshot[0].shape[0], npairs, "pairs" if shape.pe else "reads"))
        log("rank %d: %s: %d blocks (%.1f MB FASTQ) ready in %.1fs" % (rank, label, w.n, w.fastq_bytes / 1e6, time.perf_counter() - t0))
        for _ in range(warmup):
            w.step_device()
        if warmup:
            w.check_roundtrip()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record(stream)
            for _ in range(steps):
                w.step_device(record=True)
            e1.record(stream)
            barrier()
        tmax = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return w, h, float(tmax.item()), clk

    # ---- the job, sharded: rank r holds the contiguous block range block_range(job_blocks, world, r)
    lo, hi = block_range(job_blocks, world, rank)
    w, h, ms, clk = measure(range(lo, hi), a.steps, a.warmup, "job share [%d, %d)" % (lo, hi))
    value = job_fastq * a.steps / 1e6 / (ms / 1e3)
    # host-side ordered gather: every rank learns the archive offset of each of its records (first record at 16, each
    # next at previous offset + previous length: the layout SeqArc-1.6 -t 1 writes, addBlockPEInfo @0x412f6d)
    lens, offs = ordered_offsets(w.record_lens(), job_blocks)
    arc_bytes = (offs[-1] + lens[-1] - 16) if lens else 0
    assert len(lens) == job_blocks and lens[lo:hi] == w.record_lens() and all(offs[i + 1] == offs[i] + lens[i] for i in range(len(lens) - 1))
    n_launch = w.launches
    mean_ms = {k: float(np.mean(v)) for k, v in w.stage_ms.items()}
    enc_ms = float(np.mean(w.split_ms["enc"])) if w.split_ms["enc"] else 0.0      # the chain kernels of the encoder overlap on side streams:
    dec_ms = float(np.mean(w.split_ms["dec"])) if w.split_ms["dec"] else 0.0      # stage_ms entries are per-stage spans, not a partition of enc_ms
    tms = torch.tensor([enc_ms, dec_ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms = float(tms[0].item()), float(tms[1].item())
    wave_blocks = int(L_.sq_last_wave_blocks(h))

    in_bytes_rank, n_rank = w.in_bytes, w.n
    cpu_sample = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu_sample = [to_host(torch, w.blocks[i]) for i in range(min(a.cpu_sample_blocks or max(16, 2 * cores), w.n))]

    # ---- e2e through the host-buffer C ABI (pinned host memory, copies inside the timed region).  What a caller with
    # host buffers sees: nothing of the job is resident on the device beforehand, so the device-resident copy above is
    # released and the context is created anew (its workspace is sized from the free HBM).
    e2e = None
    if not a.no_e2e and w.n:
        blocks = w.blocks
        per_block = int(blocks[0]["ids"].numel() + 2 * blocks[0]["seq"].numel() + 6 * blocks[0]["nreads"]) + w.cap
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        eblocks = int(max(1, min(w.n, (0.5 * avail / local_world) // per_block)))
        hb = [to_host(torch, b, pin=True) for b in blocks[:eblocks]]
        hrec = [torch.empty(w.cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(eblocks)]
        cap = w.cap
        L_.sq_destroy(h)
        del blocks
        w.release()
        torch.cuda.empty_cache()
        h = make_ctx(L_, _lib, local, stream, cfg, usemodel)
        if usemodel:
            load_snapshot(L_, h, snapshot[0])
        want0 = (hb[0]["seq"].tobytes(), hb[0]["ids"].tobytes())
        hins = (_lib.SqBlockIn * eblocks)()
        houts = (_lib.SqBlockOut * eblocks)()
        hrecs = (C.c_void_p * eblocks)()
        hlens = (C.c_uint64 * eblocks)()
        hdecs = (_lib.SqBlockDec * eblocks)()
        for i, b in enumerate(hb):
            hins[i].nreads = b["nreads"]
            hins[i].ids, hins[i].seq, hins[i].qual = b["ids"].ctypes.data, b["seq"].ctypes.data, b["qual"].ctypes.data
            hins[i].idlen, hins[i].seqlen = b["idlen"].ctypes.data, b["seqlen"].ctypes.data
            hins[i].idbytes, hins[i].seqbytes = b["ids"].shape[0], b["seq"].shape[0]
            houts[i].data, houts[i].cap = hrec[i].ctypes.data, cap
            hrecs[i] = hrec[i].ctypes.data
            hdecs[i].ids, hdecs[i].idcap = b["ids"].ctypes.data, b["ids"].shape[0]
            hdecs[i].seq, hdecs[i].qual, hdecs[i].seqcap = b["seq"].ctypes.data, b["qual"].ctypes.data, b["seq"].shape[0]
            hdecs[i].idlen, hdecs[i].seqlen, hdecs[i].readcap = b["idlen"].ctypes.data, b["seqlen"].ctypes.data, b["nreads"]

        e2e_ms = {}

        def last_copies(prefix, wall_s):
            names = (C.c_char_p * 64)()
            ms = (C.c_float * 64)()
            for j in range(L_.sq_last_timings(h, names, ms, 64)):
                if names[j].decode() in ("h2d", "d2h"):
                    e2e_ms[prefix + names[j].decode()] = round(float(ms[j]), 1)
            e2e_ms[prefix + "call"] = round(wall_s * 1e3, 1)

        def step_host():
            t = time.perf_counter()
            r = L_.sq_encode_blocks(h, hins, eblocks, houts)
            if r:
                raise SystemExit("host encode failed: %d %s" % (r, L_.sq_last_error(h)))
            last_copies("enc.", time.perf_counter() - t)
            for i in range(eblocks):
                hlens[i] = houts[i].len
            t = time.perf_counter()
            r = L_.sq_decode_blocks(h, hrecs, hlens, eblocks, hdecs)
            if r:
                raise SystemExit("host decode failed: %d %s" % (r, L_.sq_last_error(h)))
            last_copies("dec.", time.perf_counter() - t)

        step_host()
        assert (hb[0]["seq"].tobytes(), hb[0]["ids"].tobytes()) == want0
        barrier()
        ksteps = max(1, min(a.steps, 2))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        sums = torch.tensor([float(eblocks * shape.block_fastq_bytes(units)), float(in_bytes_rank * eblocks // n_rank),
                             float(sum(int(houts[i].len) for i in range(eblocks))), float(eblocks)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(sums)
        e_fastq, e_in, e_arc, e_blocks = [float(x) for x in sums.tolist()]
        e2e = {"value": round(e_fastq * ksteps / 1e6 / float(dt.item()), 2), "unit": "MB/s",
               "h2d_bytes_per_step": int(e_in + e_arc), "d2h_bytes_per_step": int(e_arc + e_in),
               "steps": ksteps, "blocks": int(e_blocks), "rank0_ms": e2e_ms}
        del hb, hrec

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference binary on a bounded sample
    cpu = None
    if cpu_sample:
        sample, nb = cpu_sample, len(cpu_sample)
        rr = reference_run(sample, shape, cfg, cores, 1, 0, a.verbose)
        if rr:
            cpu = {"value": round(rr["value"], 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                   "sample": "%d blocks (%.0f MB FASTQ) of the workload, SeqArc-1.6 -t %d, compress then decompress wall, files in /dev/shm"
                             % (nb, rr["fastq_bytes"] / 1e6, cores),
                   "compress_MBps": round(rr["compress"], 2), "decompress_MBps": round(rr["decompress"], 2),
                   "entropy_stage_MBps": round(rr.get("entropy_stage_value", 0), 2)}
        del sample, cpu_sample

    tot = torch.tensor([float(in_bytes_rank), float(n_launch)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    in_bytes_job, n_launch_job = int(tot[0].item()), int(tot[1].item())

    # ---- weak-scaling figure (N > 1): every rank codes a whole job of its own
    replica = None
    if world > 1 and not a.no_replica:
        L_.sq_destroy(h)
        w.release()
        del w
        torch.cuda.empty_cache()
        w2, h, ms2, _ = measure(range(rank * job_blocks, (rank + 1) * job_blocks), 2, 1, "replica (a whole job per rank)")
        replica = round(world * job_fastq * 2 / 1e6 / (ms2 / 1e3), 2)
        del w2

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the stage with the largest device time
        top = max(mean_ms, key=mean_ms.get)
        nb_r = max(n_rank, 1)
        per_block_in, per_block_arc = in_bytes_rank / nb_r, (arc_bytes / max(job_blocks, 1))
        alg_bytes = ALG_BYTES.get(top, ALG_BYTES_DEFAULT)(per_block_in * nb_r, per_block_arc * nb_r, nb_r, shape)
        if top == "dec.decode" and usemodel:
            alg_bytes += nb_r * 27_203_584             # the snapshot is read once per block to restore the models
        achieved = alg_bytes / 1e9 / (mean_ms[top] / 1e3)
        traffic, tsrc = measured_traffic(top, a.config, job_blocks, world, usemodel)
        out = {
            "metric": "FASTQ MB/s compress+decompress (entropy stage)", "value": round(value, 2), "unit": "MB/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 2),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16/u32 integer",
            "data": "synthetic",
            "config": {"workload": workload, "name": a.config, "fastq_MB": round(job_fastq / 1e6, 1), "blocks": job_blocks,
                       "blocks_per_gpu": n_rank, "ratio": round(job_fastq / arc_bytes, 3) if arc_bytes else None, "usemodel": usemodel,
                       "l2": "inputs larger than L2 (>= 52 MB per block stream set; no flush)", "wave_blocks": wave_blocks,
                       "arc_bytes": int(arc_bytes)},
            "compress_MBps": round(job_fastq / 1e6 / (enc_ms / 1e3), 2) if enc_ms else None,
            "decompress_MBps": round(job_fastq / 1e6 / (dec_ms / 1e3), 2) if dec_ms else None,
            "replica_MBps": replica,
            "stage_ms": {k: round(v, 3) for k, v in sorted(mean_ms.items(), key=lambda kv: -kv[1])},
            "gpu_launches": int(n_launch_job),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": tsrc,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "rank 0's launch; the chains of a block are serial, so the kernel is instruction-latency bound (profiles/)"},
            "e2e": e2e, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    L_.sq_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def measured_traffic(top, config, job_blocks, world, usemodel):
    """DRAM bytes of the dominant kernel per launch (ncu: dram__bytes_read.sum + dram__bytes_write.sum), only for the
    configuration a capture exists for (profiles/traffic.json, written by tools/ncu_traffic.py)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, None
    key = "%s:%s:%d:%d:%d" % (top, config, job_blocks, world, usemodel)
    if key in t:
        return int(t[key]["bytes"]), t[key]["source"]
    return None, None


def to_host(torch, b, pin=False):
    out = {"nreads": b["nreads"]}
    for k in ("ids", "seq", "qual", "idlen", "seqlen"):
        t = b[k]
        if t.is_cuda:
            if pin:
                ht = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                ht.copy_(t)
            else:
                ht = t.cpu()
        else:
            ht = t
        out[k] = ht.numpy()
        out["_keep_" + k] = ht
    out["idlen"] = out["idlen"].view(np.uint16)
    return out


# Algorithmic HBM bytes per launch of each stage (what the stage must read + write once); args: input bytes, record
# bytes, blocks of the launch, shape
ALG_BYTES_DEFAULT = lambda inb, arc, nb, sh: inb + arc                                    # noqa: E731
ALG_BYTES = {
    # decode: read the record, write ids+seq+qual, reset the per-block model tables (35.3 MB per block)
    "dec.decode": lambda inb, arc, nb, sh: arc + inb + nb * 35_270_656,
    # range-coder pass of one stream: read one 8-byte (bases: 4-byte) triple per coded symbol, write the coded bytes
    "enc.qual.rc": lambda inb, arc, nb, sh: 8 * (inb * sh.L // (2 * sh.L + sh.id_w + 6)) + arc * 3 // 4,
    "enc.seq.rc": lambda inb, arc, nb, sh: 4 * (inb * sh.L // (2 * sh.L + sh.id_w + 6)) + arc // 4,
}

if __name__ == "__main__":
    main()
