#!/usr/bin/env python
"""bench.py -- entropy-stage compress + decompress throughput of the B200 SeqArc path.

A "step" is one pass of the hot path over one batch of synthetic FASTQ blocks of the shape
BASELINE.json `configs[1]` names (Illumina-style 150 bp PE, no reference, lossless): every block
is compressed (doFqzEncode seam) and the records are decompressed again (doDecode seam).

  value   whole-job MB/s of input FASTQ, inputs already resident in HBM (device pointers through
          sq_encode_blocks_dev / sq_decode_blocks_dev), CUDA events on the launching stream
  e2e     the same through the host-buffer C-ABI calls (sq_encode_blocks / sq_decode_blocks):
          pinned host buffers in, host records out, host<->device copies inside the timed region
  --impl reference   the UNMODIFIED reference binary (oracle/_ref/SeqArc-1.6) on the host cores,
          -t <cores>, on a bounded sample of the same workload

Multi-GPU: blocks are independent units (models reset per block), so ranks take disjoint block
sets with no data-path collective ("weak": per-GPU work is fixed); torch.distributed is used only
for the barrier and the max-over-ranks time.
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_PAIRS = 20_000_000           # BASELINE.json configs[1]: 20 M read pairs
BLOCK_BYTES = 50 << 20            # SeqArc default BlockSize = 50 MiB of FASTQ per block (25 MiB per mate file)
L = 150
ID_W = len(b"A00123:45:HXXXXDSXX:1:") + 4 + 1 + 5 + 1 + 5 + len(b" 1:N:0:ACGTACGT")   # fixed-width Illumina-style ID
REC_BYTES = (1 + ID_W + 1) + (L + 1) + 2 + (L + 1)     # FASTQ text bytes per read
PAIRS_PER_BLOCK = BLOCK_BYTES // (2 * REC_BYTES)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=0, help="blocks per GPU (0 = the full configs[1] job: 20 M pairs)")
    ap.add_argument("--cpu-sample-blocks", type=int, default=0, help="blocks of the CPU baseline sample (0 = one per host core, max 16)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    return ap.parse_args()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic blocks, generated on the device (SURVEY 8d input 2: Illumina-style IDs, iid ACGT,
# qualities clipped N(36,5)+33); fixed-width fields so that a block is a handful of matrix ops
def gen_block_device(torch, dev, block_index, pairs, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 1_000_003 + block_index)
    n = 2 * pairs
    first = block_index * pairs
    idx = torch.arange(first, first + pairs, device=dev, dtype=torch.int64)
    tile = 1101 + (idx // 4096) % 8000
    x = torch.randint(1000, 32000, (pairs,), device=dev, generator=g)
    y = torch.randint(1000, 32000, (pairs,), device=dev, generator=g)

    def digits(v, w):
        p = torch.tensor([10 ** k for k in range(w - 1, -1, -1)], device=dev, dtype=torch.int64)
        return ((v[:, None] // p[None, :]) % 10 + 48).to(torch.uint8)

    def const(s):
        return torch.tensor(list(s), device=dev, dtype=torch.uint8)[None, :].expand(pairs, len(s))

    def ids(mate):
        return torch.cat([const(b"A00123:45:HXXXXDSXX:1:"), digits(tile, 4), const(b":"), digits(x, 5), const(b":"),
                          digits(y, 5), const(b" %d:N:0:ACGTACGT" % mate)], dim=1)

    m1, m2 = ids(1), ids(2)
    assert m1.shape[1] == ID_W
    idm = torch.stack([m1, m2], dim=1).reshape(n, ID_W).contiguous()
    acgt = torch.tensor(list(b"ACGT"), device=dev, dtype=torch.uint8)
    seq = acgt[torch.randint(0, 4, (n, L), device=dev, generator=g)].contiguous()
    q = torch.empty((n, L), device=dev, dtype=torch.float32).normal_(36.0, 5.0, generator=g)
    qual = (q.round().clamp_(2, 40) + 33).to(torch.uint8).contiguous()
    return dict(ids=idm.reshape(-1), seq=seq.reshape(-1), qual=qual.reshape(-1),
                idlen=torch.full((n,), ID_W, device=dev, dtype=torch.int16),
                seqlen=torch.full((n,), L, device=dev, dtype=torch.int32), nreads=n)


def fastq_text(blk_host, mate):
    """Render one mate of a (host, numpy) block as FASTQ text, vectorised (fixed-width records)."""
    n = blk_host["nreads"] // 2
    ids = blk_host["ids"].reshape(2 * n, ID_W)[mate::2]
    seq = blk_host["seq"].reshape(2 * n, L)[mate::2]
    qual = blk_host["qual"].reshape(2 * n, L)[mate::2]
    rec = np.empty((n, REC_BYTES), dtype=np.uint8)
    o = 0
    rec[:, o] = ord("@"); o += 1
    rec[:, o:o + ID_W] = ids; o += ID_W
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
def reference_run(sample_host, cores, steps, warmup, verbose=False):
    """Time the unmodified reference binary (-c then -d, -t cores) on FASTQ text of the sample blocks.
    Returns dict(value MB/s over compress+decompress wall, compress/decompress MB/s, entropy-stage MB/s)."""
    from oracle import pyoracle as po
    if not po.have_ref():
        return None
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="sqref_", dir=base)
    try:
        f1, f2 = os.path.join(tmp, "s_1.fq"), os.path.join(tmp, "s_2.fq")
        with open(f1, "wb") as a, open(f2, "wb") as b:
            for blk in sample_host:
                fastq_text(blk, 0).tofile(a)
                fastq_text(blk, 1).tofile(b)
        nbytes = os.path.getsize(f1) + os.path.getsize(f2)
        tc, td, ent = [], [], []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            rc, out_c = po.run_ref(["-c", "-t", str(cores), "-f", "-1", f1, "-2", f2, "-o", "s"], tmp)
            t1 = time.perf_counter()
            assert rc == 0, out_c[-1500:]
            rc, out_d = po.run_ref(["-d", "-t", str(cores), "-f", "s.arc", "-o", "back"], tmp)
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
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    full_blocks = -(-FULL_PAIRS // PAIRS_PER_BLOCK)
    nblocks = a.blocks or full_blocks
    workload = ("synthetic Illumina 150 bp PE, no reference, lossless; %d blocks x %d pairs per GPU (%s)"
                % (nblocks, PAIRS_PER_BLOCK, "the full 20M-pair job" if nblocks == full_blocks else "subset of the 20M-pair job"))
    cores = len(os.sched_getaffinity(0))

    if a.impl == "reference":
        if rank != 0:
            return
        import torch
        nb = a.cpu_sample_blocks or min(cores, 16)
        dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        sample = [to_host(torch, gen_block_device(torch, dev, i, PAIRS_PER_BLOCK, 7)) for i in range(nb)]
        r = reference_run(sample, cores, a.steps, a.warmup, a.verbose)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/SeqArc-1.6 is not staged"}))
            return
        sample_txt = "%d blocks (%.0f MB FASTQ) of the workload, -t %d, compress then decompress, files in /dev/shm" % (nb, r["fastq_bytes"] / 1e6, cores)
        print(json.dumps({
            "impl": "reference", "metric": "FASTQ MB/s compress+decompress (entropy stage)", "value": round(r["value"], 2),
            "unit": "MB/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(r["ms_per_step"], 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16/u32 integer", "data": "synthetic",
            "config": {"workload": workload, "sample": sample_txt},
            "compress_MBps": round(r["compress"], 2), "decompress_MBps": round(r["decompress"], 2),
            "entropy_stage_MBps": round(r.get("entropy_stage_value", 0), 2),
            "cpu_baseline": {"value": round(r["value"], 2), "unit": "MB/s", "cores": cores, "kind": "reference", "sample": sample_txt},
            "e2e": {"value": round(r["value"], 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import fastqueeze_b200 as fq
    from fastqueeze_b200 import _lib
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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

    # ---- workload resident in HBM
    t0 = time.perf_counter()
    # weak scaling: rank r owns blocks [r * nblocks, (r + 1) * nblocks) of an N-times larger job
    from fastqueeze_b200.shard import block_range
    lo, hi = block_range(nblocks * world, world, rank)
    blocks = [gen_block_device(torch, dev, i, PAIRS_PER_BLOCK, 7) for i in range(lo, hi)]
    torch.cuda.synchronize()
    fastq_bytes = nblocks * 2 * PAIRS_PER_BLOCK * REC_BYTES
    in_bytes = sum(int(b["ids"].numel() + b["seq"].numel() + b["qual"].numel() + 6 * b["nreads"]) for b in blocks)
    log("rank %d: generated %d blocks (%.1f MB FASTQ) in %.1fs" % (rank, nblocks, fastq_bytes / 1e6, time.perf_counter() - t0))
    cap = int(0.55 * (blocks[0]["ids"].numel() + 2 * blocks[0]["seq"].numel())) + 4096
    recbuf = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(nblocks)]
    dec = [dict(ids=torch.empty_like(b["ids"]), seq=torch.empty_like(b["seq"]), qual=torch.empty_like(b["qual"]),
                idlen=torch.empty_like(b["idlen"]), seqlen=torch.empty_like(b["seqlen"])) for b in blocks]

    stream = torch.cuda.current_stream()
    # the reference turns usemodel on above 14,069,722,771 input bytes (doCheckSetEncodeOpt @0x408104): the full
    # configs[1] job is above that, so it runs from a snapshot trained on the five head sample blocks
    usemodel = int(fastq_bytes * 0.0023947358 > 33693269.0)
    p = _lib.SqParams()
    p.device, p.md5, p.slevel, p.qlevel, p.id_rule, p.stream = local, 1, 3, 2, 0, stream.cuda_stream
    p.usemodel = usemodel
    h = C.c_void_p()
    r = L_.sq_create(C.byref(p), C.byref(h))
    if r:
        raise SystemExit("sq_create failed: %d %s" % (r, L_.sq_last_error(None)))

    if usemodel:
        # SeqArcPreProcess sample: 5 blocks of ceil(BlockSize/2 / record) pairs = the first 5 * (PAIRS_PER_BLOCK + 1) pairs
        t0 = time.perf_counter()
        npairs = 5 * (PAIRS_PER_BLOCK + 1)
        cat = {k: torch.cat([blocks[i][k] for i in range(6)]) for k in ("ids", "seq", "qual", "idlen", "seqlen")}
        samp = {"ids": cat["ids"][:2 * npairs * ID_W], "seq": cat["seq"][:2 * npairs * L], "qual": cat["qual"][:2 * npairs * L],
                "idlen": cat["idlen"][:2 * npairs], "seqlen": cat["seqlen"][:2 * npairs], "nreads": 2 * npairs}
        hs = to_host(torch, samp)
        del cat, samp
        sin = (_lib.SqBlockIn * 1)()
        sin[0].nreads = hs["nreads"]
        sin[0].ids, sin[0].seq, sin[0].qual = hs["ids"].ctypes.data, hs["seq"].ctypes.data, hs["qual"].ctypes.data
        sin[0].idlen, sin[0].seqlen = hs["idlen"].ctypes.data, hs["seqlen"].ctypes.data
        sin[0].idbytes, sin[0].seqbytes = hs["ids"].shape[0], hs["seq"].shape[0]
        blob = np.empty(int(L_.sq_model_size(h)), np.uint8)
        r = L_.sq_train_model(h, sin, 1, blob.ctypes.data, blob.shape[0])
        if r:
            raise SystemExit("train failed: %d %s" % (r, L_.sq_last_error(h)))
        r = L_.sq_load_model(h, blob.ctypes.data, blob.shape[0])
        if r:
            raise SystemExit("load model failed: %d %s" % (r, L_.sq_last_error(h)))
        log("rank %d: trained the %d-byte model snapshot on %d pairs in %.1fs (pre-process, untimed)"
            % (rank, blob.shape[0], npairs, time.perf_counter() - t0))
        del hs

    ins = (_lib.SqBlockIn * nblocks)()
    outs = (_lib.SqBlockOut * nblocks)()
    for i, b in enumerate(blocks):
        ins[i].nreads = b["nreads"]
        ins[i].ids, ins[i].seq, ins[i].qual = b["ids"].data_ptr(), b["seq"].data_ptr(), b["qual"].data_ptr()
        ins[i].idlen, ins[i].seqlen = b["idlen"].data_ptr(), b["seqlen"].data_ptr()
        ins[i].idbytes, ins[i].seqbytes = b["ids"].numel(), b["seq"].numel()
        outs[i].data, outs[i].cap = recbuf[i].data_ptr(), cap
    recs = (C.c_void_p * nblocks)()
    lens = (C.c_uint64 * nblocks)()
    decs = (_lib.SqBlockDec * nblocks)()
    for i, d in enumerate(dec):
        decs[i].ids, decs[i].idcap = d["ids"].data_ptr(), d["ids"].numel()
        decs[i].seq, decs[i].qual, decs[i].seqcap = d["seq"].data_ptr(), d["qual"].data_ptr(), d["seq"].numel()
        decs[i].idlen, decs[i].seqlen, decs[i].readcap = d["idlen"].data_ptr(), d["seqlen"].data_ptr(), blocks[i]["nreads"]

    stage_ms = {}
    launches = [0]

    def timings(prefix):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        k = L_.sq_last_timings(h, names, ms, 64)
        for j in range(k):
            stage_ms.setdefault(prefix + names[j].decode(), []).append(float(ms[j]))
        launches[0] += int(L_.sq_last_launches(h))

    split_ms = {"enc": [], "dec": []}

    def step_device(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if record:
            ev[0].record(stream)
        r = L_.sq_encode_blocks_dev(h, ins, nblocks, outs)
        if r:
            raise SystemExit("encode failed: %d %s" % (r, L_.sq_last_error(h)))
        if record:
            ev[1].record(stream)
            timings("enc.")
        for i in range(nblocks):
            recs[i], lens[i] = recbuf[i].data_ptr(), outs[i].len
        r = L_.sq_decode_blocks_dev(h, recs, lens, nblocks, decs)
        if r:
            raise SystemExit("decode failed: %d %s" % (r, L_.sq_last_error(h)))
        if record:
            ev[2].record(stream)
            timings("dec.")
            torch.cuda.synchronize()
            split_ms["enc"].append(ev[0].elapsed_time(ev[1]))
            split_ms["dec"].append(ev[1].elapsed_time(ev[2]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_device()
    # round trip is exact (size-independent property at the full size)
    for i in (0, nblocks - 1):
        assert torch.equal(dec[i]["seq"], blocks[i]["seq"]) and torch.equal(dec[i]["qual"], blocks[i]["qual"])
        assert torch.equal(dec[i]["ids"], blocks[i]["ids"]) and torch.equal(dec[i]["seqlen"], blocks[i]["seqlen"])
    arc_bytes = sum(int(outs[i].len) for i in range(nblocks))

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(a.steps):
            step_device(record=True)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * fastq_bytes * a.steps / 1e6 / (ms / 1e3)
    n_launch = launches[0]
    mean_ms = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    enc_ms = float(np.mean(split_ms["enc"]))      # the chain kernels of the encoder overlap on side streams:
    dec_ms = float(np.mean(split_ms["dec"]))      # stage_ms entries are per-stage spans, not a partition of enc_ms

    # ---- e2e through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        # host memory budget: every rank of the box pins its own buffers (inputs + record buffers; the decoder
        # writes back into the input buffers, which then hold the same bytes again)
        per_block = int(blocks[0]["ids"].numel() + 2 * blocks[0]["seq"].numel() + 6 * blocks[0]["nreads"]) + cap
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        except Exception:
            avail = 64 << 30
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        eblocks = int(max(1, min(nblocks, (0.5 * avail / local_world) // per_block)))
        hb = [to_host(torch, b, pin=True) for b in blocks[:eblocks]]
        hrec = [torch.empty(cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(eblocks)]
        want0 = (hb[0]["qual"].tobytes(), hb[0]["seq"].tobytes(), hb[0]["ids"].tobytes())
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

        def step_host():
            r = L_.sq_encode_blocks(h, hins, eblocks, houts)
            if r:
                raise SystemExit("host encode failed: %d %s" % (r, L_.sq_last_error(h)))
            for i in range(eblocks):
                hlens[i] = houts[i].len
            r = L_.sq_decode_blocks(h, hrecs, hlens, eblocks, hdecs)
            if r:
                raise SystemExit("host decode failed: %d %s" % (r, L_.sq_last_error(h)))

        step_host()
        assert (hb[0]["qual"].tobytes(), hb[0]["seq"].tobytes(), hb[0]["ids"].tobytes()) == want0
        barrier()
        ksteps = max(1, min(a.steps, 2))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e_fastq = eblocks * 2 * PAIRS_PER_BLOCK * REC_BYTES
        e_in = in_bytes * eblocks // nblocks
        e_arc = sum(int(houts[i].len) for i in range(eblocks))
        e2e = {"value": round(world * e_fastq * ksteps / 1e6 / float(dt.item()), 2), "unit": "MB/s",
               "h2d_bytes_per_step": int(e_in + e_arc), "d2h_bytes_per_step": int(e_arc + e_in),
               "steps": ksteps, "blocks_per_gpu": eblocks}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference binary on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        nb = a.cpu_sample_blocks or min(cores, 16, nblocks)
        sample = [to_host(torch, blocks[i]) for i in range(nb)]
        rr = reference_run(sample, cores, 1, 0, a.verbose)
        if rr:
            cpu = {"value": round(rr["value"], 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                   "sample": "%d blocks (%.0f MB FASTQ) of the workload, SeqArc-1.6 -t %d, compress then decompress wall, files in /dev/shm"
                             % (nb, rr["fastq_bytes"] / 1e6, cores),
                   "compress_MBps": round(rr["compress"], 2), "decompress_MBps": round(rr["decompress"], 2),
                   "entropy_stage_MBps": round(rr.get("entropy_stage_value", 0), 2)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the stage with the largest device time
        top = max(mean_ms, key=mean_ms.get)
        # algorithmic HBM bytes of the whole path per input FASTQ byte (DESIGN.md / SURVEY 8d):
        #   compress 1.9 B/B (read id+seq+qual once, write the record, per-block model reset), decompress the same
        alg_bytes = ALG_BYTES.get(top, ALG_BYTES_DEFAULT)(fastq_bytes, in_bytes, arc_bytes, nblocks)
        if top == "dec.decode" and usemodel:
            alg_bytes += nblocks * 27_203_584          # the snapshot is read once per block to restore the models
        achieved = alg_bytes / 1e9 / (mean_ms[top] / 1e3)
        # DRAM bytes of that kernel per launch, from the ncu capture of this command on the full job
        # (profiles/r1f_traffic_k_dec_main.csv: dram__bytes_read.sum + dram__bytes_write.sum); only quoted for
        # the configuration it was measured on
        traffic = None
        if top == "dec.decode" and nblocks == 274 and usemodel:
            traffic = 1_074_329_448_960 + 549_005_104_896
        out = {
            "metric": "FASTQ MB/s compress+decompress (entropy stage)", "value": round(value, 2), "unit": "MB/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16/u32 integer",
            "data": "synthetic",
            "config": {"workload": workload, "fastq_MB_per_gpu": round(fastq_bytes / 1e6, 1), "blocks_per_gpu": nblocks,
                       "ratio": round(fastq_bytes / arc_bytes, 3), "usemodel": usemodel, "l2": "inputs larger than L2 (>= 52 MB per block stream set; no flush)",
                       "wave_blocks": int(L_.sq_last_wave_blocks(h))},
            "compress_MBps": round(world * fastq_bytes / 1e6 / (enc_ms / 1e3), 2) if enc_ms else None,
            "decompress_MBps": round(world * fastq_bytes / 1e6 / (dec_ms / 1e3), 2) if dec_ms else None,
            "stage_ms": {k: round(v, 3) for k, v in sorted(mean_ms.items(), key=lambda kv: -kv[1])},
            "gpu_launches": int(n_launch),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic,
                         "traffic_source": "profiles/r1f_traffic_k_dec_main.csv (ncu, k_dec_main, one launch)" if traffic else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": int(alg_bytes)},
            "e2e": e2e, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    L_.sq_destroy(h)
    if world > 1:
        dist.destroy_process_group()


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


# Algorithmic HBM bytes per launch of each stage (what the stage must read + write once):
def _sym(fq, inb, arc, nb):
    return inb


ALG_BYTES_DEFAULT = lambda fq, inb, arc, nb: inb + arc                                    # noqa: E731
ALG_BYTES = {
    # decode: read the record, write ids+seq+qual, reset the per-block model tables (35.3 MB per block)
    "dec.decode": lambda fq, inb, arc, nb: arc + inb + nb * 35_270_656,
    # range-coder pass of one stream: read one 8-byte triple per coded symbol, write the coded bytes
    "enc.qual.rc": lambda fq, inb, arc, nb: 8 * (inb * 150 // 363) + arc * 3 // 4,
    "enc.seq.rc": lambda fq, inb, arc, nb: 8 * (inb * 150 // 363) + arc // 4,
}

if __name__ == "__main__":
    main()
