"""Whole-file parity at scale (not part of pytest: minutes of reference CPU time).

Generates the bench's synthetic FASTQ of one BASELINE config for N default-size blocks (default: the whole config),
writes it to /dev/shm, and checks

  --mode t1      seqarc_b200 -c == SeqArc-1.6 -c -t 1 (byte-identical .arc), seqarc_b200 -d of the reference's archive and
                 of a -t <cores> archive (blocks written out of order) == the input
  --mode records (full-size jobs: -t 1 would take an hour) SeqArc-1.6 -c -t <cores> against seqarc_b200 -c: every block
                 record of the reference's archive is found through the block index (source offset) and compared byte for
                 byte with ours, the trailer parameters and the model snapshot (usemodel = 1 above 14.07 GB) too; then
                 seqarc_b200 -d of the REFERENCE's archive == the input, and SeqArc-1.6 -d of OUR archive == the input

Prints one JSON line.  Run on the GPU box, e.g.  python tools/scale_check.py --config pe150 --mode records
"""
import argparse
import hashlib
import json
import mmap
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import arcwalk, pyoracle as po  # noqa: E402

CLI = os.path.join(ROOT, "fastqueeze_b200", "seqarc_b200")


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for c in iter(lambda: f.read(1 << 24), b""):
            h.update(c)
    return h.hexdigest()


def run(args, cwd):
    t0 = time.perf_counter()
    r = subprocess.run(args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, (args, r.stdout.decode("latin1")[-2000:])
    return time.perf_counter() - t0


def trailer_of(path):
    """(mapped archive, parsed trailer) without walking the gigabytes of block records."""
    f = open(path, "rb")
    m = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    assert m[:7] == arcwalk.MAGIC
    tag, p = arcwalk.read_vint(m, 7)
    size, p = arcwalk.read_vint(m, p)
    assert tag == 2
    tpos = p + size
    tag, q = arcwalk.read_vint(m, tpos)
    tsize, q = arcwalk.read_vint(m, q)
    assert tag == 3
    tr = {}
    for t, a, b in arcwalk.walk_children(m, q, q + tsize):
        if t == 1:
            tr["params"] = {pid: bytes(m[x:y]) for pid, x, y in arcwalk.walk_children(m, a, b)}
        elif t == 6:
            tr["model_md5"] = hashlib.md5(m[a:b]).hexdigest()
            tr["model_bytes"] = b - a
        elif t == 7:
            tr["index"] = bytes(m[a:b])
    return m, tr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="pe150", choices=sorted(bench.CONFIGS))
    ap.add_argument("--blocks", type=int, default=0, help="0 = the whole config")
    ap.add_argument("--mode", default="t1", choices=["t1", "records"])
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    import torch
    cfg = bench.CONFIGS[a.config]
    shape = bench.Shape(cfg["shape"])
    units = shape.units_per_block
    nblocks = a.blocks or -(-cfg["units"] // units)
    dev = torch.device("cuda:0")
    # the files live in /dev/shm: refuse to start rather than drive the box out of memory
    need = nblocks * shape.block_fastq_bytes(units) * 3.2
    avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    shm = shutil.disk_usage("/dev/shm").free if os.path.isdir("/dev/shm") else 0
    if avail < need + (24 << 30) or shm < need:
        print(json.dumps({"config": a.config, "skipped": "not enough host memory: need %.0f GB, MemAvailable %.0f GB, /dev/shm free %.0f GB"
                          % (need / 1e9, avail / 1e9, shm / 1e9)}))
        return
    tmp = tempfile.mkdtemp(prefix="sqscale_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        names = [os.path.join(tmp, "s_%d.fq" % (k + 1)) for k in range(shape.reads_per_unit)]
        files = [open(p, "wb") for p in names]
        for i in range(nblocks):
            blk = bench.to_host(torch, bench.gen_block_device(torch, dev, shape, i, units, 7))
            for k, f in enumerate(files):
                bench.fastq_text(blk, shape, k).tofile(f)
        for f in files:
            f.close()
        torch.cuda.empty_cache()
        nbytes = sum(os.path.getsize(p) for p in names)
        res = {"config": a.config, "mode": a.mode, "blocks": nblocks, "fastq_bytes": nbytes, "fastq_md5": [md5(p) for p in names],
               "threads": a.threads}
        lossy = ["-l", "1.15"] if cfg["lossy"] else []
        inputs = ["-1", names[0]] + (["-2", names[1]] if shape.pe else [])
        outs = (lambda pre: [os.path.join(tmp, pre + "_%d.fastq" % k) for k in (1, 2)]) if shape.pe else (lambda pre: [os.path.join(tmp, pre + ".fastq")])
        if a.mode == "records":
            t0 = time.perf_counter()
            rc, log = po.run_ref(["-c", "-t", str(a.threads), "-f"] + lossy + inputs + ["-o", "ref"], tmp, lossy_shim=bool(lossy))
            assert rc == 0, log[-1500:]
            res["ref_c_s"] = round(time.perf_counter() - t0, 1)
            res["ref_usemodel"] = int("usemodel=1" in log)
            res["ours_c_s"] = round(run([CLI, "-c", "-f"] + lossy + inputs + ["-o", "ours"], tmp), 1)
            mr, tr = trailer_of(os.path.join(tmp, "ref.arc"))
            mo, to = trailer_of(os.path.join(tmp, "ours.arc"))
            ir = arcwalk.block_index({"trailer": tr}, shape.pe)
            io = arcwalk.block_index({"trailer": to}, shape.pe)
            by_src = {b["offs"][1]: b for b in io}
            same = 0
            for b in ir:
                o = by_src.get(b["offs"][1])
                if o and o["len"] == b["len"] and o["src1"] == b["src1"] and o["src2"] == b["src2"] and \
                        mr[b["offs"][0]:b["offs"][0] + b["len"]] == mo[o["offs"][0]:o["offs"][0] + o["len"]]:
                    same += 1
            res.update(records_identical="%d/%d" % (same, len(ir)), blocks_ours=len(io),
                       arc_bytes_ref=os.path.getsize(os.path.join(tmp, "ref.arc")), arc_bytes_ours=os.path.getsize(os.path.join(tmp, "ours.arc")),
                       params_identical=tr["params"] == to["params"],
                       model_md5_ref=tr.get("model_md5"), model_md5_ours=to.get("model_md5"), model_bytes=tr.get("model_bytes"),
                       ours_in_block_order=[b["offs"][0] for b in io] == sorted(b["offs"][0] for b in io) and
                       [b["offs"][1] for b in io] == sorted(b["offs"][1] for b in io))
            mr.close(); mo.close()
            res["ours_d_of_ref_s"] = round(run([CLI, "-d", "-f", "ref.arc", "-o", "back"], tmp), 1)
            if lossy:
                rc, log = po.run_ref(["-d", "-t", str(a.threads), "-f", "ref.arc", "-o", "rback"], tmp)
                assert rc == 0, log[-1500:]
                res["decode_equals_reference_decode"] = [md5(p) for p in outs("back")] == [md5(p) for p in outs("rback")]
            else:
                res["decode_of_ref_archive_ok"] = [md5(p) for p in outs("back")] == res["fastq_md5"]
                for p in outs("back"):
                    os.remove(p)
                t0 = time.perf_counter()
                rc, log = po.run_ref(["-d", "-t", str(a.threads), "-f", "ours.arc", "-o", "rback"], tmp)
                assert rc == 0, log[-1500:]
                res["ref_d_of_ours_s"] = round(time.perf_counter() - t0, 1)
                res["reference_decodes_our_archive_ok"] = [md5(p) for p in outs("rback")] == res["fastq_md5"]
            res["MBps"] = {k[:-2]: round(nbytes / 1e6 / res[k], 1) for k in ("ref_c_s", "ours_c_s", "ours_d_of_ref_s", "ref_d_of_ours_s") if k in res}
            print(json.dumps(res))
            return
        t0 = time.perf_counter()
        rc, log = po.run_ref(["-c", "-t", "1", "-f"] + lossy + inputs + ["-o", "ref"], tmp, lossy_shim=bool(lossy))
        assert rc == 0, log[-1500:]
        res["ref_c_t1_s"] = round(time.perf_counter() - t0, 1)
        res["ours_c_s"] = round(run([CLI, "-c", "-f"] + lossy + inputs + ["-o", "ours"], tmp), 1)
        res["arc_bytes"] = os.path.getsize(os.path.join(tmp, "ref.arc"))
        res["arc_md5_ref"], res["arc_md5_ours"] = md5(os.path.join(tmp, "ref.arc")), md5(os.path.join(tmp, "ours.arc"))
        res["archive_identical"] = res["arc_md5_ref"] == res["arc_md5_ours"]
        res["ours_d_s"] = round(run([CLI, "-d", "-f", "ref.arc", "-o", "back"], tmp), 1)
        if lossy:       # lossy: the decoded qualities are the smoothed ones -- compare with the reference's own decode
            rc, log = po.run_ref(["-d", "-t", str(a.threads), "-f", "ref.arc", "-o", "rback"], tmp)
            assert rc == 0, log[-1500:]
            res["decode_equals_reference_decode"] = [md5(p) for p in outs("back")] == [md5(p) for p in outs("rback")]
            print(json.dumps(res))
            return
        res["decode_of_ref_archive_ok"] = [md5(p) for p in outs("back")] == res["fastq_md5"]
        rc, log = po.run_ref(["-c", "-t", str(a.threads), "-f"] + inputs + ["-o", "mt"], tmp)
        assert rc == 0, log[-1500:]
        res["mt_archive_differs_from_t1"] = md5(os.path.join(tmp, "mt.arc")) != res["arc_md5_ref"]
        run([CLI, "-d", "-f", "mt.arc", "-o", "mtback"], tmp)
        res["decode_of_mt_archive_ok"] = [md5(p) for p in outs("mtback")] == res["fastq_md5"]
        print(json.dumps(res))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
