"""Whole-file parity at scale (not part of pytest: minutes of reference CPU time).

Generates the bench's synthetic 150 bp PE FASTQ for N default-size blocks, writes it to /dev/shm, and checks
  1. seqarc_b200 -c   == SeqArc-1.6 -c -t 1        (byte-identical .arc, default 50 MiB blocks)
  2. seqarc_b200 -d of the reference's archive      == the input FASTQ
  3. seqarc_b200 -d of a SeqArc-1.6 -t <cores> archive (blocks written out of order) == the input FASTQ
Prints one JSON line.  Run on the GPU box:  python tools/scale_check.py --blocks 24
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=24)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--lossy", type=float, default=0.0, help="R-Block ratio (-l); the reference then runs under the newline-guard shim")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda:0")
    tmp = tempfile.mkdtemp(prefix="sqscale_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        f1, f2 = os.path.join(tmp, "s_1.fq"), os.path.join(tmp, "s_2.fq")
        with open(f1, "wb") as x, open(f2, "wb") as y:
            for i in range(a.blocks):
                blk = bench.to_host(torch, bench.gen_block_device(torch, dev, i, bench.PAIRS_PER_BLOCK, 7))
                bench.fastq_text(blk, 0).tofile(x)
                bench.fastq_text(blk, 1).tofile(y)
        nbytes = os.path.getsize(f1) + os.path.getsize(f2)
        res = {"blocks": a.blocks, "fastq_bytes": nbytes, "fastq_md5": [md5(f1), md5(f2)]}
        lossy = ["-l", "%g" % a.lossy] if a.lossy else []
        res["lossy"] = a.lossy
        t0 = time.perf_counter()
        rc, log = po.run_ref(["-c", "-t", "1", "-f"] + lossy + ["-1", f1, "-2", f2, "-o", "ref"], tmp, lossy_shim=bool(lossy))
        assert rc == 0, log[-1500:]
        res["ref_c_t1_s"] = round(time.perf_counter() - t0, 1)
        res["ours_c_s"] = round(run([CLI, "-c", "-f"] + lossy + ["-1", f1, "-2", f2, "-o", "ours"], tmp), 1)
        res["arc_bytes"] = os.path.getsize(os.path.join(tmp, "ref.arc"))
        res["arc_md5_ref"], res["arc_md5_ours"] = md5(os.path.join(tmp, "ref.arc")), md5(os.path.join(tmp, "ours.arc"))
        res["archive_identical"] = res["arc_md5_ref"] == res["arc_md5_ours"]
        res["ours_d_s"] = round(run([CLI, "-d", "-f", "ref.arc", "-o", "back"], tmp), 1)
        if lossy:       # lossy: the decoded qualities are the smoothed ones -- compare with the reference's own decode
            rc, log = po.run_ref(["-d", "-t", str(a.threads), "-f", "ref.arc", "-o", "rback"], tmp)
            assert rc == 0, log[-1500:]
            res["decode_equals_reference_decode"] = [md5(os.path.join(tmp, "back_%d.fastq" % k)) for k in (1, 2)] == \
                                                    [md5(os.path.join(tmp, "rback_%d.fastq" % k)) for k in (1, 2)]
            print(json.dumps(res))
            return
        res["decode_of_ref_archive_ok"] = [md5(os.path.join(tmp, "back_1.fastq")), md5(os.path.join(tmp, "back_2.fastq"))] == res["fastq_md5"]
        rc, log = po.run_ref(["-c", "-t", str(a.threads), "-f", "-1", f1, "-2", f2, "-o", "mt"], tmp)
        assert rc == 0, log[-1500:]
        res["mt_archive_differs_from_t1"] = md5(os.path.join(tmp, "mt.arc")) != res["arc_md5_ref"]
        run([CLI, "-d", "-f", "mt.arc", "-o", "mtback"], tmp)
        res["decode_of_mt_archive_ok"] = [md5(os.path.join(tmp, "mtback_1.fastq")), md5(os.path.join(tmp, "mtback_2.fastq"))] == res["fastq_md5"]
        print(json.dumps(res))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
