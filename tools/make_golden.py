"""TEST INFRASTRUCTURE -- regenerate tests/golden/ by running the UNMODIFIED reference binary.

Run in the build container (needs oracle/_ref/SeqArc-1.6, staged by `make -C oracle ref` from
/root/reference).  Every case writes   tests/golden/<name>.arc   (what `SeqArc-1.6 -c -t 1` wrote)
and an entry in tests/golden/manifest.json (how the input is rebuilt, md5 of the FASTQ text,
options).  Inputs are either the committed tests/golden/err_pe1000_{1,2}.fq.gz (first 1000 pairs
of the reference's own test/ERR2755197_test_{1,2}.fq) or seeded synthetic blocks
(tools/synth_fastq.py), so nothing under /root/reference is needed at test time.
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tools import synth_fastq as sf  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, input spec, reference options, seqarc.config
    dict(name="err_pe1000", src="err", pe=True, opts=[], config=None),
    dict(name="err_pe1000_nomd5", src="err", pe=True, opts=["-n"], config=None),
    dict(name="illumina_pe150", src=dict(kind="illumina_pe150", n_templates=1500, seed=7), pe=True, opts=[], config=None),
    dict(name="illumina_pe150_n", src=dict(kind="illumina_pe150", n_templates=1500, seed=11, n_rate=0.004, hash_tail=0.05),
         pe=True, opts=[], config=None),
    dict(name="illumina_pe150_lossy", src=dict(kind="illumina_pe150", n_templates=1500, seed=7), pe=True,
         opts=["-l", "1.15"], config=None, shim=True),
    dict(name="se100_q8", src=dict(kind="se100_q8", n_templates=4000, seed=12345), pe=False, opts=[], config=None),
    dict(name="se100_q8_3blocks", src=dict(kind="se100_q8", n_templates=12000, seed=5), pe=False, opts=[],
         config={"BlockSize(M)": 1}),
    dict(name="illumina_pe150_3blocks", src=dict(kind="illumina_pe150", n_templates=3200, seed=31), pe=True, opts=[],
         config={"BlockSize(M)": 1}),
    dict(name="srr_pe100", src=dict(kind="srr_pe100", n_templates=1200, seed=3), pe=True, opts=[], config=None),
    dict(name="illumina_pe150_q1", src=dict(kind="illumina_pe150", n_templates=800, seed=21, hash_tail=0.1), pe=True, opts=[],
         config={"Qlevel": 1}),
    # usemodel = 1 (snapshot trained on the head blocks): forced by the size-faking shim, archive too big to commit
    # (27 MB model blob) -> only the block records, the block index and the md5 of the model blob are kept
    dict(name="se100_q8_usemodel", src=dict(kind="se100_q8", n_templates=14000, seed=77), pe=False, opts=[],
         config={"BlockSize(M)": 1}, fake_size=15000000000),
    dict(name="illumina_pe150_usemodel", src=dict(kind="illumina_pe150", n_templates=3200, seed=41), pe=True, opts=[],
         config={"BlockSize(M)": 1}, fake_size=8000000000),
    # BASELINE configs[4]: a random reference index is given and nothing maps ("Align block:0"): doAlignEncode framing
    dict(name="illumina_pe150_ref", src=dict(kind="illumina_pe150", n_templates=1500, seed=7), pe=True, opts=[], config=None,
         ref=dict(n=2000000, seed=5)),
    dict(name="se100_q8_ref", src=dict(kind="se100_q8", n_templates=4000, seed=12345), pe=False, opts=[], config=None,
         ref=dict(n=2000000, seed=5)),
    dict(name="err_pe1000_ref", src="err", pe=True, opts=[], config=None, ref=dict(n=2000000, seed=5)),
    dict(name="illumina_pe150_ref_3blocks", src=dict(kind="illumina_pe150", n_templates=3200, seed=31), pe=True, opts=[],
         config={"BlockSize(M)": 1}, ref=dict(n=2000000, seed=5)),
    # '+' lines that repeat the ID: same block records, trailer param 2 = 0 (and 0 instead of 1 in the SE index records)
    dict(name="se100_q8_plus", src=dict(kind="se100_q8", n_templates=4000, seed=12345), pe=False, opts=[], config=None, plus=True),
    dict(name="illumina_pe150_plus_3blocks", src=dict(kind="illumina_pe150", n_templates=3200, seed=31), pe=True, opts=[],
         config={"BlockSize(M)": 1}, plus=True),
    # reads longer than 65,535 bases: compressLen_long @0x423710 and the long flag in the block index
    dict(name="long_se", src=dict(kind="long_se", n_templates=4, seed=9), pe=False, opts=[], config=None),
    # more ID styles for the ID-bin analysis / name model (BGI, pre-1.8 Illumina, SRA dump with a length field)
    dict(name="bgi_pe100", src=dict(kind="bgi_pe100", n_templates=1200, seed=13), pe=True, opts=[], config=None),
    dict(name="old_illumina_pe100", src=dict(kind="old_illumina_pe100", n_templates=1200, seed=14), pe=True, opts=[], config=None),
    dict(name="sra_len_se100", src=dict(kind="sra_len_se100", n_templates=2500, seed=15), pe=False, opts=[], config=None),
    # the extreme base-context orders: Slevel 0 (order 7) and Slevel 5 (order 12)
    dict(name="illumina_pe150_s0", src=dict(kind="illumina_pe150", n_templates=700, seed=23, n_rate=0.001), pe=True, opts=[],
         config={"Slevel": 0}),
    dict(name="illumina_pe150_s5", src=dict(kind="illumina_pe150", n_templates=700, seed=24, n_rate=0.001), pe=True, opts=[],
         config={"Slevel": 5}),
    # mate files of unequal size (trimmed reads): one file runs dry first, doReadPEJob keeps cutting until both do
    dict(name="illumina_pe_trim_blocks", src=dict(kind="illumina_pe_trim", n_templates=7000, seed=51), pe=True, opts=[],
         config={"BlockSize(M)": 1}),
    # a block that DOES hold mapped reads: reads sampled from the reference FASTA itself, a tenth random.
    # Pins the coders of the aligned-mode side streams (Order / Pos / Mis / Rev / CigaL / CigaV) for the oracle.
    dict(name="aligned_se100", src="aligned", pe=False, opts=[], config=None, aligned=dict(n=200000, seed=11, reads=3000, L=100),
         ref=dict(n=200000, seed=11)),
    dict(name="aligned_se_var", src="aligned", pe=False, opts=[], config=None,
         aligned=dict(n=1000000, seed=12, reads=2500, L=150, Lmin=60, n_rate=0.004), ref=dict(n=1000000, seed=12)),
    # three sequences in the FASTA: positions run through their plain concatenation
    dict(name="aligned_se_multi", src="aligned", pe=False, opts=[], config=None,
         aligned=dict(n=0, seed=31, reads=2000, L=100, chroms=[120000, 90001, 50030]), ref=dict(n=0, seed=31)),
    dict(name="aligned_pe100", src="aligned", pe=True, opts=[], config=None, aligned=dict(n=300000, seed=21, pairs=2000, L=100),
         ref=dict(n=300000, seed=21)),
    dict(name="illumina_pe150_q3s2", src=dict(kind="illumina_pe150", n_templates=800, seed=22, n_rate=0.001), pe=True, opts=[],
         config={"Qlevel": 3, "Slevel": 2}),
]


def case_fastq(case):
    """FASTQ text of a case: (r1, r2) for PE, (r1,) for SE."""
    if case["src"] == "err":
        return tuple(gzip.open(os.path.join(GOLD, "err_pe1000_%d.fq.gz" % k), "rb").read() for k in (1, 2))
    if case["src"] == "aligned":
        return aligned_case(case)[1]
    blk = sf.make_block(**case["src"])
    t = sf.to_fastq(blk)
    t = t if case["pe"] else (t,)
    return tuple(plus_lines(x) for x in t) if case.get("plus") else t


def plus_lines(text):
    """Rewrite every bare '+' line as '+<ID>'."""
    lines = text.split(b"\n")
    for i in range(2, len(lines) - 1, 4):
        lines[i] = b"+" + lines[i - 2][1:]
    return b"\n".join(lines)


def aligned_reads(n, seed, reads, L, Lmin=None, n_rate=0.0, chroms=None):
    """(FASTA, FASTQ): one random chromosome and single-end reads sampled from it -- 0..2 substitutions each, half of them
    reverse-complemented, a tenth of the reads random (they stay unmapped).  One generator draws the reference first.
    Lmin: read lengths uniform in [Lmin, L]; n_rate: bases turned into N with quality '#' (degenerate-base overlay);
    chroms: lengths of several sequences instead of one of n bases (70 bases per line, header lines with a comment)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    if chroms:
        parts = [acgt[rng.integers(0, 4, k)] for k in chroms]
        fa = b"".join(b">chr%d some text\n" % (k + 1) + b"".join(c[i:i + 70].tobytes() + b"\n" for i in range(0, len(c), 70))
                      for k, c in enumerate(parts))
    else:
        ref = acgt[rng.integers(0, 4, n)]
        fa = b">chr1\n" + b"".join(ref[i:i + 60].tobytes() + b"\n" for i in range(0, n, 60))
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    recs = []
    Lmax = L
    for i in range(reads):
        if Lmin is not None:
            L = int(rng.integers(Lmin, Lmax + 1))
        kind = rng.random()
        if kind < 0.1:
            s = acgt[rng.integers(0, 4, L)]
        else:
            if chroms:
                ref = parts[int(rng.integers(0, len(parts)))]
                n = len(ref)
            p = int(rng.integers(0, n - L))
            s = ref[p:p + L].copy()
            for _ in range(int(rng.integers(0, 3))):
                k = int(rng.integers(0, L))
                s[k] = acgt[rng.integers(0, 4)]
            if rng.random() < 0.5:
                s = np.array([comp[int(c)] for c in s[::-1]], np.uint8)
        q = (np.clip(np.rint(rng.normal(30, 6, L)), 2, 40).astype(np.uint8) + 33)
        if n_rate:
            nm = rng.random(L) < n_rate
            s = np.where(nm, ord("N"), s).astype(np.uint8)
            q = np.where(nm, ord("#"), q).astype(np.uint8)
        recs.append(b"@r%d\n" % i + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return fa, b"".join(recs)


def aligned_pairs(n, seed, pairs, L):
    """(FASTA, FASTQ 1, FASTQ 2): pairs sampled from one random chromosome, inserts of 200..399 bp, 0..2 substitutions per mate,
    mate 2 reverse-complemented, half of the pairs with the mates swapped, 8 % of the pairs random (unmapped)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    ref = acgt[rng.integers(0, 4, n)]
    fa = b">chr1\n" + b"".join(ref[i:i + 60].tobytes() + b"\n" for i in range(0, n, 60))
    comp = {65: 84, 67: 71, 71: 67, 84: 65}

    def revcomp(s):
        return np.array([comp[int(c)] for c in s[::-1]], np.uint8)

    def mutate(s):
        s = s.copy()
        for _ in range(int(rng.integers(0, 3))):
            s[int(rng.integers(0, L))] = acgt[rng.integers(0, 4)]
        return s

    r1, r2 = [], []
    for i in range(pairs):
        q1 = np.clip(np.rint(rng.normal(30, 6, L)), 2, 40).astype(np.uint8) + 33
        q2 = np.clip(np.rint(rng.normal(30, 6, L)), 2, 40).astype(np.uint8) + 33
        if rng.random() < 0.08:
            a, b = acgt[rng.integers(0, 4, L)], acgt[rng.integers(0, 4, L)]
        else:
            ins = int(rng.integers(200, 400))
            p = int(rng.integers(0, n - ins))
            a, b = mutate(ref[p:p + L]), revcomp(mutate(ref[p + ins - L:p + ins]))
            if rng.random() < 0.5:
                a, b = b, a
        r1.append(b"@p%d/1\n" % i + a.tobytes() + b"\n+\n" + q1.tobytes() + b"\n")
        r2.append(b"@p%d/2\n" % i + b.tobytes() + b"\n+\n" + q2.tobytes() + b"\n")
    return fa, b"".join(r1), b"".join(r2)


def aligned_case(case):
    """(FASTA, (FASTQ, ...)) of a case whose reads are sampled from its own reference."""
    if case["pe"]:
        fa, f1, f2 = aligned_pairs(**case["aligned"])
        return fa, (f1, f2)
    fa, fq = aligned_reads(**case["aligned"])
    return fa, (fq,)


def random_reference(n, seed):
    """FASTA text of one random chromosome (numpy default_rng(seed), 60 bases per line)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
    return b">chr1\n" + b"".join(s[i:i + 60].tobytes() + b"\n" for i in range(0, n, 60))


def main():
    only = set(sys.argv[1:])        # optional: regenerate only the named cases, keep the other manifest entries
    os.makedirs(GOLD, exist_ok=True)
    assert po.have_ref(), "reference binary not staged (make -C oracle ref)"
    # committed real-data subset: first 1000 pairs of the reference's own test files
    for k in (1, 2):
        dst = os.path.join(GOLD, "err_pe1000_%d.fq.gz" % k)
        if not os.path.exists(dst):
            src = "/root/reference/test/ERR2755197_test_%d.fq" % k
            lines = open(src, "rb").read().split(b"\n")[:4000]
            with open(dst, "wb") as f:
                with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:
                    g.write(b"\n".join(lines) + b"\n")
    manifest = json.load(open(os.path.join(GOLD, "manifest.json"))) if only else {}
    for case in CASES:
        if only and case["name"] not in only:
            continue
        fq = case_fastq(case)
        tmp = tempfile.mkdtemp(prefix="gold_")
        try:
            names = []
            for k, t in enumerate(fq):
                p = os.path.join(tmp, "in_%d.fq" % (k + 1))
                open(p, "wb").write(t)
                names.append(p)
            args = ["-c", "-t", "1", "-f"] + case["opts"]
            ref_md5 = None
            if case.get("ref"):     # `SeqArc -i` writes a ~2 GB hash index beside the FASTA (temporary)
                fa = aligned_case(case)[0] if case.get("aligned") else random_reference(**case["ref"])
                open(os.path.join(tmp, "ref.fa"), "wb").write(fa)
                rc, log = po.run_ref(["-i", "ref.fa"], tmp)
                assert rc == 0, log[-2000:]
                ref_md5 = open(os.path.join(tmp, "ref.fa.md5"), "rb").read().hex()
                assert ref_md5 == hashlib.md5(fa).hexdigest()
                args += ["ref.fa"]
            if case["pe"]:
                args += ["-1", names[0], "-2", names[1]]
            else:
                args += ["-1", names[0]]
            args += ["-o", "out"]
            rc, log = po.run_ref(args, tmp, lossy_shim=case.get("shim", False), config=case["config"],
                                 fake_size=case.get("fake_size"))
            assert rc == 0, (case["name"], rc, log[-2000:])
            if case.get("aligned"):
                assert "Align block:1" in log, log[-1500:]
            elif case.get("ref"):
                assert "Align block:0" in log, log[-1500:]
            arc = open(os.path.join(tmp, "out.arc"), "rb").read()
            entry = dict(src=case["src"], pe=case["pe"], opts=case["opts"], config=case["config"],
                         shim=case.get("shim", False), arc_bytes=len(arc), arc_md5=hashlib.md5(arc).hexdigest(),
                         fastq_md5=[hashlib.md5(t).hexdigest() for t in fq])
            if ref_md5:
                entry.update(ref=case["ref"], ref_md5=ref_md5)
            if case.get("aligned"):
                entry.update(aligned=case["aligned"])
            if case.get("plus"):
                entry.update(plus=True)
            if case.get("fake_size"):
                from oracle import arcwalk
                assert "usemodel=1" in log, log[-1500:]
                pa = arcwalk.parse_arc(arc)
                model = pa["trailer"]["model"]
                assert model[:4] == bytes([0x81, 0x82, 0x01, 0x00])
                # stripped archive: everything except the model blob (header + blocks | trailer without tag 6)
                cut0 = arc.index(model[:64]) - 5          # tag 6 + size(4) precede the payload
                stripped = arc[:cut0] + arc[cut0 + 5 + len(model):]
                open(os.path.join(GOLD, case["name"] + ".arc_nomodel"), "wb").write(stripped)
                entry.update(usemodel=1, model_bytes=len(model) - 4, model_md5=hashlib.md5(model[4:]).hexdigest(),
                             model_cut=cut0, fake_size=case["fake_size"])
            else:
                open(os.path.join(GOLD, case["name"] + ".arc"), "wb").write(arc)
            manifest[case["name"]] = entry
            print(case["name"], len(arc), "bytes")
        finally:
            shutil.rmtree(tmp)
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
