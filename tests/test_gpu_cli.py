"""GPU: the host driver (fastqueeze_b200/seqarc_b200: C++ block cutter + .arc packer over the C ABI).

Whole-file contract of BASELINE.json: a GPU-written .arc is byte-identical to what the reference writes
at -t 1, decodes with the reference decompressor, and reference-written archives decode here."""
import hashlib
import os
import subprocess
import tempfile

import pytest

from tests import common
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
CLI = os.path.join(common.ROOT, "fastqueeze_b200", "seqarc_b200")

# every golden archive, rule-predictable IDs (ERR / SRR style) included: byte identity expected
IDENTICAL = ["illumina_pe_trim_blocks", "err_pe1000", "err_pe1000_nomd5", "srr_pe100", "illumina_pe150", "illumina_pe150_n", "illumina_pe150_3blocks", "illumina_pe150_lossy", "illumina_pe150_q1",
             "illumina_pe150_q3s2", "se100_q8", "se100_q8_3blocks", "se100_q8_usemodel", "illumina_pe150_usemodel",
             # BASELINE configs[4]: reference index given, nothing maps
             "illumina_pe150_ref", "se100_q8_ref", "err_pe1000_ref", "illumina_pe150_ref_3blocks",
             # "+<ID>" separator lines
             "se100_q8_plus", "illumina_pe150_plus_3blocks",
             # reads longer than 65,535 bases (compressLen_long, long flag in the block index)
             "long_se",
             # more ID styles (BGI, pre-1.8 Illumina, SRA dump with a length field)
             "bgi_pe100", "old_illumina_pe100", "sra_len_se100"]


def _setup(name, d):
    m = common.MANIFEST[name]
    fq = common.case_fastq(name)
    paths = []
    for k, t in enumerate(fq):
        # same base names as when the golden was made (they are stored in the trailer)
        p = os.path.join(d, "in_%d.fq" % (k + 1))
        open(p, "wb").write(t)
        paths.append(p)
    if m.get("ref"):
        # the CLI takes the reference's MD5 from the <ref>.md5 file `SeqArc -i` leaves beside the index, or else
        # digests the FASTA itself: the SE case exercises the second path
        if m["pe"]:
            open(os.path.join(d, "ref.fa.md5"), "wb").write(bytes.fromhex(m["ref_md5"]))
        else:
            from tools.make_golden import random_reference
            open(os.path.join(d, "ref.fa"), "wb").write(random_reference(**m["ref"]))
    cfg = m["config"] or {}
    if cfg:
        with open(os.path.join(d, "seqarc.config"), "w") as f:
            for k, v in cfg.items():
                f.write("%s:%s\n" % (k, v))
    return m, fq, paths


def _compress(m, paths, d, out="out", extra=()):
    args = [CLI, "-c", "-f"] + (["ref.fa"] if m.get("ref") else []) + ["-1", paths[0]] + (["-2", paths[1]] if m["pe"] else []) + ["-o", out] + m["opts"]
    if m.get("usemodel"):
        args += ["--usemodel", "1"]          # the golden was made with the size-faking shim
    args += list(extra)
    r = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return open(os.path.join(d, out + ".arc"), "rb").read()


@pytest.mark.parametrize("name", IDENTICAL)
def test_cli_archive_is_byte_identical_to_reference(name):
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        arc = _compress(m, paths, d)
    assert len(arc) == m["arc_bytes"]
    assert hashlib.md5(arc).hexdigest() == m["arc_md5"]


@pytest.mark.parametrize("name", ["err_pe1000", "illumina_pe150_3blocks", "se100_q8"])
def test_cli_gzip_input_gives_the_reference_archive(name):
    """.gz input: the reference writes the same archive as for the plain files except trailer param 4 = 1 (measured with
    SeqArc-1.6 on err_pe1000 and the 3-block PE case; the stored names drop the .gz suffix)."""
    import gzip
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        gz = []
        for p in paths:
            with open(p + ".gz", "wb") as f:
                with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:
                    g.write(open(p, "rb").read())
            os.remove(p)
            gz.append(p + ".gz")
        arc = _compress(m, gz, d)
    want = common.case_archive(name)
    k = want.rindex(b"\x84\x81\x00\x85\x81\x01")          # params 4 and 5 in the trailer
    want = want[:k + 2] + b"\x01" + want[k + 3:]
    assert arc == want


@pytest.mark.parametrize("name", [c for c in common.CASES if "lossy" not in c])
def test_cli_decodes_reference_archive(name):
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        open(os.path.join(d, "ref.arc"), "wb").write(common.case_archive(name))
        r = subprocess.run([CLI, "-d", "-f", "ref.arc", "-o", "back"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs = ["back_1.fastq", "back_2.fastq"] if m["pe"] else ["back.fastq"]
        for o, t in zip(outs, fq):
            assert open(os.path.join(d, o), "rb").read() == t


@pytest.mark.parametrize("name", ["err_pe1000", "srr_pe100", "illumina_pe150_n", "se100_q8_3blocks"])
def test_reference_decodes_cli_archive(name):
    if not po.have_ref():
        pytest.skip("reference binary not staged")
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        _compress(m, paths, d, out="mine")
        rc, log = po.run_ref(["-d", "-f", "mine.arc", "-o", "back"], d, config=m["config"])
        assert rc == 0, log[-1500:]
        outs = ["back_1.fastq", "back_2.fastq"] if m["pe"] else ["back.fastq"]
        for o, t in zip(outs, fq):
            assert open(os.path.join(d, o), "rb").read() == t


def test_config1_reference_test_pair_matches_survey_md5():
    """BASELINE configs[0]: the reference's own test/ERR2755197_test_{1,2}.fq (committed gzipped; 10,000 pairs, rule-predictable
    IDs).  SURVEY section 4 recorded what SeqArc-1.6 -c -t 1 writes for it: 821,500 bytes, md5 c8ba98391ac4607d9559db23875152ba,
    and under the newline-guard shim with -l 1.15: 224,143 bytes, md5 e21845c348c28dbe7cb54c9d627869e2."""
    import gzip
    with tempfile.TemporaryDirectory() as d:
        paths = []
        texts = []
        for k in (1, 2):
            t = gzip.open(os.path.join(common.GOLD, "ERR2755197_test_%d.fq.gz" % k), "rb").read()
            p = os.path.join(d, "ERR2755197_test_%d.fq" % k)
            open(p, "wb").write(t)
            paths.append(p)
            texts.append(t)
        for extra, size, md5 in (([], 821500, "c8ba98391ac4607d9559db23875152ba"), (["-l", "1.15"], 224143, "e21845c348c28dbe7cb54c9d627869e2")):
            r = subprocess.run([CLI, "-c", "-f", "-1", paths[0], "-2", paths[1], "-o", "pe"] + extra, cwd=d, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr
            arc = open(os.path.join(d, "pe.arc"), "rb").read()
            assert len(arc) == size
            assert hashlib.md5(arc).hexdigest() == md5
            if not extra:
                r = subprocess.run([CLI, "-d", "-f", "pe.arc", "-o", "back"], cwd=d, capture_output=True, text=True, timeout=600)
                assert r.returncode == 0, r.stderr
                assert open(os.path.join(d, "back_1.fastq"), "rb").read() == texts[0]
                assert open(os.path.join(d, "back_2.fastq"), "rb").read() == texts[1]


def _ngpu():
    import fastqueeze_b200
    return int(fastqueeze_b200.lib().sq_device_count())


@pytest.mark.parametrize("name", ["illumina_pe150_3blocks", "se100_q8_3blocks", "illumina_pe_trim_blocks"])
def test_cli_small_batches_keep_block_order(name):
    """One block per GPU call (-b 1): the batches go through the reader -> GPU -> ordered-writer pipeline one after
    the other and the archive is still the reference's -t 1 archive, block for block."""
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        arc = _compress(m, paths, d, extra=["-b", "1"])
        assert hashlib.md5(arc).hexdigest() == m["arc_md5"]
        r = subprocess.run([CLI, "-d", "-f", "out.arc", "-o", "back", "-b", "1"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs = ["back_1.fastq", "back_2.fastq"] if m["pe"] else ["back.fastq"]
        for o, t in zip(outs, fq):
            assert open(os.path.join(d, o), "rb").read() == t


@pytest.mark.parametrize("ng", [2, 4, 8])
@pytest.mark.parametrize("name", ["illumina_pe150_3blocks", "illumina_pe_trim_blocks", "se100_q8_usemodel"])
def test_cli_sharded_over_gpus_is_byte_identical(name, ng):
    """The job sharded over ng GPUs (one host thread + one context per GPU, batches of one block dealt out, records
    gathered in block order): the archive is byte-identical to SeqArc-1.6 -t 1, and the sharded decode restores the input."""
    if _ngpu() < ng:
        pytest.skip("needs %d GPUs" % ng)
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup(name, d)
        devs = ",".join(str(i) for i in range(ng))
        arc = _compress(m, paths, d, extra=["-g", devs, "-b", "1"])
        if m.get("usemodel"):
            want = common.case_archive(name)
            assert arc == want
        else:
            assert hashlib.md5(arc).hexdigest() == m["arc_md5"]
        r = subprocess.run([CLI, "-d", "-f", "out.arc", "-o", "back", "-g", devs], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs = ["back_1.fastq", "back_2.fastq"] if m["pe"] else ["back.fastq"]
        for o, t in zip(outs, fq):
            assert open(os.path.join(d, o), "rb").read() == t


def test_cli_decompress_without_prefix_reuses_input_names():
    """README.md:97: without a prefix the outputs are named exactly like the compression inputs (measured with the
    reference: `SeqArc -d x.arc` writes a_1.fq / a_2.fq into the current directory)."""
    with tempfile.TemporaryDirectory() as d:
        m, fq, paths = _setup("err_pe1000", d)
        _compress(m, paths, d)
        sub = os.path.join(d, "sub")
        os.mkdir(sub)
        r = subprocess.run([CLI, "-d", "../out.arc"], cwd=sub, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        for p, t in zip(paths, fq):
            assert open(os.path.join(sub, os.path.basename(p)), "rb").read() == t
