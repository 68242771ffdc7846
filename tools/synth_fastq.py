"""Seeded synthetic FASTQ blocks of the shapes BASELINE.json names.

Blocks are produced directly in the layout the coder seam takes (concatenated ids / seq / qual +
length arrays, PE interleaved R1,R2,...) and can be rendered to FASTQ text for the reference binary.
Used by tests/, bench.py and tools/make_golden.py; not part of the product.
"""
import numpy as np

QLV8 = np.array([2, 6, 15, 22, 27, 33, 37, 40], dtype=np.uint8) + 33
QP8 = np.array([.01, .02, .03, .04, .05, .10, .35, .40])
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _digits(v, width):
    """(n,) ints -> (n, width) ASCII zero-padded digits."""
    out = np.empty((v.shape[0], width), dtype=np.uint8)
    x = v.astype(np.int64).copy()
    for k in range(width - 1, -1, -1):
        out[:, k] = 48 + (x % 10)
        x //= 10
    return out


def _const(n, s):
    return np.broadcast_to(np.frombuffer(s, dtype=np.uint8), (n, len(s)))


def illumina_ids(n_templates, rng, first=0, mate=None, lane=1):
    """'A00123:45:HXXXXDSXX:<lane>:<tile>:<x>:<y> <mate>:N:0:ACGTACGT' (zero-padded fixed width).
    tile advances every 4096 templates, x / y uniform 1000..32000."""
    idx = first + np.arange(n_templates)
    tile = 1101 + (idx // 4096) % 8000
    x = rng.integers(1000, 32000, n_templates)
    y = rng.integers(1000, 32000, n_templates)
    cols = [_const(n_templates, b"A00123:45:HXXXXDSXX:%d:" % lane), _digits(tile, 4), _const(n_templates, b":"),
            _digits(x, 5), _const(n_templates, b":"), _digits(y, 5)]
    if mate is not None:
        cols.append(_const(n_templates, b" %d:N:0:ACGTACGT" % mate))
    return np.concatenate(cols, axis=1)


def srr_ids(n_templates, first=0, mate=None):
    """'SRR1234567.<n> <n>/<mate>' -- variable width, rule-predictable (bOrderBin 1)."""
    ids = []
    for i in range(first + 1, first + 1 + n_templates):
        s = b"SRR1234567.%d %d" % (i, i)
        if mate is not None:
            s += b"/%d" % mate
        ids.append(s)
    return ids


def bgi_ids(n_templates, first=0, mate=None):
    """'V300012345L1C001R0010000001/1' -- BGI / MGI style: lane, column, row, running number, mate suffix."""
    ids = []
    for i in range(first, first + n_templates):
        s = b"V300012345L1C%03dR%03d%07d" % (1 + (i // 9000) % 8, 1 + (i // 300) % 30, i % 300 * 13 + 1)
        if mate is not None:
            s += b"/%d" % mate
        ids.append(s)
    return ids


def old_illumina_ids(n_templates, rng, first=0, mate=None):
    """'HWUSI-EAS100R:6:73:941:1973#0/1' -- pre-CASAVA-1.8 style, coordinates not zero padded (variable width)."""
    x = rng.integers(1, 20000, n_templates)
    y = rng.integers(1, 20000, n_templates)
    ids = []
    for i in range(n_templates):
        s = b"HWUSI-EAS100R:6:%d:%d:%d#0" % (1 + ((first + i) // 2000) % 120, x[i], y[i])
        if mate is not None:
            s += b"/%d" % mate
        ids.append(s)
    return ids


def sra_len_ids(n_templates, rng, first=0):
    """'SRR001666.1 071112_SLXA-EAS1_s_7:5:1:817:345 length=100' -- SRA dump with the original name and a length field."""
    x = rng.integers(1, 2000, n_templates)
    y = rng.integers(1, 2000, n_templates)
    return [b"SRR001666.%d 071112_SLXA-EAS1_s_7:5:%d:%d:%d length=100" % (first + i + 1, 1 + ((first + i) // 500) % 300, x[i], y[i])
            for i in range(n_templates)]


def make_block(kind="illumina_pe150", n_templates=1000, seed=7, first=0, n_rate=0.0, hash_tail=0.0):
    """Return dict(ids, idlen, seq, qual, seqlen) numpy arrays of one block.

    kind: 'illumina_pe150' (config 2/3/5: 150 bp PE, Q ~ clipped N(36,5)),
          'se100_q8'       (config 4: 100 bp SE, 8-level binned qualities),
          'srr_pe100'      (rule-predictable IDs, 100 bp PE).
    n_rate: fraction of bases replaced by N with quality '#'/'!' (exercises the Dege streams).
    hash_tail: fraction of reads whose quality ends in a run of '#' (exercises symbol 94).
    """
    rng = np.random.default_rng(seed)
    if kind == "long_se":           # a few long reads (> 65,535 bases: encode_len_long), lengths fixed, n_templates <= 5
        lens = [65536, 300, 70001, 1200, 66000][:n_templates]
        idl = [b"read%d len=%d" % (first + i, n) for i, n in enumerate(lens)]
        tot = sum(lens)
        return dict(ids=np.frombuffer(b"".join(idl), dtype=np.uint8).copy(), idlen=np.array([len(x) for x in idl], np.uint16),
                    seq=BASES[rng.integers(0, 4, tot)], qual=(np.clip(np.rint(rng.normal(20, 6, tot)), 2, 40).astype(np.uint8) + 33),
                    seqlen=np.array(lens, np.int32), pe=False)
    if kind == "illumina_pe_trim":  # adapter-trimmed pairs: every read keeps a random prefix, mate 2 shorter on average,
        full = make_block("illumina_pe150", n_templates, seed, first, n_rate, hash_tail)   # so the mate files differ in size
        n = full["seqlen"].shape[0]
        keep = np.where(np.arange(n) % 2 == 0, rng.integers(100, 151, n), rng.integers(40, 151, n)).astype(np.int32)
        m = (np.arange(150)[None, :] < keep[:, None]).reshape(-1)
        return dict(ids=full["ids"], idlen=full["idlen"], seq=full["seq"][m].copy(), qual=full["qual"][m].copy(), seqlen=keep, pe=True)
    pe = kind not in ("se100_q8", "sra_len_se100")
    L = 150 if kind == "illumina_pe150" else 100
    n = n_templates * (2 if pe else 1)
    if kind in ("srr_pe100", "bgi_pe100", "old_illumina_pe100", "sra_len_se100"):
        if kind == "srr_pe100":
            a, b = srr_ids(n_templates, first, 1), srr_ids(n_templates, first, 2)
        elif kind == "bgi_pe100":
            a, b = bgi_ids(n_templates, first, 1), bgi_ids(n_templates, first, 2)
        elif kind == "old_illumina_pe100":
            a = old_illumina_ids(n_templates, np.random.default_rng(seed + 1000003), first, 1)
            b = old_illumina_ids(n_templates, np.random.default_rng(seed + 1000003), first, 2)
        else:
            a, b = sra_len_ids(n_templates, np.random.default_rng(seed + 1000003), first), None
        if pe:
            idl = [None] * n
            idl[0::2], idl[1::2] = a, b
        else:
            idl = a
        ids = np.frombuffer(b"".join(idl), dtype=np.uint8).copy()
        idlen = np.array([len(s) for s in idl], dtype=np.uint16)
    else:
        if pe:
            rng_id = np.random.default_rng(seed + 1000003)
            m1 = illumina_ids(n_templates, rng_id, first, 1)
            rng_id = np.random.default_rng(seed + 1000003)
            m2 = illumina_ids(n_templates, rng_id, first, 2)
            mat = np.empty((n, m1.shape[1]), dtype=np.uint8)
            mat[0::2], mat[1::2] = m1, m2
        else:
            mat = illumina_ids(n_templates, rng, first, None)
        ids = np.ascontiguousarray(mat).reshape(-1)
        idlen = np.full(n, mat.shape[1], dtype=np.uint16)
    seq = BASES[rng.integers(0, 4, (n, L))]
    if kind == "se100_q8":
        qual = QLV8[rng.choice(8, (n, L), p=QP8)]
    else:
        qual = (np.clip(np.rint(rng.normal(36, 5, (n, L))), 2, 40).astype(np.uint8) + 33)
    if n_rate > 0:
        m = rng.random((n, L)) < n_rate
        seq = np.where(m, np.uint8(ord("N")), seq)
        qual = np.where(m, np.where(rng.random((n, L)) < 0.5, np.uint8(ord("#")), np.uint8(ord("!"))), qual)
    if hash_tail > 0:
        rows = np.nonzero(rng.random(n) < hash_tail)[0]
        tl = rng.integers(1, L // 2, rows.shape[0])
        for r, t in zip(rows, tl):
            qual[r, L - t:] = ord("#")
    return dict(ids=ids, idlen=idlen, seq=np.ascontiguousarray(seq).reshape(-1),
                qual=np.ascontiguousarray(qual).reshape(-1), seqlen=np.full(n, L, dtype=np.int32), pe=pe)


def to_fastq(blk):
    """Render a block as FASTQ text: bytes (SE) or (bytes, bytes) (PE, de-interleaved)."""
    outs = ([], []) if blk["pe"] else ([],)
    io = so = 0
    ids, seq, qual = blk["ids"].tobytes(), blk["seq"].tobytes(), blk["qual"].tobytes()
    for i in range(blk["seqlen"].shape[0]):
        il, sl = int(blk["idlen"][i]), int(blk["seqlen"][i])
        outs[i % 2 if blk["pe"] else 0].append(b"@" + ids[io:io + il] + b"\n" + seq[so:so + sl] + b"\n+\n" +
                                                qual[so:so + sl] + b"\n")
        io += il
        so += sl
    res = tuple(b"".join(o) for o in outs)
    return res if blk["pe"] else res[0]


def fastq_bytes(blk):
    """Size of the FASTQ text of this block without rendering it ('@', '+', four newlines per read)."""
    n = blk["seqlen"].shape[0]
    return int(blk["ids"].shape[0] + 2 * blk["seq"].shape[0] + 6 * n)
