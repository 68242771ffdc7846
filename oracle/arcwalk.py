"""TEST INFRASTRUCTURE (oracle side) -- .arc container walker.

Splits a SeqArc v1.6 archive into its EBML-style tagged pieces so that the
per-block, per-stream byte strings written by the reference binary can be
compared one by one with the oracle / CUDA outputs.

Format restated from SURVEY.md App. A (Encap @0x420680-0x4207bd,
SeqArcFile ctor/writeFileInfo @0x415c20/@0x4171b0 of /root/reference/SeqArc-1.6).
Only tests/, tools/ and bench.py's reference leg may import this.
"""
import struct

MAGIC = b".arc\x01\x06\x00"


def read_vint(buf, pos):
    """EBML vint: length n from the leading 1-bit of the first byte; value = remaining bits."""
    b0 = buf[pos]
    n = 1
    mask = 0x80
    while n <= 8 and not (b0 & mask):
        n += 1
        mask >>= 1
    if n > 8:
        raise ValueError("bad vint at %d" % pos)
    v = b0 & (mask - 1)
    for i in range(1, n):
        v = (v << 8) | buf[pos + i]
    return v, pos + n


def walk_children(buf, pos, end):
    """Yield (tag, payload_start, payload_end) for the tag/size records in buf[pos:end]."""
    while pos < end:
        tag, p = read_vint(buf, pos)
        size, p = read_vint(buf, p)
        yield tag, p, p + size
        pos = p + size


STREAM_NAMES = {1: "count", 4: "len", 5: "id", 6: "seq", 7: "qual", 23: "degetip", 14: "degech",
                24: "degemaxq", 25: "ndegecnt", 26: "ndegepos"}


def parse_arc(data):
    """Return dict(blocks=[{kind, raw, streams:{tag: bytes}, order:[tags]}], trailer={...})."""
    assert data[:7] == MAGIC, "not a v1.6 .arc"
    out = {"blocks": [], "trailer": {}}
    top = list(walk_children(data, 7, len(data)))
    assert top[0][0] == 2
    _, bs, be = top[0]
    for kind, s, e in walk_children(data, bs, be):
        blk = {"kind": kind, "start": s, "end": e, "streams": {}, "order": []}
        # whole record including its own tag+size header
        hdr = 1 + 4
        blk["raw"] = bytes(data[s - hdr:e])
        for tag, ps, pe in walk_children(data, s, e):
            blk["streams"][tag] = bytes(data[ps:pe])
            blk["order"].append(tag)
        out["blocks"].append(blk)
    if len(top) > 1:
        assert top[1][0] == 3
        _, ts, te = top[1]
        out["trailer_raw"] = bytes(data[ts - 5:te])
        for tag, ps, pe in walk_children(data, ts, te):
            if tag == 1:
                params = {}
                for pid, a, b in walk_children(data, ps, pe):
                    params[pid] = bytes(data[a:b])
                out["trailer"]["params"] = params
            elif tag == 6:
                out["trailer"]["model"] = bytes(data[ps:pe])
            elif tag == 7:
                out["trailer"]["index"] = bytes(data[ps:pe])
            else:
                out["trailer"][tag] = bytes(data[ps:pe])
    return out


def block_index(parsed, pe):
    """Decode the trailer block index (32 B/SE, 40 B/PE records)."""
    raw = parsed["trailer"]["index"]
    rec = 40 if pe else 32
    res = []
    for i in range(0, len(raw), rec):
        r = raw[i:i + rec]
        l2, s1, s2, z = struct.unpack_from("<IIII", r, 0)
        offs = struct.unpack_from("<%dQ" % ((rec - 16) // 8), r, 16)
        res.append({"len": l2 >> 1, "long": l2 & 1, "src1": s1, "src2": s2, "z": z, "offs": offs})
    return res


if __name__ == "__main__":
    import sys
    d = open(sys.argv[1], "rb").read()
    p = parse_arc(d)
    for i, b in enumerate(p["blocks"]):
        print("block", i, "kind", b["kind"], "bytes", len(b["raw"]),
              {STREAM_NAMES.get(t, t): len(v) for t, v in b["streams"].items()}, b["order"])
    tr = p["trailer"]
    if "params" in tr:
        for k, v in tr["params"].items():
            print(" param", k, len(v), v[:32].hex())
    print(" index", tr.get("index", b"").hex())
