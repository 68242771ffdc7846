"""CPU: the C-ABI library loads, exports every symbol include/seqarc_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from tests.common import ROOT


def _build():
    import __graft_entry__
    __graft_entry__.build()


def test_exports_match_header():
    _build()
    from fastqueeze_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "seqarc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import fastqueeze_b200 as fq
    with pytest.raises(fq.SeqArcB200Error) as e:
        fq.EncapFqzComp()
    assert e.value.code == -1


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fastqueeze_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "codec.py" and "oracle" not in txt.lower(), f


def test_ctypes_structs_match_the_header(tmp_path):
    """The Python mirror declares the ABI structs by hand: their sizes and field offsets must equal what a C
    compiler sees in include/seqarc_b200.h (a drifted layout would corrupt every call silently)."""
    import subprocess
    from fastqueeze_b200 import _lib
    structs = {"sq_params": _lib.SqParams, "sq_block_in": _lib.SqBlockIn, "sq_block_out": _lib.SqBlockOut,
               "sq_block_dec": _lib.SqBlockDec, "sq_side_stream": _lib.SqSideStream,
               "sq_align_block": _lib.SqAlignBlock}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "seqarc_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        src.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            src.append('printf(" %s=%%zu", offsetof(%s, %s));' % (fname, cname, fname))
        src.append('printf("\\n");')
    src += ['return 0;', '}']
    c = tmp_path / "abi.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)])
    out = subprocess.check_output([str(exe)]).decode().splitlines()
    for line in out:
        parts = line.split()
        ct = structs[parts[0]]
        assert int(parts[1]) == ctypes.sizeof(ct), line
        for kv in parts[2:]:
            k, v = kv.split("=")
            assert getattr(ct, k).offset == int(v), (parts[0], k)
