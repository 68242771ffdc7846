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
