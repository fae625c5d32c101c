"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and the host mirror fails loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "stark_rings_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", txt))
    names = {n for n in names if "##" not in n}
    for tag in ("gl", "bb", "sp"):
        for f in ("crt_batch", "icrt_batch", "ntt_mul_batch", "ring_mul_batch", "matvec"):
            names.add("sr_%s_%s" % (tag, f))
    names.discard("sr_")
    return names


def test_library_exports_every_declared_symbol():
    from stark_rings_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 35
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.sr_elem_limbs(0) == 24 and _lib.lib.sr_elem_limbs(1) == 72 and _lib.lib.sr_elem_limbs(2) == 64
    assert _lib.lib.sr_elem_limbs(9) == 0
    assert b"sm_100a" in _lib.lib.sr_version()


def test_library_carries_sm100a_code_only():
    import subprocess
    so = os.path.join(ROOT, "stark_rings_b200", "libstarkrings_cuda.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import stark_rings_b200 as S
    with pytest.raises(S.StarkRingsError):
        S.Context(0)
    import numpy as np
    with pytest.raises(S.StarkRingsError):
        S.GoldilocksRingConfig.crt_batch(np.zeros(24, dtype=np.uint64))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "stark_rings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                bad = re.findall(r"from oracle|import oracle|c_oracle|ref_py|libsr_oracle|sr_oracle\.|sro_\w+\(", txt)
                assert not bad, (os.path.join(dirpath, f), bad)
