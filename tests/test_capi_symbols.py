"""CPU: the C-ABI library loads here (no GPU) and exports every symbol include/dlb200.h declares; host-only helpers
work; anything that computes fails loudly without a device instead of falling back to the CPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import deeplearning_b200 as dl
from deeplearning_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dlb200.h")).read()
    names = set()
    for name, templ in re.findall(r"DLB_API\s+[\w\s\*]+?\b(dlb_\w+?)(_##SFX)?\s*\(", text):
        names.update({name + "_f32", name + "_f64"} if templ else {name})
    return names


def test_header_and_binding_agree():
    assert declared_symbols() == set(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "deeplearning_b200", "libdlb200.so")
    assert os.path.exists(lib_path), "build it: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = declared_symbols() - exported
    assert not missing, f"not exported: {sorted(missing)}"
    stray = {s for s in exported if not s.startswith("dlb_") and not s.startswith("_")}
    assert not stray, f"unexpected exported symbols: {sorted(stray)[:10]}"


def test_sass_is_sm_100a():
    lib_path = os.path.join(ROOT, "deeplearning_b200", "libdlb200.so")
    out = subprocess.check_output(["cuobjdump", "-lelf", lib_path], text=True)
    assert "sm_100a" in out, out


def test_host_only_helpers():
    lib = _capi.load()
    assert lib.dlb_version() >= 100
    p = dl.shuffle_sequence(3, 50, 2)
    assert p.shape == (2, 50) and sorted(p[0]) == list(range(50)) and sorted(p[1]) == list(range(50))
    assert not np.array_equal(p[0], p[1])
    assert np.array_equal(p, dl.shuffle_sequence(3, 50, 2))
    # data-parallel partitioning rule: contiguous, complete, ceil-sized shares
    for count in (0, 1, 7, 256, 8192, 8195):
        for world in (1, 2, 3, 8):
            parts = [dl.rank_slice(count, r, world) for r in range(world)]
            assert sum(m for _, m in parts) == count
            pos = 0
            for first, mine in parts:
                assert first == min(pos, count) and mine >= 0
                pos += mine
    with pytest.raises(dl.DlbError):
        dl.rank_slice(10, 4, 4)


def test_no_cpu_fallback_without_a_device():
    """Where there is no CUDA device the library refuses to create a context (and says why)."""
    from tests.conftest import _cuda_device_present
    if _cuda_device_present():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dl.DlbError) as e:
        dl.Context(0)
    assert str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "deeplearning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "dl_oracle" not in text and "liboracle" not in text, f
