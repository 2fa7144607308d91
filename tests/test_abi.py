"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/morsi_cuda.h declares, the host logic (element grammar,
operation table, CLI error paths) matches the reference, and compute calls
fail loudly without a GPU (no CPU fallback).  No compute on the GPU here."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

import imscript_b200 as M
from imscript_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "imscript_b200", "lib", "morsi")


def test_library_exports_every_declared_symbol():
    L = M.lib()
    header = open(os.path.join(ROOT, "include", "morsi_cuda.h")).read()
    declared = set(re.findall(r"\b(morsi_[a-zA-Z0-9_]+)\s*\(", header))
    declared -= {"morsi_" + o for o in M.OPS} | {"morsi_all"}     # those live in libmorsi_compat
    assert declared == set(binding.EXPORTED_SYMBOLS)
    for name in binding.EXPORTED_SYMBOLS:
        assert hasattr(L, name), name
    C = ctypes.CDLL(M.lib_path("libmorsi_compat.so"))
    for name in binding.COMPAT_SYMBOLS:
        assert hasattr(C, name), name


def test_elements_match_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "elements.json")))
    for name, e in gold.items():
        got = M.parse_element(name)
        if e is None:
            assert got is None, name
        else:
            assert got is not None and list(got) == e, name
    assert M.parse_element("square3") is None
    assert list(M.parse_element("kids2.5")) == gold["disk2.5"]
    assert list(M.parse_element("rrrr3")) == list(M.parse_element("Drec3"))
    assert M.build_element("disk", 1.0) is None
    assert M.build_element("disk", float("nan")) is None


def test_operation_table():
    for i, op in enumerate(M.OPS):
        assert M.parse_operation(op) == i
        assert M.lib().morsi_operation_name(i).decode() == op
    assert M.parse_operation("top-hat") == -1 and M.parse_operation("") == -1


def test_element_classification():
    assert M.describe_element("cross").startswith("small3x3 n=5")
    assert M.describe_element("square").startswith("small3x3 n=9")
    assert M.describe_element("disk7").startswith("rowrun n=145")
    assert M.describe_element("disk15").startswith("rowrun n=697 box=[-14,14]x[-14,14]")
    assert M.describe_element("hrec5").startswith("rowrun")
    assert M.describe_element("vrec5").startswith("rowrun")
    assert M.describe_element("dysk5").startswith("direct")
    assert M.describe_element("drec5").startswith("direct")
    assert "dup" in M.describe_element([3, 0, 0, 0, 0, 0, 0, 0, 5, 5])


def test_halo_rows():
    assert M.halo_rows("erosion", "disk15") == (14, 14)
    assert M.halo_rows("tophat", "disk15") == (28, 28)
    assert M.halo_rows("gradient", "cross") == (1, 1)
    assert M.halo_rows("oscillation", "hrec9") == (0, 0)
    assert M.halo_rows("opening", [2, 0, 0, 0, 0, 3, 0, 1]) == (0, 6)


def test_invalid_arguments_are_rejected_before_any_device_work():
    x = np.zeros((4, 4), np.float32)
    with pytest.raises(M.MorsiError):
        M.apply(99, "cross", x)
    with pytest.raises(M.MorsiError):
        M.apply("erosion", "square3", x)
    rc = M.lib().morsi_cuda_apply(0, None, x.ctypes.data, x.ctypes.data, 4, 4, 1)
    assert rc == 1


@pytest.mark.skipif(M.device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    with pytest.raises(M.MorsiError) as ei:
        M.apply("erosion", "cross", np.zeros((4, 4), np.float32))
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_synth_host_is_deterministic_and_band_consistent():
    a = M.synth_host(33, 20, row0=0, plane=2, seed=5)
    b = M.synth_host(33, 7, row0=9, plane=2, seed=5)
    assert np.array_equal(a[9:16], b)
    assert a.min() >= 0 and a.max() < 1 and len(np.unique(a)) > 600
    c = M.synth_host(64, 64, seed=3, dist=2)
    assert np.isnan(c).any() and np.isinf(c).any() and (c == 0).any()
    d = M.synth_host(16, 16, seed=3, dist=1)
    assert np.array_equal(d, np.round(d)) and d.max() <= 255


@pytest.mark.skipif(not os.path.exists(CLI), reason="morsi CLI not built")
def test_cli_error_paths_match_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "cli.json")))
    for key, g in gold.items():
        argv = key.split() if key else []
        if g["rc"] == 0 and "IN" in argv:
            continue                                   # needs the GPU: tests/test_gpu_cli.py
        argv = ["/nonexistent.npy" if a == "IN" else a for a in argv]
        p = subprocess.run([CLI] + argv, capture_output=True)
        assert p.returncode == g["rc"], key
        assert p.stderr.decode().replace(CLI, "morsi") == g["stderr"], key
        if "stdout" in g:
            assert p.stdout.decode() == g["stdout"], key


def test_long_line_with_repeats_is_flagged():
    """a one-row list of 100 offsets spanning 100 columns, one column repeated and one missing: the line
    kernels must not take it for a contiguous run (round-1 advice)"""
    offs = list(range(100))
    offs[50] = 10                                   # column 50 missing, column 10 twice
    e = np.array([100, 0, 0, 0] + [v for x in offs for v in (x, 0)], dtype=np.int32)
    assert " dup" in M.describe_element(e)
    clean = np.array([100, 0, 0, 0] + [v for x in range(100) for v in (x, 0)], dtype=np.int32)
    assert " dup" not in M.describe_element(clean)
