"""CPU checks of host-checkable logic that the CUDA kernels are built from: the compile-time MMA schedule of K2c and
the bit tricks of K1's digit fast path.  The code under test is cut out of the .cu sources and compiled with g++
against small harnesses in tests/host/ (no GPU, no oracle)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "macrodna_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def _cut(path, start, end):
    src = open(path).read()
    a, b = src.index(start), src.index(end)
    assert a < b
    return src[a:b]


def _build_and_run(tmp_path, harness, macro, snippet):
    snip = tmp_path / "snippet.h"
    snip.write_text(snippet)
    exe = tmp_path / "check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-D%s=\"%s\"" % (macro, snip), os.path.join(HOST, harness), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_k2c_pass_plan_is_a_valid_schedule(tmp_path):
    snippet = _cut(os.path.join(CSRC, "corr_ozaki.cu"), "struct PassPlan {",
                   "template <int NLOAD, int NG, int MODE>\n__host__ __device__ constexpr uint32_t packed_load_order")
    assert "pass plan ok" in _build_and_run(tmp_path, "pass_plan_check.cpp", "PLAN_SNIPPET", snippet)


def test_k1_digit_fast_path_bits_match_reference_digits(tmp_path):
    src = os.path.join(CSRC, "standardize.cu")
    snippet = (_cut(src, "__device__ __forceinline__ void ozaki_digits", "// Split-precision operand") +
               _cut(src, "template <int NSL>\n__device__ __forceinline__ void ozaki_bits", "// T threads per row, NV4 sweeps"))
    assert "0 mismatches" in _build_and_run(tmp_path, "digit_bits_check.cpp", "DIGIT_SNIPPET", snippet)


def test_fork_guard_refuses_an_inherited_context(monkeypatch):
    """The reference's sweep scripts use the class in the parent and then fork a Pool
    (clonal_proportions_resampling.py:266-267, :300): a handle created by another pid must never be touched."""
    import os

    from macrodna_b200 import api

    class FakeHandle:
        h = 1
        pid = os.getpid() + 12345  # "created by the parent"

    monkeypatch.setitem(api._HANDLES, 0, FakeHandle())
    with pytest.raises(RuntimeError, match="fork"):
        api.get_handle(0)
    monkeypatch.delitem(api._HANDLES, 0)


def test_duplicate_dna_cells_are_detected_only_with_identical_data():
    from macrodna_b200.api import MaCroDNA

    rng = np.random.default_rng(0)
    base = rng.random((4, 9))
    cells = ["a", "b", "a", "c", "b", "a"]
    data = np.stack([base[{"a": 0, "b": 1, "c": 2}[c]] for c in cells])
    uniq, cols = MaCroDNA._duplicate_dna_cells(cells, data)
    assert uniq.tolist() == [0, 1, 3] and cols.tolist() == [0, 1, 0, 2, 1, 0]
    assert MaCroDNA._duplicate_dna_cells(["a", "b", "c"], data[:3]) is None
    data2 = data.copy()
    data2[2, 4] += 1e-9  # same label, different data: genuinely different cells
    assert MaCroDNA._duplicate_dna_cells(cells, data2) is None


def test_matrix_hand_off_validation():
    from macrodna_b200 import _lib

    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    out = _lib._matrix(a, 3, 4, "rna")
    assert out.dtype == np.float64 and out.flags.c_contiguous and (out == a).all()
    f = np.asfortranarray(np.arange(12.0).reshape(3, 4))
    assert _lib._matrix(f, 3, 4, "rna").flags.c_contiguous
    with pytest.raises(ValueError, match="shape"):
        _lib._matrix(a, 4, 3, "rna")
    assert _lib._matrix(12345, 3, 4, "rna") == 12345  # device pointers pass through
    with pytest.raises(ValueError, match="assign"):
        _lib._out(np.empty(3, dtype=np.int64), 3, np.int32, "assign")
