"""CPU checks of host-checkable logic that the CUDA kernels are built from: the compile-time MMA schedule of K2c and
the bit tricks of K1's digit fast path.  The code under test is cut out of the .cu sources and compiled with g++
against small harnesses in tests/host/ (no GPU, no oracle)."""
import os
import shutil
import subprocess

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
