"""CPU tests: the C-ABI library builds/loads and exports every symbol include/*.h declares; host-side
validation of the drop-in class fails loudly (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT
from macrodna_b200 import MaCroDNA, _lib, build


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "macrodna_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", txt)))


def test_library_is_built_in_tree():
    build.build()
    assert os.path.exists(_lib.LIB_PATH)
    assert os.path.dirname(_lib.LIB_PATH).endswith("macrodna_b200")


def test_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)


def test_every_handle_option_is_documented_in_the_header():
    """mcd_set_option's table (csrc/api.cu) and the option list in the header's comment must not drift apart."""
    src = open(os.path.join(ROOT, "macrodna_b200", "csrc", "api.cu")).read()
    names = re.findall(r'MCD_OPT_[ID]\("([a-z0-9_.]+)"', src)
    assert len(names) >= 30 and len(set(names)) == len(names)
    hdr = open(os.path.join(ROOT, "include", "macrodna_b200.h")).read()
    missing = [n for n in names if '"%s"' % n not in hdr]
    assert not missing, missing


def test_pure_host_entry_points():
    lib = _lib.load_library()
    assert lib.mcd_abi_version() == 2
    assert lib.mcd_padded_k(20000) == 20000 and lib.mcd_padded_k(20001) == 20016 and lib.mcd_padded_k(6) == 16
    assert lib.mcd_padded_k_split(6) == 64
    assert [lib.mcd_num_steps(m, n) for m, n in [(9, 4), (3, 7), (5, 5), (8, 4), (5, 1)]] == [3, 1, 1, 2, 5]
    assert lib.mcd_strerror(0) == b"ok" and b"non-finite" in lib.mcd_strerror(-4)
    assert ctypes.sizeof(_lib.McdStats) == 8 * (13 + 8 + 3 * 64 + 4 + 64 + 1)


def _frames(m=5, n=3, g=8):
    rng = np.random.default_rng(0)
    genes = ["g%d" % i for i in range(g)]
    rna = pd.DataFrame(rng.random((g, m)), index=genes, columns=["r%d" % i for i in range(m)])
    dna = pd.DataFrame(rng.random((g, n)), index=genes, columns=["d%d" % i for i in range(n)])
    return rna, dna


def test_host_validation_errors():
    rna, dna = _frames()
    bad = rna.copy()
    bad.columns = ["r0", "r0", "r2", "r3", "r4"]
    with pytest.raises(ValueError, match="duplicate RNA"):
        MaCroDNA(bad, dna).cell2cell_assignment()
    other = dna.copy()
    other.index = ["z%d" % i for i in range(len(dna))]
    with pytest.raises(ValueError, match="share no genes"):
        MaCroDNA(rna, other).cell2cell_assignment()
    txt = rna.astype(object)
    txt.iloc[0, 0] = "abc"
    with pytest.raises(ValueError, match="non-numeric"):
        MaCroDNA(txt, dna).cell2cell_assignment()
    with pytest.raises(ValueError, match="dna_label"):
        MaCroDNA(rna, dna).cell2clone_assignment()


def test_no_cpu_fallback_without_gpu():
    """Without a GPU the product path must raise, never compute on the host."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    rna, dna = _frames()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MaCroDNA(rna, dna).cell2cell_assignment()


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "macrodna_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "scipy" not in src or f == "synth.py", f


def test_instances_pickle_before_use():
    import pickle

    rna, dna = _frames()
    m = pickle.loads(pickle.dumps(MaCroDNA(rna, dna)))
    assert m.rna_df.equals(rna)
