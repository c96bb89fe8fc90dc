import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


def load_reference_cases():
    with open(os.path.join(GOLDEN, "reference_cases.json")) as f:
        return json.load(f)["cases"]


def case_frames(case):
    dt = np.int64 if case["int_data"] else np.float64
    rna = pd.DataFrame(np.asarray(case["rna"], dtype=dt), index=case["genes_rna"], columns=case["rna_cells"])
    dna = pd.DataFrame(np.asarray(case["dna"], dtype=dt), index=case["genes_dna"], columns=case["dna_cells"])
    lab = pd.DataFrame({"clone": case["label_clone"], "cell": case["label_cell"]})
    return rna, dna, lab


@pytest.fixture(scope="session")
def reference_cases():
    return load_reference_cases()


@pytest.fixture(scope="session")
def handle():
    from macrodna_b200 import get_handle

    return get_handle(0)


def tie_report(corrs, assign_a, step_a, assign_b, step_b, rel=1e-9):
    """Compare two step-tagged assignments on oracle correlations.

    Returns (identical, report).  Differences are acceptable only as exact ties: every step's
    objective must agree to `rel` relative (SURVEY.md section 8d parity gates).
    """
    identical = bool((assign_a == assign_b).all() and (step_a == step_b).all())
    rep = {"differing_cells": int(((assign_a != assign_b) | (step_a != step_b)).sum()), "steps": []}
    ok = True
    for s in range(1, int(max(step_a.max(), step_b.max())) + 1):
        ia, ib = np.flatnonzero(step_a == s), np.flatnonzero(step_b == s)
        oa, ob = corrs[ia, assign_a[ia]].sum(), corrs[ib, assign_b[ib]].sum()
        gap = abs(oa - ob) / max(1e-300, abs(ob))
        rep["steps"].append((s, float(oa), float(ob), float(gap)))
        if len(ia) != len(ib) or gap > rel:
            ok = False
    rep["objective_ok"] = ok
    return identical, rep
