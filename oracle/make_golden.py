"""Generate the committed fixtures in ``tests/golden``.  TEST INFRASTRUCTURE ONLY.

Run in the build container (``python -m oracle.make_golden``): it needs
``/root/reference`` for part (1).

1. ``reference_cases.json`` -- small seeded cases pushed through the reference file
   itself, byte-unmodified (``oracle/run_reference.py``: scipy-backed ``gurobipy``
   stand-in + pandas set-indexer patch).  Inputs and the reference's outputs
   (``predict_cell``, ``step``, clone column, the ``Obj:`` lines it printed) are stored,
   so the GPU box can check against the reference without having it.
2. ``synth_<config>.npz`` -- oracle (``oracle/restatement.py``) outputs for the seeded
   synthetic configs of SURVEY.md section 8d; inputs are regenerated from the seed by
   ``macrodna_b200/synth.py`` at test time.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from macrodna_b200 import synth  # noqa: E402
from oracle import restatement as R  # noqa: E402
from oracle import run_reference as RR  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _frames(rng, n_rna, n_dna, n_genes, ints=False, extra=2, dup_dna=False, const_rna=False, const_dna=False):
    genes = ["g%03d" % i for i in range(n_genes)]
    if ints:
        dna = rng.integers(0, 7, size=(n_genes, n_dna)).astype(np.int64)
        rna = rng.integers(0, 30, size=(n_genes + extra, n_rna)).astype(np.int64)
    else:
        dna = np.log1p(rng.integers(1, 5, size=(n_genes, n_dna)) * (1 + 0.1 * rng.standard_normal((n_genes, n_dna))) ** 2)
        rna = np.log1p(rng.poisson(3.0, size=(n_genes + extra, n_rna)).astype(np.float64))
    if const_rna:
        rna[:, 1] = 2
    if const_dna:
        dna[:, 0] = 2
    dna_ids = ["d%d" % i for i in range(n_dna)]
    if dup_dna:
        half = (n_dna + 1) // 2
        dna[:, half:] = dna[:, : n_dna - half]
        dna_ids = dna_ids[:half] + dna_ids[: n_dna - half]
    rna_genes = genes + ["x%03d" % i for i in range(extra)]
    perm = rng.permutation(len(rna_genes))
    rna_df = pd.DataFrame(rna[perm], index=[rna_genes[i] for i in perm], columns=["r%d" % i for i in range(n_rna)])
    dna_df = pd.DataFrame(dna, index=genes, columns=dna_ids)
    uniq = list(dict.fromkeys(dna_ids))
    lab = pd.DataFrame({"clone": [i % 3 for i in range(len(uniq))], "cell": uniq})
    return rna_df, dna_df, lab


def _case(name, rna_df, dna_df, lab, variant="src"):
    (res, tagged), out = RR.run_reference(rna_df, dna_df, lab, variant=variant, method="cell2cell_assignment")
    clone, _ = RR.run_reference(rna_df, dna_df, lab, variant=variant, method="cell2clone_assignment")
    clone_col = [c for c in clone.columns if c != "predict_cell"][0]
    objs = [float(x) for x in re.findall(r"^Obj: (\S+)$", out, flags=re.M)]
    assert list(res.index) == list(rna_df.columns)
    return {
        "name": name,
        "variant": variant,
        "genes_rna": list(rna_df.index),
        "genes_dna": list(dna_df.index),
        "rna_cells": list(rna_df.columns),
        "dna_cells": list(dna_df.columns),
        "rna": rna_df.to_numpy().tolist(),
        "dna": dna_df.to_numpy().tolist(),
        "int_data": bool(np.issubdtype(rna_df.to_numpy().dtype, np.integer)),
        "label_clone": lab["clone"].tolist(),
        "label_cell": lab["cell"].tolist(),
        "predict_cell": res["predict_cell"].tolist(),
        "tagged_predict_cell": tagged["predict_cell"].tolist(),
        "step": [int(s) for s in tagged["step"].tolist()],
        "clone_column": clone_col,
        "predict_clone": clone[clone_col].tolist(),
        "printed_obj": objs,
    }


def make_reference_cases():
    cases = []
    for variant in ("src", "crc"):
        rna, dna, lab = R.tiny_frames(variant)
        cases.append(_case("tiny_" + variant, rna, dna, lab, variant))
    rng = np.random.default_rng(20231)
    specs = [
        ("m9_n4", dict(n_rna=9, n_dna=4, n_genes=30)),  # M = 2N+1 -> np.squeeze 0-d path (macrodna.py:143)
        ("m3_n7", dict(n_rna=3, n_dna=7, n_genes=30)),  # M < N
        ("m5_n5", dict(n_rna=5, n_dna=5, n_genes=30)),  # square
        ("m8_n4", dict(n_rna=8, n_dna=4, n_genes=30)),  # M mod N == 0
        ("m5_n1", dict(n_rna=5, n_dna=1, n_genes=30)),  # N = 1 -> 5 steps
        ("m1_n1", dict(n_rna=1, n_dna=1, n_genes=12)),
        ("m11_n4_int", dict(n_rna=11, n_dna=4, n_genes=50, ints=True)),
        ("m13_n6_dupdna", dict(n_rna=13, n_dna=6, n_genes=40, dup_dna=True)),
        ("m7_n3_const", dict(n_rna=7, n_dna=3, n_genes=25, const_rna=True, const_dna=True)),
        ("m40_n12", dict(n_rna=40, n_dna=12, n_genes=120)),
        ("m24_n31", dict(n_rna=24, n_dna=31, n_genes=64)),
    ]
    for name, kw in specs:
        rna, dna, lab = _frames(rng, **kw)
        cases.append(_case(name, rna, dna, lab))
    with open(os.path.join(GOLDEN, "reference_cases.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": cases}, f)
    print("reference cases:", [c["name"] for c in cases])


def make_loo_cases():
    """Leave-one-out fixtures from the reference's own class (run_loo_experiment.py, byte-unmodified, shimmed)."""
    import contextlib
    import io

    mod = RR.load_reference_module("loo")
    rng = np.random.default_rng(4711)
    cases = []
    for name, (m, n, g) in (("loo_m9_n4", (9, 4, 30)), ("loo_m6_n7", (6, 7, 24)), ("loo_m12_n4", (12, 4, 40))):
        genes = ["g%03d" % i for i in range(g)]
        rna = pd.DataFrame(np.log1p(rng.poisson(3.0, (g, m)).astype(float)), index=genes, columns=["r%02d" % i for i in range(m)])
        dna = pd.DataFrame(np.log1p(rng.integers(1, 5, (g, n)) * (1 + 0.05 * rng.standard_normal((g, n)))), index=genes,
                           columns=["d%02d" % i for i in range(n)])
        with RR._pandas_set_indexer_patch(), contextlib.redirect_stdout(io.StringIO()):
            obj = mod.MaCroDNA(rna.copy(), dna.copy())
            _, tagged, total, k = obj.cell2cell_assignment()
            runs = []
            for q in range(m):
                t, s_ = obj.leave_one_out(cell_idx=q, K_steps=k, biopsy_name=name)
                runs.append({"rna_cell": list(t.index), "predicted_dna_cell": t["predicted_dna_cell"].tolist(),
                             "step": [x if isinstance(x, str) else int(x) for x in t["step"].tolist()],
                             "corr_val": [float(x) for x in t["corr_val"].tolist()], "objective": float(s_)})
        cases.append({"name": name, "genes": genes, "rna_cells": list(rna.columns), "dna_cells": list(dna.columns),
                      "rna": rna.to_numpy().tolist(), "dna": dna.to_numpy().tolist(), "K": int(k),
                      "full_objective": float(total), "full_corr_val": [float(x) for x in tagged["corr_val"].tolist()],
                      "loo": runs})
    with open(os.path.join(GOLDEN, "loo_cases.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py::make_loo_cases", "cases": cases}, f)
    print("loo cases:", [c["name"] for c in cases])


def make_synth(names=("C2", "C3", "C4")):
    for name in names:
        inst = synth.make_config_arrays(name)
        corrs, assign, step, objs = R.cell2cell_arrays(inst.rna, inst.dna)
        rng = np.random.default_rng(7)
        si = rng.integers(0, corrs.shape[0], size=4096)
        sj = rng.integers(0, corrs.shape[1], size=4096)
        np.savez_compressed(
            os.path.join(GOLDEN, "synth_%s.npz" % name),
            assign=assign,
            step=step,
            objs=objs,
            sample_i=si,
            sample_j=sj,
            sample_corr=corrs[si, sj],
            corr_sum=np.array([corrs.sum()]),
            rna_sum=np.array([inst.rna.sum()]),
            dna_sum=np.array([inst.dna.sum()]),
            clone_acc=np.array([(inst.dna_clone[assign] == inst.rna_clone).mean()]),
        )
        print(name, corrs.shape, "objs", objs, "acc", (inst.dna_clone[assign] == inst.rna_clone).mean())


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    if RR.reference_available():
        make_reference_cases()
        make_loo_cases()
    else:
        print("reference not available: skipping reference_cases.json")
    make_synth()
