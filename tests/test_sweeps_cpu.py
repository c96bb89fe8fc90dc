"""CPU tests of the rows either side of the hot path: the oracle's leave-one-out and
null-test restatements against the reference's own classes (golden fixtures / live reference when present),
and the CSV / normaliser helpers against the reference scripts' formulas."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN
from oracle import restatement as R


def _loo_cases():
    with open(os.path.join(GOLDEN, "loo_cases.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _loo_cases(), ids=lambda c: c["name"])
def test_oracle_leave_one_out_matches_reference_run(case):
    rna = np.asarray(case["rna"], dtype=np.float64).T  # cells x genes
    dna = np.asarray(case["dna"], dtype=np.float64).T
    corrs = R.correlation_matrix(rna, dna)
    a, s, o = R.step_loop(corrs)
    assert abs(float(o.sum()) - case["full_objective"]) < 1e-12
    assert np.allclose(corrs[np.arange(len(a)), a], case["full_corr_val"], rtol=0, atol=1e-12)
    for q, ref in enumerate(case["loo"]):
        assign, step, objs, tdna, tval, total = R.leave_one_out(corrs, q, case["K"])
        rest = [c for k, c in enumerate(case["rna_cells"]) if k != q]
        assert ref["rna_cell"] == rest + [case["rna_cells"][q]]
        assert ref["predicted_dna_cell"] == [case["dna_cells"][j] for j in assign] + [case["dna_cells"][tdna]]
        assert ref["step"] == [int(x) for x in step] + ["TEST"]
        assert abs(total - ref["objective"]) < 1e-12
        assert abs(tval - ref["corr_val"][-1]) < 1e-12


def test_oracle_random_assign_matches_reference_distribution():
    """Same statistic as the reference's `random_test.assign` (when /root/reference is present), and the exact
    expectation sum_i mean_j C[i, j] in any case: every RNA cell meets a uniformly random DNA cell."""
    rng = np.random.default_rng(1)
    corrs = rng.standard_normal((11, 4)) * 0.1
    draws = np.array([R.random_assign(corrs, rng) for _ in range(20000)])
    expect = corrs.mean(axis=1).sum()
    assert abs(draws.mean() - expect) < 4 * draws.std() / np.sqrt(len(draws))
    from oracle import run_reference as RR

    if not RR.reference_available():
        return
    mod = RR.load_reference_module("random")
    obj = mod.random_test.__new__(mod.random_test)  # the constructor only builds corrs / n_iters from frames
    obj.corrs = corrs
    obj.rna_np = np.zeros((11, 1))
    obj.dna_np = np.zeros((4, 1))
    obj.quotient, obj.remainder = divmod(11, 4)
    obj.n_iters = 3
    np.random.seed(2023)
    ref = np.array([obj.assign() for _ in range(20000)])
    assert abs(ref.mean() - draws.mean()) < 5 * np.hypot(ref.std(), draws.std()) / np.sqrt(20000)
    assert abs(ref.std() - draws.std()) < 0.05 * ref.std()


def test_normalisers_follow_reference_scripts():
    from macrodna_b200 import io as mio

    rng = np.random.default_rng(3)
    counts = pd.DataFrame(rng.poisson(40, (120, 6)).astype(float), index=["g%d__chr1" % i for i in range(120)],
                          columns=["c%d" % i for i in range(6)])
    counts["c5"] = 1.0  # coverage 120 <= 3000: dropped
    counts.iloc[3, 0] = np.nan
    d = mio.normalize_dna_counts(counts)
    assert list(d.columns) == ["c0", "c1", "c2", "c3", "c4"]
    x = counts.fillna(0)[d.columns] + 1
    assert np.allclose(d.to_numpy(), np.log1p(2 * x / x.median()).to_numpy())
    counts.iloc[7, :] = 2.0  # never reaches 3 transcripts: gene dropped
    r = mio.normalize_rna_counts(counts)
    assert "g7" not in r.index and r.index[0] == "g0" and list(r.columns) == list(d.columns)
    y = counts.fillna(0)[r.columns].drop(index="g7__chr1") + 1
    assert np.allclose(r.to_numpy(), np.log1p(y / y.sum() * 1e6).to_numpy())
    assert (r.to_numpy() >= 0).all()


def test_indexed_csv_roundtrip(tmp_path):
    from macrodna_b200 import io as mio

    tagged = pd.DataFrame({"predict_cell": ["d1", "d0", "d1"], "cell": ["r0", "r1", "r2"], "step": [1, 1, 2]}).set_index("cell")
    p = tmp_path / "PAT_cell2cell_assignment_indexed.csv"
    mio.write_indexed_csv(tagged, p)
    assert p.read_text().splitlines()[0] == "cell,predict_cell,step"
    back = mio.read_indexed_csv(p)
    assert back.equals(tagged)
    with pytest.raises(ValueError):
        mio.write_indexed_csv(tagged.reset_index(), p)
