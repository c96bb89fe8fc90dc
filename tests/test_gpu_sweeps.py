"""GPU tests of the views on the resident correlation matrix: replicate sub-instances, leave-one-out, null test."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, tie_report

pytestmark = pytest.mark.gpu


def _frames(rng, m, n, g):
    genes = ["g%03d" % i for i in range(g)]
    rna = pd.DataFrame(np.log1p(rng.poisson(3.0, (g, m)).astype(float)), index=genes, columns=["r%03d" % i for i in range(m)])
    dna = pd.DataFrame(np.log1p(rng.integers(1, 5, (g, n)) * (1 + 0.05 * rng.standard_normal((g, n)))), index=genes,
                       columns=["d%03d" % i for i in range(n)])
    return rna, dna


def test_subinstance_resampled_dna_and_dropped_cells(handle):
    """A replicate with duplicated / dropped DNA cells and a subset of RNA cells equals a from-scratch run on the
    gathered frames (clonal_proportions_resampling.py:184-190) and the oracle on the gathered matrix."""
    from macrodna_b200 import MaCroDNA
    from oracle import restatement as R

    rng = np.random.default_rng(8)
    rna, dna = _frames(rng, 57, 13, 200)
    m = MaCroDNA(rna.copy(), dna.copy(), variant="resampling")
    base = m.cell2cell_assignment()
    names = list(rng.choice(dna.columns, size=9, replace=True))
    keep_rna = [c for k, c in enumerate(rna.columns) if k % 5 != 2]
    sub = m.subinstance_assignment(rna_cells=keep_rna, dna_cells=names)
    corrs = R.correlation_matrix(rna.T.to_numpy(), dna.T.to_numpy())
    rows = [list(rna.columns).index(c) for c in keep_rna]
    cols = [list(dna.columns).index(c) for c in names]
    a_ref, s_ref, o_ref = R.step_loop(corrs[np.ix_(rows, cols)])
    assert np.allclose(m.last_sub["objs"], o_ref, rtol=1e-12, atol=1e-13)
    ident, rep = tie_report(corrs[np.ix_(rows, cols)], m.last_sub["assign"], m.last_sub["step"], a_ref, s_ref, 1e-12)
    assert rep["objective_ok"], rep  # duplicated DNA columns are exact ties: objectives must agree step by step
    assert sub["rna_cell"].tolist() == keep_rna and set(sub["predicted_dna_cell"]) <= set(names)
    # the resident matrix survived: the base run is still what the handle answers for
    again = m.subinstance_assignment()
    assert again.equals(base)
    fresh = MaCroDNA(rna[keep_rna].copy(), dna.loc[:, names].copy(), variant="objective").cell2cell_assignment()
    assert abs(fresh[2] - o_ref.sum()) <= 1e-12 * abs(o_ref.sum())


def test_leave_one_out_matches_reference_fixtures(handle):
    from macrodna_b200 import MaCroDNA

    with open(os.path.join(GOLDEN, "loo_cases.json")) as f:
        cases = json.load(f)["cases"]
    for case in cases:
        rna = pd.DataFrame(np.asarray(case["rna"]), index=case["genes"], columns=case["rna_cells"])
        dna = pd.DataFrame(np.asarray(case["dna"]), index=case["genes"], columns=case["dna_cells"])
        m = MaCroDNA(rna, dna, variant="loo")
        _, tagged, total, k = m.cell2cell_assignment()
        assert k == case["K"] and abs(total - case["full_objective"]) < 1e-11
        for q, ref in enumerate(case["loo"]):
            t, s = m.leave_one_out(cell_idx=q, K_steps=k, biopsy_name=case["name"])
            assert list(t.index) == ref["rna_cell"]
            assert t["predicted_dna_cell"].tolist() == ref["predicted_dna_cell"]
            assert t["step"].tolist() == ref["step"]
            assert np.allclose(t["corr_val"].to_numpy(dtype=float), ref["corr_val"], rtol=0, atol=1e-11)
            assert abs(s - ref["objective"]) < 1e-11


@pytest.mark.parametrize("mn", [(11, 4), (6, 9), (40, 40), (230, 57)])
def test_null_assignments_distribution(handle, mn):
    """Every RNA cell meets a uniformly random DNA cell, injectively within a step: the mean of the statistic is
    exactly sum_i mean_j C[i, j]; the spread must match the oracle's draw-by-draw restatement of the reference."""
    from macrodna_b200 import random_test
    from oracle import restatement as R

    M, N = mn
    rng = np.random.default_rng(M + N)
    rna, dna = _frames(rng, M, N, 150)
    rt = random_test(rna, dna)
    assert rt.n_iters == R.n_steps(M, N)
    sums, med = rt.assign_many(20000, seed=5, medians=True)
    corrs = R.correlation_matrix(rna.T.to_numpy(), dna.T.to_numpy())
    expect = corrs.mean(axis=1).sum()
    assert abs(sums.mean() - expect) < 4.5 * sums.std() / np.sqrt(len(sums))
    ref = np.array([R.random_assign(corrs, rng) for _ in range(4000)])
    assert abs(sums.std() - ref.std()) < 0.08 * ref.std()
    assert np.quantile(sums, 0.999) <= R.step_loop(corrs)[2].sum() + 1e-9  # nothing beats the optimum
    assert (med >= corrs.min() - 1e-12).all() and (med <= corrs.max() + 1e-12).all()
    again = rt.assign_many(20000, seed=5)
    assert (again == sums).all()  # deterministic in (seed, trial)
    assert isinstance(rt.assign(), float)


def test_concurrent_sweep_equals_one_at_a_time(handle):
    """mcd_subinstance_sweep keeps several replicates in flight on worker streams (each on a slice of the chip); every
    replicate must come out as the one-at-a-time call gives it: same objectives always, same assignments on tie-free
    replicates (DNA cells dropped / permuted); resampled replicates (duplicated DNA cells = exact ties) are optimal
    step by step (certificate) with equal objectives."""
    from macrodna_b200 import dist, synth
    from oracle import restatement as R

    inst = synth.make_arrays(700, 120, 500, 4, seed=17)
    M, N, G = 700, 120, 500
    handle.cell2cell(inst.rna, inst.dna, M, N, G)
    rng = np.random.default_rng(3)
    subsets = np.stack([rng.permutation(N)[:100] for _ in range(9)]).astype(np.int32)      # tie-free
    resampled = np.stack([synth.resample_dna_columns(inst.dna_clone, seed=s) for s in range(9)]).astype(np.int32)
    for cols, tie_free in ((subsets, True), (resampled, False)):
        a, s, o, gaps, st = handle.subinstance_sweep(cols, M=M, concurrency=4)
        assert (gaps >= 0).all() and gaps.max() <= 1e-12 and st.as_dict()["cert_bad"] == 0
        for r in range(cols.shape[0]):
            a1, s1, o1, _ = handle.subinstance(None, cols[r], M=M, N=N)
            assert np.allclose(o[r], o1, rtol=1e-12)
            if tie_free:
                assert (a[r] == a1).all() and (s[r] == s1).all()
            assert np.bincount(s[r])[1:].sum() == M
    # against the oracle, and the accuracy of the reference's sweep (clonal_proportions_resampling.py:191-201)
    c_ref = R.correlation_matrix(inst.rna, inst.dna)
    a, s, o, gaps, _ = handle.subinstance_sweep(subsets[:3], M=M, concurrency=3)
    for r in range(3):
        a_ref, s_ref, o_ref = R.step_loop(c_ref[:, subsets[r]])
        assert (a[r] == a_ref).all() and (s[r] == s_ref).all() and np.allclose(o[r], o_ref, rtol=1e-12)
        acc = dist.replicate_accuracy(a[r], subsets[r], inst.rna_clone, inst.dna_clone)
        assert acc == np.mean(inst.dna_clone[subsets[r][a_ref]] == inst.rna_clone) and acc > 0.5
    # the multi-rank driver: ranks own replicates r mod P, results identical to the single-rank sweep
    whole = dist.sweep_assignments(handle, inst.rna, inst.dna, list(subsets), world=1, rank=0, concurrency=4)
    part = {}
    for rank in range(2):
        part.update(dist.sweep_assignments(handle, inst.rna, inst.dna, list(subsets), world=2, rank=rank, concurrency=2))
    assert sorted(part) == sorted(whole) == list(range(9))
    for r in whole:
        assert (whole[r][0] == part[r][0]).all() and (whole[r][1] == part[r][1]).all()


def test_resampled_replicates_with_duplicated_cells(handle):
    """Resampling WITH replacement (clonal_proportions_resampling.py:184-187) duplicates DNA cells: exact ties by
    construction.  Round 1 spun to the round guard on them (264 000 rounds, seconds per replicate).  The copies of a
    cell now bid as a class of similar persons; every step must be optimal (SciPy on the oracle matrix, objective
    within 1e-12), carry a certificate at rounding level, and need about as many rounds as a tie-free instance."""
    from macrodna_b200 import MaCroDNA, synth
    from oracle import restatement as R
    from scipy.optimize import linear_sum_assignment

    inst = synth.make_arrays(1500, 300, 800, 4, seed=31)
    M, N, G = 1500, 300, 800
    _, _, _, base = handle.cell2cell(inst.rna, inst.dna, M, N, G)
    base_rounds = base.as_dict()["lap_rounds"]
    c_ref = R.correlation_matrix(inst.rna, inst.dna)
    for seed in range(3):
        cols = synth.resample_dna_columns(inst.dna_clone, seed=seed).astype(np.int32)
        assert len(set(cols.tolist())) < 0.8 * N  # plenty of duplicates
        a, s, o, st = handle.subinstance(None, cols, M=M, N=N)
        d = st.as_dict()
        assert d["lap_rounds"] < 6 * base_rounds, (d["lap_rounds"], base_rounds)
        assert 0.0 <= d["cert_rel_gap"] <= 1e-12 and d["cert_bad"] == 0
        sub = c_ref[:, cols]
        remaining = np.arange(M)
        for k in range(5):
            rows = np.flatnonzero(s == k + 1)
            r, c = linear_sum_assignment(sub[remaining], maximize=True)
            best = sub[remaining][r, c].sum()
            mine = sub[rows, a[rows]].sum()
            assert abs(mine - best) <= 1e-12 * abs(best), (seed, k, mine, best)
            assert abs(o[k] - mine) <= 1e-12 * abs(mine)
            assert len(set(a[rows].tolist())) == len(rows)         # injective within the step
            remaining = np.setdiff1d(remaining, rows)
    # the drop-in class on a resampled FRAME (duplicate column labels, identical data): same machinery
    rna_df, dna_df, lab = synth.make_frames(inst)
    names = [dna_df.columns[j] for j in cols]
    m = MaCroDNA(rna_df, dna_df.loc[:, names], variant="resampling")
    frame = m.cell2cell_assignment()
    assert list(frame.columns) == ["predicted_dna_cell", "rna_cell", "step"] and len(frame) == M
    assert m.last_stats["cert_rel_gap"] <= 1e-12 and m.last_stats["lap_rounds"] < 6 * base_rounds
    assert np.allclose(m.last_objective, o, rtol=1e-12)
