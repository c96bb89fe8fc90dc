"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed
golden fixtures.  Tolerances: correlations <= 1e-6 abs (north star; FP64 path is checked at
1e-12), per-step objective <= 1e-9 relative, assignments bit-identical except exact ties
(reported, objective-equal)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_frames, load_reference_cases, tie_report

pytestmark = pytest.mark.gpu

CASES = load_reference_cases()


def _torch():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def _dev(a):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------
# the drop-in class against the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_class_matches_reference_run(case):
    from macrodna_b200 import MaCroDNA
    from oracle import restatement as R

    rna, dna, lab = case_frames(case)
    m = MaCroDNA(rna.copy(), dna.copy(), lab.copy(), clone_column=case["clone_column"])
    res, tagged = m.cell2cell_assignment()
    assert list(res.index) == case["rna_cells"] and res.index.name == "cell"
    assert list(res.columns) == ["predict_cell"] and list(tagged.columns) == ["predict_cell", "step"]
    assert len(m.last_objective) == len(case["printed_obj"])
    for a, b in zip(m.last_objective, case["printed_obj"]):
        assert abs(a - b) <= 5e-6 * max(1.0, abs(b))
    # the side effect of macrodna.py:90-91: frames are gene-filtered in place
    assert len(m.rna_df.index) == len(m.dna_df.index) == len(set(case["genes_rna"]) & set(case["genes_dna"]))
    o = R.OracleMaCroDNA(rna.copy(), dna.copy(), lab.copy())
    o.cell2cell_assignment()
    assert np.allclose(m.last_objective, o.last["objs"], rtol=1e-12, atol=1e-13)
    if "dup" in case["name"] or "const" in case["name"]:
        ident, rep = tie_report(o.last["corrs"], m.last_assign, m.last_step, o.last["assign"], o.last["step"], 1e-12)
        assert rep["objective_ok"], rep
        return
    assert res["predict_cell"].tolist() == case["predict_cell"]
    assert tagged["step"].tolist() == case["step"]
    m2 = MaCroDNA(rna.copy(), dna.copy(), lab.copy(), clone_column=case["clone_column"])
    clone = m2.cell2clone_assignment()
    assert list(clone.columns) == ["predict_cell", case["clone_column"]]
    assert clone[case["clone_column"]].tolist() == case["predict_clone"]


def test_tiny_test_known_answer(capsys):
    from macrodna_b200 import MaCroDNA

    m = MaCroDNA(verbose=True)
    out = m.tiny_test()
    assert abs(m.last_objective[0] - 2.8300077180864673) < 1e-13  # README.md:138 of the reference
    assert out["predict_cell"].tolist() == ["cell1", "cell2", "cell3", "cell4"]
    assert out["predict_clone"].tolist() == [0, 1, 2, 3]
    txt = capsys.readouterr().out
    assert "MaCroDNA will be run for 1 steps" in txt and "Obj: 2.83001" in txt and "Test Success" in txt


def test_nan_input_raises():
    from macrodna_b200 import MaCroDNA

    rna, dna, lab = case_frames(CASES[2])
    rna = rna.astype(float)
    rna.iloc[3, 2] = np.nan
    with pytest.raises(ValueError, match="NaN|non-finite"):
        MaCroDNA(rna, dna).cell2cell_assignment()


# ------------------------------------------------------------------------------------------------
# kernels, one by one, through the C ABI with device pointers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [1, 6, 255, 256, 257, 1000, 2000, 4097, 8192, 10000, 15000, 20000, 24576, 30001])
@pytest.mark.parametrize("odd_ld", [False, True])
def test_standardize_kernel(handle, G, odd_ld):
    torch = _torch()
    from oracle import restatement as R

    rng = np.random.default_rng(G)
    ncells = 37
    ld = G + (3 if odd_ld else 0)
    buf = np.full((ncells, ld), 7.25)
    x = np.log1p(rng.poisson(4.0, size=(ncells, G)).astype(np.float64)) + rng.random((ncells, G))
    x[5] = 3.0  # constant cell
    buf[:, :G] = x
    d_x = _dev(buf)
    ldk = handle.lib.mcd_padded_k(G)
    d_y = torch.full((ncells, ldk), -1.0, dtype=torch.float64, device="cuda")
    d_n = torch.empty(ncells, dtype=torch.float64, device="cuda")
    handle.check(handle.lib.mcd_standardize(handle.h, d_x.data_ptr(), ncells, G, ld, d_y.data_ptr(), d_n.data_ptr()))
    handle.check(handle.lib.mcd_check_finite(handle.h))
    xc, nrm = R.standardise(x)
    y = d_y.cpu().numpy()
    assert np.abs(y[:, :G] - xc).max() <= 1e-13 * max(1.0, np.abs(x).max())
    assert (y[:, G:] == 0).all()
    assert np.abs(d_n.cpu().numpy() - nrm).max() <= 1e-12 * max(1.0, nrm.max())
    assert d_n.cpu().numpy()[5] <= 1e-12 * np.sqrt(G) * 3.0  # zero-variance cell


@pytest.mark.parametrize("shape", [(4, 4, 6), (130, 129, 100), (300, 257, 1000), (129, 260, 2001), (1000, 64, 333)])
def test_corr_fp64_kernel(handle, shape):
    torch = _torch()
    from oracle import restatement as R

    M, N, G = shape
    rng = np.random.default_rng(M * 7 + N)
    rna = np.log1p(rng.poisson(4.0, size=(M, G)).astype(np.float64))
    dna = np.log1p(rng.integers(1, 5, size=(N, G)) * (1 + 0.05 * rng.standard_normal((N, G))))
    if N > 2:
        dna[1] = 2.0
    ldk = handle.lib.mcd_padded_k(G)
    d_a = torch.empty((M, ldk), dtype=torch.float64, device="cuda")
    d_b = torch.empty((N, ldk), dtype=torch.float64, device="cuda")
    d_na = torch.empty(M, dtype=torch.float64, device="cuda")
    d_nb = torch.empty(N, dtype=torch.float64, device="cuda")
    lib, h = handle.lib, handle.h
    handle.check(lib.mcd_standardize(h, _dev(rna).data_ptr(), M, G, G, d_a.data_ptr(), d_na.data_ptr()))
    handle.check(lib.mcd_standardize(h, _dev(dna).data_ptr(), N, G, G, d_b.data_ptr(), d_nb.data_ptr()))
    ldc, ldct = N + 3, M + 1
    d_c = torch.full((M, ldc), 9.0, dtype=torch.float64, device="cuda")
    d_ct = torch.full((N, ldct), 9.0, dtype=torch.float64, device="cuda")
    handle.check(lib.mcd_corr_fp64(h, d_a.data_ptr(), M, d_b.data_ptr(), N, G, ldk, d_na.data_ptr(), d_nb.data_ptr(),
                                   d_c.data_ptr(), ldc, d_ct.data_ptr(), ldct))
    handle.synchronize()
    c = d_c.cpu().numpy()
    ct = d_ct.cpu().numpy()
    ref = R.correlation_matrix(rna, dna)
    assert np.abs(c[:, :N] - ref).max() < 1e-12
    assert (c[:, N:] == 9.0).all() and (ct[:, M:] == 9.0).all()  # padding untouched
    assert (ct[:, :M] == c[:, :N].T).all()  # transpose is bit-identical
    if N > 2:
        assert (c[:, 1] == 0).all()  # zero-variance cell gives exactly 0.0 (macrodna.py:25)


def _lap_gpu(handle, w):
    torch = _torch()
    n, m = w.shape
    d_w = _dev(w)
    d_col = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    d_obj = torch.zeros(1, dtype=torch.float64, device="cuda")
    handle.check(handle.lib.mcd_lap_max(handle.h, d_w.data_ptr(), n, m, m, d_col.data_ptr(), d_obj.data_ptr()))
    handle.synchronize()
    return d_col.cpu().numpy(), float(d_obj.cpu().numpy()[0])


@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (2, 2), (5, 5), (4, 9), (50, 50), (100, 300), (257, 1000),
                                   (300, 301), (400, 400), (64, 9000), (1000, 1000)])
def test_lap_random_vs_scipy(handle, shape):
    from scipy.optimize import linear_sum_assignment

    n, m = shape
    rng = np.random.default_rng(n * 31 + m)
    w = rng.standard_normal((n, m)) * 0.1 + 0.05
    col, obj = _lap_gpu(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert len(np.unique(col)) == n and col.min() >= 0 and col.max() < m
    ref = w[r, c].sum()
    assert abs(obj - ref) <= 1e-12 * max(1.0, abs(ref))
    assert abs(w[np.arange(n), col].sum() - obj) <= 1e-12 * max(1.0, abs(ref))
    assert (col == c).all()  # generic real costs: the optimum is unique


@pytest.mark.parametrize("kind", ["zeros", "small_ints", "dup_cols", "dup_rows", "planted_flat", "negative"])
def test_lap_ties_and_degenerate(handle, kind):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(11)
    n, m = 60, 90
    if kind == "zeros":
        w = np.zeros((n, m))
    elif kind == "small_ints":
        w = rng.integers(0, 4, size=(n, n)).astype(np.float64)
    elif kind == "dup_cols":
        base = rng.random((n, m // 2))
        w = np.concatenate([base, base], axis=1)
    elif kind == "dup_rows":
        base = rng.random((n // 2, m))
        w = np.concatenate([base, base], axis=0)
    elif kind == "planted_flat":
        w = 0.17 + 1e-9 * rng.standard_normal((n, n))
    else:
        w = -rng.random((n, n)) - 5.0
    col, obj = _lap_gpu(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert len(np.unique(col)) == w.shape[0] and col.min() >= 0
    ref = w[r, c].sum()
    assert abs(obj - ref) <= 1e-12 * max(1.0, abs(ref)), (obj, ref)


@pytest.mark.parametrize("mn", [(9, 4), (3, 7), (5, 5), (8, 4), (5, 1), (1, 1), (1, 3), (101, 10), (64, 64), (130, 64),
                                (700, 150)])
def test_step_loop_vs_oracle(handle, mn):
    torch = _torch()
    from oracle import restatement as R

    M, N = mn
    rng = np.random.default_rng(M * 3 + N)
    corrs = 0.2 * rng.random((M, N)) - 0.03
    a_ref, s_ref, o_ref = R.step_loop(corrs)
    d_c = _dev(corrs)
    d_ct = _dev(corrs.T)
    a, s, o, stats = handle.lap_steps(d_c.data_ptr(), N, d_ct.data_ptr(), M, M, N)
    assert (s == s_ref).all() and (a == a_ref).all()
    assert np.allclose(o, o_ref, rtol=1e-12, atol=1e-14)
    assert stats.n_steps == R.n_steps(M, N)


# ------------------------------------------------------------------------------------------------
# whole path on the seeded synthetic configs against the committed oracle outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_synthetic_config_vs_golden(handle, name):
    from macrodna_b200 import synth

    g = np.load(os.path.join(GOLDEN, "synth_%s.npz" % name))
    inst = synth.make_config_arrays(name)
    M, G = inst.rna.shape
    N = inst.dna.shape[0]
    corr = np.empty((M, N))
    assign, step, objs, stats = handle.cell2cell(inst.rna, inst.dna, M, N, G, corr_out=corr)
    # correlations: north-star gate 1e-6 abs; the FP64 path is held to 1e-12
    assert np.abs(corr[g["sample_i"], g["sample_j"]] - g["sample_corr"]).max() < 1e-12
    assert abs(corr.sum() - g["corr_sum"][0]) <= 1e-9 * abs(g["corr_sum"][0])
    assert np.allclose(objs, g["objs"], rtol=1e-9)
    assert (assign >= 0).all()
    ident, rep = tie_report(corr, assign, step, g["assign"], g["step"])
    assert rep["objective_ok"], rep
    assert ident, rep  # tie-free instance: assignments and step tags are bit-identical to the oracle's
    q, r = divmod(M, N)
    assert np.bincount(step)[1:].tolist() == [N] * q + ([r] if r else [])


def test_exact_tie_instance_is_stepwise_optimal(handle):
    """A zero-variance DNA cell (all correlations exactly 0.0) makes step 1 an exact tie: which RNA cell
    it takes is arbitrary (scipy, Gurobi and this solver each pick one) and changes the later steps'
    sub-problems.  Gate: step-1 objective equals the oracle's, and every step is the exact optimum of
    its own sub-problem (certified by scipy on the GPU's active set); the differing cells are reported."""
    from scipy.optimize import linear_sum_assignment

    from macrodna_b200 import synth
    from oracle import restatement as R

    inst = synth.make_config_arrays("C3", scale=0.3, ties=True)
    M, G = inst.rna.shape
    N = inst.dna.shape[0]
    corr = np.empty((M, N))
    assign, step, objs, stats = handle.cell2cell(inst.rna, inst.dna, M, N, G, corr_out=corr)
    c_ref, a_ref, s_ref, o_ref = R.cell2cell_arrays(inst.rna, inst.dna)
    assert np.abs(corr - c_ref).max() < 1e-12
    assert (corr[:, N // 2] == 0).all()
    assert abs(objs[0] - o_ref[0]) <= 1e-12 * abs(o_ref[0])
    active = np.arange(M)
    for s in range(1, step.max() + 1):
        rows = np.flatnonzero(step == s)
        assert np.isin(rows, active).all() and len(np.unique(assign[rows])) == len(rows)
        sub = c_ref[active]
        r, c = linear_sum_assignment(sub, maximize=True)
        opt = sub[r, c].sum()
        assert abs(c_ref[rows, assign[rows]].sum() - opt) <= 1e-12 * abs(opt)
        assert abs(objs[s - 1] - opt) <= 1e-12 * abs(opt)
        active = np.setdiff1d(active, rows)
    assert len(active) == 0
    print("TIE REPORT: cells differing from the scipy-tie-broken oracle:", int((assign != a_ref).sum()))


def test_determinism(handle):
    from macrodna_b200 import synth

    inst = synth.make_config_arrays("C3", scale=0.25)
    M, G = inst.rna.shape
    N = inst.dna.shape[0]
    a1, s1, o1, _ = handle.cell2cell(inst.rna, inst.dna, M, N, G)
    a2, s2, o2, _ = handle.cell2cell(inst.rna, inst.dna, M, N, G)
    assert (a1 == a2).all() and (s1 == s2).all() and (o1 == o2).all()


def test_nonfinite_cost_matrix_is_reported(handle):
    """mcd_lap_max / mcd_lap_steps on a user matrix with NaN must report MCD_ERR_NONFINITE, not crash."""
    torch = _torch()
    w = np.random.default_rng(1).random((40, 60))
    w[3, 7] = np.nan
    d_w = _dev(w)
    d_col = torch.zeros(40, dtype=torch.int32, device="cuda")
    st = handle.lib.mcd_lap_max(handle.h, d_w.data_ptr(), 40, 60, 60, d_col.data_ptr(), None)
    assert st == -4
    # and the handle is still usable afterwards
    w[3, 7] = 0.5
    col, obj = _lap_gpu(handle, w)
    assert len(np.unique(col)) == 40


def test_variant_return_signatures():
    """The reference repo's copies of the class drift in their return signature."""
    from macrodna_b200 import MaCroDNA, synth
    from oracle import restatement as R

    inst = synth.make_arrays(90, 25, 300, 3, seed=5)
    rna, dna, lab = synth.make_frames(inst)
    o = R.OracleMaCroDNA(rna.copy(), dna.copy(), lab)
    o.cell2cell_assignment()
    ref = o.last
    matched = ref["corrs"][np.arange(90), ref["assign"]]

    res, tagged, total = MaCroDNA(rna.copy(), dna.copy(), variant="objective").cell2cell_assignment()
    assert abs(total - ref["objs"].sum()) <= 1e-12 * abs(ref["objs"].sum())
    assert list(tagged.columns) == ["predict_cell", "step"]

    res, tagged, total, med = MaCroDNA(rna.copy(), dna.copy(), variant="median").cell2cell_assignment()
    assert abs(med - np.median(matched)) < 1e-13

    res, tagged, total, n_iters = MaCroDNA(rna.copy(), dna.copy(), variant="loo").cell2cell_assignment()
    assert n_iters == 4 and tagged.index.name == "rna_cell"
    assert list(tagged.columns) == ["predicted_dna_cell", "step", "corr_val"]
    assert np.abs(tagged["corr_val"].to_numpy() - matched).max() < 1e-13
    assert tagged["step"].tolist() == ref["step"].tolist()

    df = MaCroDNA(rna.copy(), dna.copy(), variant="resampling").cell2cell_assignment()
    assert list(df.columns) == ["predicted_dna_cell", "rna_cell", "step"]
    assert df["rna_cell"].tolist() == list(rna.columns)
    assert df["predicted_dna_cell"].tolist() == [dna.columns[j] for j in ref["assign"]]


def test_replicate_sweep_matches_oracle(handle):
    """Config-4 style sweep: DNA columns resampled with replacement per clone (exact duplicate columns)."""
    from macrodna_b200 import dist as mdist
    from macrodna_b200 import synth
    from oracle import restatement as R

    inst = synth.make_arrays(120, 30, 400, 3, seed=9)
    cols = [synth.resample_dna_columns(inst.dna_clone, seed=r) for r in range(4)]
    out = mdist.sweep_assignments(handle, inst.rna, inst.dna, cols, world=2, rank=1)
    assert sorted(out) == [1, 3]  # replicas only: rank 1 of 2 owns the odd replicates
    for r, (assign, step, objs) in out.items():
        sub = inst.dna[cols[r]]
        c_ref, a_ref, s_ref, o_ref = R.cell2cell_arrays(inst.rna, sub)
        assert np.allclose(objs, o_ref, rtol=1e-12)
        assert (step == s_ref).all()
        # duplicated DNA cells are interchangeable: compare the ORIGINAL cell each RNA cell was given
        assert (cols[r][assign] == cols[r][a_ref]).all()


def test_class_level_config3_with_device_gene_gather():
    """Whole drop-in path at config 3: frames with shuffled RNA gene order and 3 % extra RNA genes, so the gene
    intersection is applied by the device-side gather; result must equal the oracle's golden assignment."""
    from macrodna_b200 import MaCroDNA, synth

    g = np.load(os.path.join(GOLDEN, "synth_C3.npz"))
    inst = synth.make_config_arrays("C3")
    rna, dna, lab = synth.make_frames(inst)
    assert list(rna.index) != list(dna.index) and len(rna.index) > len(dna.index)
    m = MaCroDNA(rna, dna, lab)
    res, tagged = m.cell2cell_assignment()
    dna_cells = list(dna.columns)
    assert res["predict_cell"].tolist() == [dna_cells[j] for j in g["assign"]]
    assert tagged["step"].tolist() == g["step"].tolist()
    assert np.allclose(m.last_objective, g["objs"], rtol=1e-9)
    # the lazily materialised side effect of macrodna.py:90-91
    assert list(m.rna_df.index) == list(dna.index) and m.rna_df.shape == (len(dna.index), rna.shape[1])
    clone = MaCroDNA(rna, dna, lab).cell2clone_assignment()
    assert (clone["predict_clone"].to_numpy() == inst.dna_clone[g["assign"]]).all()


@pytest.mark.parametrize("mn", [(700, 5), (300, 299), (50, 700), (2500, 1200)])
def test_step_loop_shapes_many_steps_and_wide(handle, mn):
    """140-step schedule (more steps than the per-step stats slots), near-square, M << N, and a mid-size case that
    exercises the candidate-list kernels and the eps-scaled square-free last step."""
    from oracle import restatement as R

    M, N = mn
    rng = np.random.default_rng(M + N)
    base = rng.standard_normal((M, 8)) @ rng.standard_normal((8, N))  # low-rank structure: flat, contested costs
    corrs = 0.15 * np.tanh(base / 3.0) + 0.01 * rng.standard_normal((M, N))
    a_ref, s_ref, o_ref = R.step_loop(corrs)
    d_c, d_ct = _dev(corrs), _dev(corrs.T)  # keep the tensors alive across the call
    a, s, o, stats = handle.lap_steps(d_c.data_ptr(), N, d_ct.data_ptr(), M, M, N)
    assert np.allclose(o, o_ref, rtol=1e-12, atol=1e-13)
    assert (s == s_ref).all() and (a == a_ref).all()
    assert stats.n_steps == R.n_steps(M, N)


def test_standardize_split_long_rows(handle):
    """G beyond the register-resident capacity (declared 3-sweep variant), split-precision output."""
    torch = _torch()
    from oracle import restatement as R

    G, n = 30001, 9
    x = np.log1p(np.random.default_rng(4).poisson(5.0, size=(n, G)).astype(np.float64))
    lib, h = handle.lib, handle.h
    ldk = lib.mcd_padded_k_split(G)
    a2 = torch.empty((2, n, ldk), dtype=torch.int16, device="cuda")
    na = torch.empty(n, dtype=torch.float64, device="cuda")
    handle.check(lib.mcd_standardize_split(h, _dev(x).data_ptr(), n, G, G, a2.data_ptr(), na.data_ptr()))
    handle.synchronize()
    xc, nrm = R.standardise(x)
    a = a2.cpu().numpy()
    rec = (a[0].view(np.float16).astype(np.float64) + a[1].view(np.float16).astype(np.float64))[:, :G] / 256.0
    assert np.abs(rec - xc / nrm[:, None]).max() < 2.0 ** -21
    assert (a[:, :, G:] == 0).all()
