"""GPU tests of the solver's long-row machinery: lists in every wide round with the cooperative chunked rebuild,
and the master/helper narrow-round kernel (8- and 16-CTA clusters).  Rows this long only occur at the large
config, so three cases are genuinely long (m > 16384 / m >= 32768); the rest run the same (default) path on small
problems, and once more with the optional single-CTA list tail (handle option "lap.list_max_m" = 16384)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lap(handle, w):
    import torch

    n, m = w.shape
    d_w = torch.from_numpy(np.ascontiguousarray(w)).cuda()
    d_col = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    d_obj = torch.zeros(1, dtype=torch.float64, device="cuda")
    handle.check(handle.lib.mcd_lap_max(handle.h, d_w.data_ptr(), n, m, m, d_col.data_ptr(), d_obj.data_ptr()))
    handle.synchronize()
    return d_col.cpu().numpy(), float(d_obj.cpu().numpy()[0])


@pytest.fixture(params=["0", "16384"], ids=["master_helper_tail", "single_cta_list_tail"])
def long_row_path(request, handle):
    handle.set_option("lap.list_max_m", int(request.param))
    yield
    handle.set_option("lap.list_max_m", 0)


def _clustered(rng, n, m, k):
    """Persons and objects in k groups with a common group effect: near-flat values inside a group, the regime
    where candidate lists run out and the narrow rounds are long."""
    gp, go = rng.integers(0, k, n), rng.integers(0, k, m)
    return 0.15 * (gp[:, None] == go[None, :]) + 0.02 * rng.standard_normal((n, m))


@pytest.mark.parametrize("shape", [(300, 20000), (500, 40000), (2000, 17000)])
def test_genuinely_long_rows_vs_scipy(handle, shape):
    from scipy.optimize import linear_sum_assignment

    n, m = shape
    rng = np.random.default_rng(n + m)
    w = _clustered(rng, n, m, 7)
    col, obj = _lap(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert len(np.unique(col)) == n and col.min() >= 0 and col.max() < m
    ref = w[r, c].sum()
    assert abs(obj - ref) <= 1e-12 * max(1.0, abs(ref))
    assert (col == c).all()


@pytest.mark.parametrize("shape", [(1, 2), (2, 3), (5, 9), (31, 130), (64, 127), (200, 1000), (700, 1500), (900, 901)])
def test_forced_long_row_path_vs_scipy(handle, long_row_path, shape):
    from scipy.optimize import linear_sum_assignment

    n, m = shape
    rng = np.random.default_rng(7 * n + m)
    w = _clustered(rng, n, m, 3)
    col, obj = _lap(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert len(np.unique(col)) == n and col.min() >= 0 and col.max() < m
    ref = w[r, c].sum()
    assert abs(obj - ref) <= 1e-12 * max(1.0, abs(ref))
    assert (col == c).all()


@pytest.mark.parametrize("kind", ["zeros", "dup_cols", "dup_rows", "planted_flat", "small_ints"])
def test_forced_long_row_path_ties(handle, long_row_path, kind):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(5)
    n, m = 80, 400
    if kind == "zeros":
        w = np.zeros((n, m))
    elif kind == "dup_cols":
        base = rng.random((n, m // 2))
        w = np.concatenate([base, base], axis=1)
    elif kind == "dup_rows":
        base = rng.random((n // 2, m))
        w = np.concatenate([base, base], axis=0)
    elif kind == "planted_flat":
        w = 0.17 + 1e-9 * rng.standard_normal((n, m))
    else:
        w = rng.integers(0, 4, size=(n, m)).astype(np.float64)
    col, obj = _lap(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert len(np.unique(col)) == n and col.min() >= 0
    ref = w[r, c].sum()
    assert abs(obj - ref) <= 1e-12 * max(1.0, abs(ref)), (obj, ref)


def test_forced_long_row_step_loop_vs_oracle(handle, long_row_path):
    import torch
    from oracle import restatement as R

    rng = np.random.default_rng(3)
    M, N = 2300, 400
    corrs = _clustered(rng, M, N, 5)
    c = torch.from_numpy(corrs).cuda()
    ct = torch.from_numpy(np.ascontiguousarray(corrs.T)).cuda()
    assign, step, objs, _ = handle.lap_steps(c.data_ptr(), N, ct.data_ptr(), M, M, N)
    a_ref, s_ref, o_ref = R.step_loop(corrs)
    assert np.allclose(objs, o_ref, rtol=1e-12, atol=1e-13)
    assert (assign == a_ref).all() and (step == s_ref).all()
