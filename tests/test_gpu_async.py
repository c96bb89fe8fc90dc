"""GPU tests of the asynchronous wide kernel of the rectangular steps (lap_async_kernel): CTA workers that follow
eviction chains and apply bids with 128-bit compare-and-swaps.  A tie-free instance has one optimum, so the
asynchronous path, the round-synchronous path (handle option "deterministic" = 1) and SciPy must agree object for
object; instances with exact ties must come out the same run after run (the kernel notices the ties and the step is
redone round-synchronously)."""
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


def _clustered(rng, n, m, k):
    gp, go = rng.integers(0, k, n), rng.integers(0, k, m)
    return 0.15 * (gp[:, None] == go[None, :]) + 0.02 * rng.standard_normal((n, m))


@pytest.fixture
def synchronous(handle):
    def run(fn):
        handle.set_option("deterministic", 1)
        try:
            return fn()
        finally:
            handle.set_option("deterministic", 0)

    return run


@pytest.mark.parametrize("shape", [(64, 256), (300, 301), (257, 1000), (1000, 5000), (2000, 17000), (700, 40000)])
def test_async_equals_synchronous_and_scipy(handle, synchronous, shape):
    from scipy.optimize import linear_sum_assignment

    n, m = shape
    rng = np.random.default_rng(n * 7 + m)
    w = _clustered(rng, n, m, 5)
    col_a, obj_a = _lap(handle, w)
    col_s, obj_s = synchronous(lambda: _lap(handle, w))
    r, c = linear_sum_assignment(w, maximize=True)
    assert (col_a == c).all() and (col_s == c).all()
    assert obj_a == obj_s  # same assignment, same fixed-order sum
    for _ in range(2):  # run after run
        col_b, obj_b = _lap(handle, w)
        assert (col_b == col_a).all() and obj_b == obj_a


@pytest.mark.parametrize("shape", [(1, 256), (3, 1000), (8, 5000), (9, 5000), (255, 256), (2, 40000)])
def test_async_few_persons(handle, shape):
    """At most 8 persons: the asynchronous kernel stops before its first bid and the tail starts from scratch; 9: one
    bid is enough to hand over."""
    from scipy.optimize import linear_sum_assignment

    n, m = shape
    w = np.random.default_rng(n + m).standard_normal((n, m)) * 0.1
    col, _ = _lap(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert (col == c).all()


def test_async_starved_group(handle, synchronous):
    """More persons than objects in one group: the excess persons fight a price war over the group's objects before
    they leave it (long eviction chains, the regime the chain-following workers are for)."""
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(5)
    n, m = 600, 1500
    gp = np.repeat(np.arange(3), [400, 100, 100])
    go = np.repeat(np.arange(3), [300, 600, 600])
    w = 0.2 * (gp[:, None] == go[None, :]) + 0.01 * rng.standard_normal((n, m))
    col, obj = _lap(handle, w)
    r, c = linear_sum_assignment(w, maximize=True)
    assert (col == c).all()
    col_s, _ = synchronous(lambda: _lap(handle, w))
    assert (col_s == c).all()


def test_async_exact_ties_are_reproducible(handle):
    """Duplicated persons (identical rows) and duplicated objects (identical columns): many optima.  Every run must
    return the same one, and it must be optimal."""
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(11)
    base = _clustered(rng, 150, 700, 3)
    w = np.concatenate([base, base[:120]], axis=0)  # 120 persons twice
    w[:, 350:] = w[:, :350]  # every object twice
    r, c = linear_sum_assignment(w, maximize=True)
    best = w[r, c].sum()
    outs = [_lap(handle, w) for _ in range(4)]
    for col, obj in outs:
        assert len(set(col.tolist())) == w.shape[0]
        assert abs(obj - best) <= 1e-9 * abs(best)
        assert (col == outs[0][0]).all()


def test_async_whole_path_certified(handle, synchronous):
    """cell2cell through the C ABI: every step certified, identical to the synchronous path."""
    from macrodna_b200 import synth

    inst = synth.make_arrays(3000, 500, 1500, 4, seed=3)
    a, s, o, st = handle.cell2cell(inst.rna, inst.dna, 3000, 500, 1500)
    d = st.as_dict()
    assert d["cert_steps"] == 6 and d["cert_rel_gap"] <= 1e-12 and d["cert_bad"] == 0
    a2, s2, o2, _ = synchronous(lambda: handle.cell2cell(inst.rna, inst.dna, 3000, 500, 1500))
    assert (a == a2).all() and (s == s2).all() and (o == o2).all()
