"""Research: warm start of the square step from the exact duals of a row/column subsample, extended to all objects by
dual feasibility (p_j = max_{i in S} (W_ij - pi_i)); then eps-scaling from a small eps.  Counts Jacobi rounds."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_scaling import phase

def scaled_solve(W, p0, eps0_rel, theta, eps_min_rel=1e-7):
    n, m = W.shape
    rng = W.max() - W.min()
    p = p0.copy()
    f = eps0_rel
    tot = 0; nar = 0; bids = 0
    per = []
    while True:
        eps = f * rng if f >= eps_min_rel else 0.0
        col, owner, hist = phase(W, p, eps)
        tot += len(hist); nar += int((hist <= 32).sum()); bids += int(hist.sum())
        per.append(len(hist))
        if eps == 0.0: break
        f /= theta
    return col, p, tot, nar, bids, per

def exact_duals(W):
    """prices / profits of an exact solve (eps-scaling then eps = 0) of a square block"""
    col, p, tot, nar, bids, per = scaled_solve(W, np.zeros(W.shape[1]), 1/3, 3)
    n = W.shape[0]
    profit = W[np.arange(n), col] - p[col]
    return col, p, profit, tot

if __name__ == "__main__":
    wl, step = sys.argv[1], int(sys.argv[2])
    d = np.load("../../.scratch/corr_torch_%s.npz" % wl); corr = d["corr"]
    for s, W in step_blocks(corr):
        if s != step: continue
        n, m = W.shape
        r, c = linear_sum_assignment(W, maximize=True); ref = W[r, c].sum()
        col, p, tot, nar, bids, per = scaled_solve(W, np.zeros(m), 1/3, 3)
        print("cold: rounds", tot, "narrow", nar, "bids", bids, "gap", ref - W[np.arange(n), col].sum())
        rng_ = np.random.default_rng(0)
        for frac in (4, 8):
            S = np.sort(rng_.choice(n, n // frac, replace=False))
            T = np.sort(rng_.choice(m, m // frac, replace=False))   # square subsample
            Ws = W[np.ix_(S, T)]
            cs, ps, prof_s, tot_s = exact_duals(Ws)
            # extension by dual feasibility
            p0 = (W[S] - prof_s[:, None]).max(axis=0)
            p0 -= p0.min()
            for e0, th in ((1e-2, 3), (3e-3, 3), (1e-3, 3), (1e-2, 10), (1e-3, 10)):
                col, p, tot, nar, bids, per = scaled_solve(W, p0, e0, th)
                print("sub 1/%d (its own cold rounds %d): eps0 %.0e theta %d -> rounds %d narrow %d bids %d gap %.1e %s" % (
                    frac, tot_s, e0, th, tot, nar, bids, ref - W[np.arange(n), col].sum(), per), flush=True)
