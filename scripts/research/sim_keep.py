"""Research: square eps-scaling with partial restarts (assignments that still satisfy eps'-CS are kept)."""
import sys, numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_scaling import phase

def run(W, theta, keep, eps_min=1e-7):
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m); f = 1.0 / theta
    col = owner = None
    tot = 0; per = []; kept = []
    while True:
        eps = f * rng if f >= eps_min else 0.0
        if col is not None and keep:
            best = (W - p).max(axis=1)
            ok = (W[np.arange(n), col] - p[col]) >= best - eps
            if eps == 0.0:
                ok = (W[np.arange(n), col] - p[col]) >= best   # exact CS
            owner[col[~ok]] = -1; col[~ok] = -1
            kept.append(int(ok.sum()))
        else:
            col = owner = None
        col, owner, hist = phase(W, p, eps, col, owner)
        tot += len(hist); per.append(len(hist))
        if eps == 0.0: break
        f /= theta
    return tot, per, kept, W[np.arange(n), col].sum()

wl, step = sys.argv[1], int(sys.argv[2])
d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
for s, W in step_blocks(d["corr"]):
    if s != step: continue
    r, c = linear_sum_assignment(W, maximize=True); ref = W[r, c].sum()
    for theta in (3, 5, 10):
        for keep in (False, True):
            tot, per, kept, obj = run(W, theta, keep)
            print("theta", theta, "keep", keep, "rounds", tot, per, "kept", kept, "gap %.1e" % (ref - obj), flush=True)
