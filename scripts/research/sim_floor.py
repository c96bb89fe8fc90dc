"""Research: eps-scaling forward auction for n < m with a FLOOR-RAISE repair: after the exact (eps = 0) phase, unassigned
objects whose price exceeds the lowest assigned price are stale; raise the floor F = max stale price (every price below F
becomes F, owners of raised objects are unassigned) and continue the exact phase: afterwards every unassigned object sits
at F <= every assigned price, i.e. the state is optimal.  Counts Jacobi rounds."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_scaling import phase


def solve_floor(W, theta, eps0_rel, eps_min_rel, prefloor_q=None):
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m)
    eps = eps0_rel * rng
    info = []
    tot = 0; narrow = 0
    while eps >= eps_min_rel * rng:
        col, owner, hist = phase(W, p, eps)
        info.append((eps / rng, len(hist), int((hist <= 32).sum())))
        tot += len(hist); narrow += int((hist <= 32).sum())
        eps /= theta
    if prefloor_q is not None:
        # floor before the exact phase: q-quantile of the positive prices of objects unassigned at the end of scaling
        st = (owner < 0) & (p > 0)
        if st.any():
            F = np.quantile(p[st], prefloor_q)
            p = np.maximum(p, F)
    col, owner, hist = phase(W, p, 0.0)
    info.append((0.0, len(hist), int((hist <= 32).sum())))
    tot += len(hist); narrow += int((hist <= 32).sum())
    reps = 0
    while True:
        lam = p[col].min()
        stale = (owner < 0) & (p > lam)
        if not stale.any():
            break
        F = p[stale].max()
        low = p < F
        kicked = owner[low & (owner >= 0)]
        col[kicked] = -1
        owner[low] = -1
        p[low] = F
        col, owner, hist = phase(W, p, 0.0, col, owner)
        info.append(("repair", int(stale.sum()), kicked.size, len(hist), int((hist <= 32).sum())))
        tot += len(hist); narrow += int((hist <= 32).sum())
        reps += 1
    obj = W[np.arange(n), col].sum()
    return obj, tot, narrow, info


if __name__ == "__main__":
    wl = sys.argv[1]
    only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else []
    d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
    corr = d["corr"]
    for s, W in step_blocks(corr):
        if only and s not in only:
            continue
        n, m = W.shape
        if n == m:
            continue
        r, c = linear_sum_assignment(W, maximize=True)
        ref = W[r, c].sum()
        p0 = np.zeros(m)
        t0 = time.time()
        col, owner, hist = phase(W, p0, 0.0)
        print("step", s, W.shape, "naive rounds", len(hist), "narrow", int((hist <= 32).sum()), "gap", ref - W[np.arange(n), col].sum(), "%.1fs" % (time.time() - t0), flush=True)
        for theta, e0, emin, pq in [(4, .25, 1e-4, None), (4, .25, 1e-6, None), (6, 1 / 27., 1e-5, None), (4, .25, 1e-4, 0.5), (4, .25, 1e-4, 1.0)]:
            t0 = time.time()
            obj, tot, narrow, info = solve_floor(W, theta, e0, emin, pq)
            print("   theta %g e0 %g emin %g prefloor %s: rounds %d narrow %d gap %.3g  %.1fs" % (theta, e0, emin, pq, tot, narrow, ref - obj, time.time() - t0))
            print("      ", info, flush=True)
