"""Research: rect steps solved by eps-scaled forward phases (restart, prices kept) + final eps = 0 forward phase +
naive reverse phase for unassigned objects priced above lambda.  Counts rounds, reverse iterations, verifies optimum."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from sim_scaling import phase
from make_inst import step_blocks


def reverse_phase(W, p, col, owner):
    n, m = W.shape
    profit = W[np.arange(n), col] - p[col]
    trivial = real = 0
    while True:
        lam = p[col].min()
        viol = np.flatnonzero((owner < 0) & (p > lam))
        if viol.size == 0:
            return trivial, real
        # Jacobi-style would do all at once; sequential here (counts only)
        j = viol[np.argmax(p[viol])]
        b = W[:, j] - profit
        i1 = int(b.argmax()); beta = b[i1]
        b[i1] = -np.inf
        omega = b.max() if n > 1 else -np.inf
        if beta <= lam:
            p[j] = lam if beta <= lam else beta
            p[j] = min(p[j], lam)
            trivial += 1
            continue
        real += 1
        newp = max(omega, lam)
        jold = col[i1]
        owner[jold] = -1
        owner[j] = i1
        col[i1] = j
        p[j] = newp
        profit[i1] = W[i1, j] - newp


def solve_fr(W, sched):
    n, m = W.shape
    rngW = W.max() - W.min()
    p = np.zeros(m)
    per = []
    for f in sched:
        col, owner, hist = phase(W, p, f * rngW, max_rounds=100000)
        per.append((len(hist), int((hist <= 32).sum())))
    nviol = int(((owner < 0) & (p > p[col].min())).sum())
    triv, real = reverse_phase(W, p, col, owner)
    obj = W[np.arange(n), col].sum()
    # certificate
    lam = p[col].min(); q = np.maximum(p - lam, 0)
    u = (W - q).max(axis=1); t = W[np.arange(n), col] - q[col]
    un = np.ones(m, bool); un[col] = False
    gap = (u - t).sum() + q[un].sum()
    return obj, per, nviol, triv, real, gap


if __name__ == "__main__":
    wl = sys.argv[1]
    only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else []
    d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
    for s, W in step_blocks(d["corr"]):
        if only and s not in only: continue
        if W.shape[0] == W.shape[1]: continue
        r, c = linear_sum_assignment(W, maximize=True); ref = W[r, c].sum()
        print("step", s, W.shape)
        for sched in ([0.0], [1e-2, 0.0], [1e-2, 1e-3, 0.0], [1e-2, 1e-3, 1e-4, 0.0], [3e-2, 3e-3, 3e-4, 3e-5, 0.0], [1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 0.0]):
            t0 = time.time()
            obj, per, nviol, triv, real, gap = solve_fr(W, sched)
            print("  sched %-40s rounds %6d %s | violators %d reverse trivial %d real %d | optgap %.1e cert %.1e  t=%.1f" % (
                sched, sum(x[0] for x in per), per, nviol, triv, real, ref - obj, gap, time.time() - t0), flush=True)
