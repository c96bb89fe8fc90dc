"""Research: eps-scaling forward auction for n <= m, final eps = 0 phase, reverse fix-up of stale-priced unassigned
objects; counts Jacobi rounds per phase."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_auction import top2


def phase(W, p, eps, col=None, owner=None, max_rounds=10**6):
    """Jacobi forward auction phase from prices p.  col/owner: kept partial assignment (None = all unassigned)."""
    n, m = W.shape
    if col is None:
        col = -np.ones(n, int)
        owner = -np.ones(m, int)
    un = np.flatnonzero(col < 0)
    profit = np.zeros(n)
    hist = []
    while un.size and len(hist) < max_rounds:
        V = W[un] - p
        v1, j1, v2, j2 = top2(V)
        gam = (v1 - v2) + eps
        order = np.lexsort((un, gam.astype(np.float32)))
        win = {}
        for k in order:
            win[j1[k]] = k
        won = np.zeros(un.size, bool)
        nxt = []
        for j, k in win.items():
            if owner[j] < 0 or gam[k] > 0:
                if owner[j] >= 0:
                    col[owner[j]] = -1
                    nxt.append(owner[j])
                owner[j] = un[k]
                col[un[k]] = j
                p[j] += gam[k]
                won[k] = True
        nxt.extend(un[~won].tolist())
        hist.append(un.size)
        un = np.array(sorted(nxt), int)
    return col, owner, np.array(hist)


def reverse_fix(W, p, col, owner):
    """eps = 0 reverse iterations for unassigned objects priced above lambda = min assigned price."""
    n, m = W.shape
    profit = W[np.arange(n), col] - p[col]
    it = 0
    while True:
        lam = p[col].min()
        viol = np.flatnonzero((owner < 0) & (p > lam))
        if viol.size == 0:
            return it
        j = viol[0]
        b = W[:, j] - profit
        i1 = b.argmax(); beta = b[i1]
        b[i1] = -np.inf
        omega = b.max() if n > 1 else -np.inf
        it += 1
        if beta <= lam:
            p[j] = beta
            continue
        newp = max(omega, lam)
        jold = col[i1]
        owner[jold] = -1
        owner[j] = i1
        col[i1] = j
        p[j] = newp
        profit[i1] = W[i1, j] - newp


def solve(W, theta=4.0, eps0_rel=0.25, eps_min_rel=1e-6, keep=False, verbose=True):
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m)
    eps = eps0_rel * rng
    tot = 0; narrow = 0; totb = 0
    col = owner = None
    phases = []
    while True:
        if eps < eps_min_rel * rng:
            eps = 0.0
        if col is not None:
            lam = p[col].min()
            un_obj = owner < 0
            p[un_obj] = np.minimum(p[un_obj], lam)
            if keep and True:
                # keep assignments that still satisfy eps-CS at the new eps
                best = (W - p).max(axis=1)
                ok = (W[np.arange(n), col] - p[col]) >= best - eps
                if eps == 0.0:
                    ok &= False
                owner[col[~ok]] = -1
                col[~ok] = -1
            else:
                col = owner = None
        col, owner, hist = phase(W, p, eps, col, owner)
        phases.append((eps / rng, len(hist), int((hist <= 32).sum()), int(hist.sum())))
        tot += len(hist); narrow += int((hist <= 32).sum()); totb += int(hist.sum())
        if eps == 0.0:
            break
        eps /= theta
    nrev = reverse_fix(W, p, col, owner) if n < m else 0
    obj = W[np.arange(n), col].sum()
    return obj, tot, narrow, totb, nrev, phases


if __name__ == "__main__":
    wl = sys.argv[1]
    only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else []
    d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
    corr = d["corr"]
    for s, W in step_blocks(corr):
        if only and s not in only:
            continue
        r, c = linear_sum_assignment(W, maximize=True)
        ref = W[r, c].sum()
        for theta, e0, emin, keep in [(4, .25, 1e-6, False), (4, .25, 1e-6, True), (8, .125, 1e-6, False), (3, .33, 1e-7, False), (10, .1, 1e-5, False)]:
            t0 = time.time()
            obj, tot, narrow, totb, nrev, phases = solve(W, theta, e0, emin, keep)
            print("step", s, W.shape, "theta", theta, "e0", e0, "emin", emin, "keep", keep, "| rounds", tot, "narrow", narrow, "bids", totb,
                  "rev", nrev, "gap", ref - obj, "t=%.1f" % (time.time() - t0), flush=True)
            print("    ", [(("%.1e" % e), r_, nr) for e, r_, nr, b in phases])
