"""Research: eps-scaling for n < m with ONE reset before the exact phase: scaling phases run forward from empty
assignments with rising prices; before the final eps = 0 phase every object that ended the last scaling phase unowned gets
price 0 again, so the only priced objects are the n that the last scaling phase assigned.  If the exact phase ends with
the same object set there is no stale price and the result is optimal (checked against SciPy)."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_auction import top2


def phase(W, p, eps, max_rounds=200000):
    n, m = W.shape
    col = -np.ones(n, int)
    owner = -np.ones(m, int)
    un = np.arange(n)
    hist = []
    stalled = 0
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
        prog = 0
        for j, k in win.items():
            if owner[j] < 0 or gam[k] > 1.4e-14:
                if owner[j] >= 0:
                    col[owner[j]] = -1
                    nxt.append(owner[j])
                owner[j] = un[k]
                col[un[k]] = j
                p[j] += gam[k]
                won[k] = True
                prog += 1
        nxt.extend(un[~won].tolist())
        hist.append(un.size)
        un = np.array(sorted(nxt), int)
        if prog == 0:
            stalled = un.size
            break
    return col, owner, np.array(hist), stalled


def cost(h):
    h = np.asarray(h)
    return 15.0 * (h > 32).sum() + 2.4 * (h <= 32).sum()


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
        rng = W.max() - W.min()
        col, owner, hist, st = phase(W, np.zeros(m), 0.0)
        print("step", s, W.shape, "naive: rounds", len(hist), "wide", int((hist > 32).sum()), "est us %.0f" % cost(hist), "stalled", st, flush=True)
        for theta, e0, emin in [(4, .25, 1e-9), (6, 1 / 27., 1e-9), (10, 0.1, 1e-9), (6, 1 / 27., 1e-7), (6, 1 / 27., 1e-11)]:
            p = np.zeros(m)
            eps = e0 * rng
            info = []
            tot_us = 0.0
            while eps >= emin * rng:
                col, owner, hist, st = phase(W, p, eps)
                info.append((len(hist), int((hist > 32).sum())))
                tot_us += cost(hist)
                eps /= theta
            objs_scaled = set(np.flatnonzero(owner >= 0).tolist())
            gap_scaled = ref - W[np.arange(n), col].sum()
            p[owner < 0] = 0.0
            col, owner, hist, st = phase(W, p, 0.0)
            info.append(("final", len(hist), int((hist > 32).sum()), "stalled", st))
            tot_us += cost(hist)
            stale = int(((owner < 0) & (p > 0)).sum())
            gap = ref - W[np.arange(n), col].sum() if st == 0 else float("nan")
            print("   theta %g e0 %.3g emin %g: est us %.0f  stale %d  gap_scaled %.3g gap %.3g" % (theta, e0, emin, tot_us, stale, gap_scaled, gap))
            print("      ", info, flush=True)
