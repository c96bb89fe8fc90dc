"""Research: naive Jacobi auction until few bidders remain (or a round budget), then shortest augmenting paths (JV)
from the auction's duals.  Counts auction rounds and JV Dijkstra steps."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from make_inst import step_blocks
from sim_auction import top2


def auction_until(W, stop_nu, max_rounds):
    n, m = W.shape
    p = np.zeros(m)
    owner = -np.ones(m, int)
    col = -np.ones(n, int)
    profit = np.zeros(n)
    un = np.arange(n)
    hist = []
    while un.size > stop_nu and len(hist) < max_rounds:
        V = W[un] - p
        v1, j1, v2, j2 = top2(V)
        gam = v1 - v2
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
                profit[un[k]] = v1[k] - gam[k]
                won[k] = True
        nxt.extend(un[~won].tolist())
        hist.append(un.size)
        un = np.array(sorted(nxt), int)
    return p, owner, col, profit, un, np.array(hist)


def jv_finish(W, p, owner, col, profit, free):
    """Sequential shortest augmenting paths in max-form duals: reduced cost r_ij = profit_i + p_j - W_ij >= 0."""
    n, m = W.shape
    steps_per = []
    for f in free:
        dist = np.full(m, np.inf)
        pred = np.full(m, -1)
        done = np.zeros(m, bool)
        # free person's profit: max_j (W - p) makes its row feasible
        profit[f] = (W[f] - p).max()
        i = f
        di = 0.0
        steps = 0
        scanned_rows = [(f, 0.0)]
        while True:
            steps += 1
            r = di + (profit[i] + p - W[i])
            upd = (r < dist) & ~done
            dist[upd] = r[upd]
            pred[upd] = i
            cand = np.where(done, np.inf, dist)
            j = int(cand.argmin())
            dj = cand[j]
            done[j] = True
            if owner[j] < 0:
                sink = j
                D = dj
                break
            i = owner[j]
            di = dj
            scanned_rows.append((i, dj))
        # dual update
        for (i, di) in scanned_rows:
            profit[i] -= (D - di)
        sc = done.copy(); sc[sink] = False
        p[sc] += (D - dist[sc])
        # augment
        j = sink
        while True:
            i = pred[j]
            owner[j] = i
            jn = col[i]
            col[i] = j
            if i == f:
                break
            j = jn
        steps_per.append(steps)
    return np.array(steps_per)


if __name__ == "__main__":
    wl = sys.argv[1]
    only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] else []
    stops = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.02, 0.1]
    d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
    corr = d["corr"]
    for s, W in step_blocks(corr):
        if only and s not in only:
            continue
        n, m = W.shape
        r, c = linear_sum_assignment(W, maximize=True)
        ref = W[r, c].sum()
        for frac in stops:
            for maxr in (50, 200, 10**6):
                t0 = time.time()
                p, owner, col, profit, un, hist = auction_until(W, int(frac * n), maxr)
                sp = jv_finish(W, p, owner, col, profit, list(un))
                obj = W[np.arange(n), col].sum()
                print("step", s, W.shape, "stop_nu", int(frac * n), "maxr", maxr, "| auction rounds", len(hist), "bids", hist.sum(), "free", un.size,
                      "JV steps total", sp.sum(), "max", sp.max() if sp.size else 0, "mean %.1f" % (sp.mean() if sp.size else 0),
                      "gap %.2e" % (ref - obj), "t=%.1f" % (time.time() - t0), flush=True)
