"""Research: NumPy simulation of the product's Jacobi auction (eps = 0, naive increments) to study round counts."""
import sys
import time

import numpy as np

from make_inst import corr_for, step_blocks


def top2(V):
    j1 = V.argmax(axis=1)
    r = np.arange(V.shape[0])
    v1 = V[r, j1]
    V[r, j1] = -np.inf
    j2 = V.argmax(axis=1)
    v2 = V[r, j2]
    V[r, j1] = v1
    return v1, j1, v2, j2


def jacobi_auction(W, eps=0.0, price=None, max_rounds=10**6, verbose=False):
    n, m = W.shape
    p = np.zeros(m) if price is None else price.copy()
    owner = -np.ones(m, int)
    col = -np.ones(n, int)
    un = np.arange(n)
    rounds = 0
    hist = []
    while un.size and rounds < max_rounds:
        V = W[un] - p
        v1, j1, v2, j2 = top2(V)
        gam = (v1 - v2) + eps
        # winner per object: max gamma, then max person id
        order = np.lexsort((un, gam.astype(np.float32)))
        win = {}
        for k in order:  # later = larger wins
            win[j1[k]] = k
        nxt = []
        won = np.zeros(un.size, bool)
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
        rounds += 1
    return col, p, rounds, np.array(hist)


if __name__ == "__main__":
    name = sys.argv[1]
    corr = corr_for(name)
    for s, W in step_blocks(corr):
        t0 = time.time()
        col, p, rounds, hist = jacobi_auction(W)
        obj = W[np.arange(W.shape[0]), col].sum()
        print("step", s, W.shape, "rounds", rounds, "bids", hist.sum(), "narrow(<=32)", (hist <= 32).sum(),
              "nu==1", (hist == 1).sum(), "obj", obj, "t=%.1fs" % (time.time() - t0), flush=True)
