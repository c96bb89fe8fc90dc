import numpy as np, sys
from sim_scaling import *
d = np.load("../../.scratch/corr_torch_C4.npz"); corr = d["corr"]; rc, dc = d["rna_clone"], d["dna_clone"]
for s, W in step_blocks(corr):
    if s != 4: continue
    # persons = remaining RNA rows (square): need their clones
    M, N = corr.shape
    from scipy.optimize import linear_sum_assignment
    act = np.arange(M)
    for t in range(4):
        r, c = linear_sum_assignment(corr[act], maximize=True)
        keep = np.ones(act.size, bool); keep[r] = False; act = act[keep]
    pclone = rc[act]; oclone = dc
    print("persons per clone", np.bincount(pclone), "objects per clone", np.bincount(oclone))
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m)
    col, owner, hist = phase(W, p, rng / 3)
    # phase 2 traced
    eps = rng / 9
    col = -np.ones(n, int); owner = -np.ones(m, int)
    un = np.arange(n); rounds = 0; log = []
    while un.size:
        V = W[un] - p
        v1, j1, v2, j2 = top2(V)
        gam = (v1 - v2) + eps
        order = np.lexsort((un, gam.astype(np.float32)))
        win = {}
        for k in order: win[j1[k]] = k
        won = np.zeros(un.size, bool); nxt = []
        for j, k in win.items():
            if owner[j] >= 0:
                col[owner[j]] = -1; nxt.append(owner[j])
            owner[j] = un[k]; col[un[k]] = j; p[j] += gam[k]; won[k] = True
        if un.size == 1:
            log.append((un[0], pclone[un[0]], j1[0], oclone[j1[0]], gam[0] / eps, v1[0]))
        nxt.extend(un[~won].tolist()); un = np.array(sorted(nxt), int); rounds += 1
    print("rounds", rounds, "single-bidder rounds", len(log))
    lg = np.array(log)
    print("person clones in chain", np.bincount(lg[:, 1].astype(int), minlength=8))
    print("object clones in chain", np.bincount(lg[:, 3].astype(int), minlength=8))
    print("gam/eps quantiles", np.quantile(lg[:, 4], [0, .25, .5, .75, 1]))
    print("distinct persons", len(set(lg[:, 0])), "distinct objects", len(set(lg[:, 2])))
    print("v1 start/end", lg[:5, 5], lg[-5:, 5])
    for q in range(0, len(lg), 100): print(q, lg[q])
