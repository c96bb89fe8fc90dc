"""Research: rect steps with heavy clone deficits (resampled replicate analog, no exact duplicates): naive auction vs
eps phases (cold eps_0 phase from zero prices, then restarts down to eps = 0) + count of lambda-violators."""
import sys, time
import numpy as np
from scipy.optimize import linear_sum_assignment
from sim_scaling import phase

d = np.load("../../.scratch/corr_torch_C4.npz")
corr, rc, dc = d["corr"], d["rna_clone"], d["dna_clone"]
M, N = corr.shape
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(seed)
clones = np.unique(dc)
props = rng.multinomial(N, rng.dirichlet(np.ones(len(clones))))
cols = []
for k, cnt in zip(clones, props):
    mem = np.flatnonzero(dc == k)
    cols.append(rng.choice(mem, size=min(cnt, len(mem)), replace=False))
cols = np.concatenate(cols)
print("persons per clone", [len(c) for c in np.split(cols, 1)], props, "n =", len(cols))
sub = corr[:, cols]
act = np.arange(M)
n = len(cols)
for s in range(4):
    W = np.ascontiguousarray(sub[act].T)   # persons = DNA
    r, c = linear_sum_assignment(W, maximize=True); ref = W[r, c].sum()
    m = W.shape[1]
    print("step", s, W.shape, "objects per clone", np.bincount(rc[act], minlength=8))
    rngW = W.max() - W.min()
    # naive
    p = np.zeros(m); t0 = time.time()
    col, owner, hist = phase(W, p, 0.0, max_rounds=400000)
    print("  naive: rounds %d narrow %d bids %d  t=%.1f" % (len(hist), (hist <= 32).sum(), hist.sum(), time.time() - t0), flush=True)
    for sched in ([1e-2, 0.0], [1e-3, 0.0], [1e-2, 1e-3, 1e-4, 0.0], [3e-2, 3e-3, 3e-4, 3e-5, 0.0]):
        p = np.zeros(m); tot = 0; per = []
        t0 = time.time()
        for f in sched:
            col, owner, hist = phase(W, p, f * rngW, max_rounds=400000)
            tot += len(hist); per.append(len(hist))
            if f > 0:   # prices of objects left unassigned go back to the floor for the next phase
                lam = p[col].min(); un = owner < 0; p[un] = np.minimum(p[un], 0.0)
        lam = p[col].min()
        viol = int(((owner < 0) & (p > lam)).sum())
        obj = W[np.arange(n), col].sum()
        print("  sched %s: rounds %d %s violators %d gap %.2e t=%.1f" % (sched, tot, per, viol, ref - obj, time.time() - t0), flush=True)
    keep = np.ones(act.size, bool); keep[c] = False; act = act[keep]
