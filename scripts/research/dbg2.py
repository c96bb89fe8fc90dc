import numpy as np, time, sys
from sim_scaling import *
wl, step, theta = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
d = np.load("../../.scratch/corr_torch_%s.npz" % wl); corr = d["corr"]
for s, W in step_blocks(corr):
    if s != step: continue
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m)
    f = 1.0 / theta
    tot = 0
    while True:
        eps = f * rng if f >= 1e-7 else 0.0
        t0 = time.time()
        col, owner, hist = phase(W, p, eps, max_rounds=200000)
        tot += len(hist)
        print("eps %.2e rounds %d narrow %d nu1 %d bids %d  t=%.1f" % (f if eps else 0, len(hist), (hist <= 32).sum(), (hist == 1).sum(), hist.sum(), time.time() - t0), flush=True)
        if eps == 0: break
        f /= theta
    print("total rounds", tot)
