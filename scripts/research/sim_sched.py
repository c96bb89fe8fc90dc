"""Research: eps schedules of the square step (start factor, theta)."""
import sys, numpy as np
from make_inst import step_blocks
from sim_scaling import phase
wl, step = sys.argv[1], int(sys.argv[2])
d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
for s, W in step_blocks(d["corr"]):
    if s != step: continue
    n, m = W.shape; rng = W.max() - W.min()
    for f0 in (1/3, 1/9, 1/27, 1/81):
        for theta in (3, 4, 6):
            p = np.zeros(m); f = f0; tot = 0; per = []
            while True:
                eps = f * rng if f >= 1e-7 else 0.0
                col, owner, hist = phase(W, p, eps)
                tot += len(hist); per.append(len(hist))
                if eps == 0: break
                f /= theta
            print("f0 1/%d theta %d rounds %d %s" % (round(1/f0), theta, tot, per), flush=True)
