import sys, time
import numpy as np
from make_inst import step_blocks
from sim_auction import jacobi_auction
wl = sys.argv[1]
d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
corr = d["corr"]
only = [int(x) for x in sys.argv[2:]]
for s, W in step_blocks(corr):
    if only and s not in only: continue
    t0 = time.time()
    col, p, rounds, hist = jacobi_auction(W)
    print("step", s, W.shape, "rounds", rounds, "bids", hist.sum(), "narrow", (hist <= 32).sum(), "nu==1", (hist == 1).sum(), "t=%.1f" % (time.time() - t0), flush=True)
