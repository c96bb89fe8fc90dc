import numpy as np, time
from sim_scaling import *
d = np.load("../../.scratch/corr_torch_C3.npz"); corr = d["corr"]
for s, W in step_blocks(corr):
    if s != 7: continue
    n, m = W.shape
    rng = W.max() - W.min()
    p = np.zeros(m)
    eps = .25 * rng
    for it in range(12):
        t0 = time.time()
        col, owner, hist = phase(W, p, eps, max_rounds=20000)
        print(it, eps / rng, len(hist), hist[-5:], time.time() - t0, flush=True)
        lam = p[col].min(); un_obj = owner < 0; p[un_obj] = np.minimum(p[un_obj], lam)
        eps /= 4
