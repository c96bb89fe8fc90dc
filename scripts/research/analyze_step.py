import sys
import numpy as np
wl, step = sys.argv[1], int(sys.argv[2])
d = np.load("../../.scratch/corr_torch_%s.npz" % wl)
corr, rc, dc = d["corr"], d["rna_clone"], d["dna_clone"]
M, N = corr.shape
from scipy.optimize import linear_sum_assignment
act = np.arange(M)
for s in range(step):
    r, c = linear_sum_assignment(corr[act], maximize=True)
    keep = np.ones(act.size, bool); keep[r] = False; act = act[keep]
print("remaining RNA", act.size)
k = max(rc.max(), dc.max()) + 1
print("clone: dna persons / rna objects remaining")
for q in range(k):
    print(q, (dc == q).sum(), (rc[act] == q).sum())
sub = corr[act]
r, c = linear_sum_assignment(sub, maximize=True)
cross = (rc[act][r] != dc[c]).sum()
print("optimal: matched pairs with clone mismatch:", cross, "of", len(r))
vals = sub[r, c]
print("matched value quantiles", np.quantile(vals, [0, .1, .5, .9, 1]))
