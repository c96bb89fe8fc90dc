import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.optimize import linear_sum_assignment
from macrodna_b200 import get_handle, synth
from oracle import restatement as R
h = get_handle(0)
inst = synth.make_arrays(300, 70, 500, 3, seed=42)
M, N, G = 300, 70, 500
corr = np.empty((M, N))
assign, step, objs, stats = h.cell2cell(inst.rna, inst.dna, M, N, G, corr_out=corr)
print("assign<0:", (assign < 0).sum(), "stats", stats.as_dict())
c_ref, a_ref, s_ref, o_ref = R.cell2cell_arrays(inst.rna, inst.dna)
print("objs gpu", objs, "\nobjs ref", o_ref)
active = np.arange(M)
for s in range(1, step.max() + 1):
    rows = np.flatnonzero(step == s)
    sub = corr[active]
    r, c = linear_sum_assignment(sub, maximize=True)
    print("step", s, "gpu matched", len(rows), "uniq dna", len(np.unique(assign[rows])), "gpu obj", corr[rows, assign[rows]].sum(),
          "scipy on gpu-active-set", sub[r, c].sum(), "rows subset of active", np.isin(rows, active).all())
    active = np.setdiff1d(active, rows)
# single LAP on step-2 problem
rows1 = np.flatnonzero(s_ref == 1)
act = np.setdiff1d(np.arange(M), rows1)
W = np.ascontiguousarray(c_ref[act].T)  # N x R
def lap(Wm):
    n, m = Wm.shape
    d_w = torch.from_numpy(Wm).cuda(); d_col = torch.full((n,), -7, dtype=torch.int32, device="cuda"); d_obj = torch.zeros(1, dtype=torch.float64, device="cuda")
    h.check(h.lib.mcd_lap_max(h.h, d_w.data_ptr(), n, m, m, d_col.data_ptr(), d_obj.data_ptr())); h.synchronize()
    cnt = torch.empty(0)
    return d_col.cpu().numpy(), float(d_obj.cpu()[0])
col, obj = lap(W)
r, c = linear_sum_assignment(W, maximize=True)
print("single LAP step2:", W.shape, "gpu", obj, "scipy", W[r, c].sum(), "uniq", len(np.unique(col)), "same", (col == c).all())
for env in [("MCD_LAP_MAX_ROUNDS", "0")]:
    os.environ[env[0]] = env[1]
    col, obj = lap(W)
    print("pure JV (max_rounds=0): gpu", obj, "uniq", len(np.unique(col)), "same", (col == c).all())
    del os.environ[env[0]]
# which persons differ
col, obj = lap(W)
diff = np.flatnonzero(col != c)
print("differing persons", diff, "zero-variance dna idx", N // 2)
