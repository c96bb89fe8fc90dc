import numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scipy.optimize import linear_sum_assignment
from macrodna_b200 import get_handle
h = get_handle(0)
def lap(w):
    n, m = w.shape
    d_w = torch.from_numpy(np.ascontiguousarray(w)).cuda()
    d_col = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    d_obj = torch.zeros(1, dtype=torch.float64, device="cuda")
    h.check(h.lib.mcd_lap_max(h.h, d_w.data_ptr(), n, m, m, d_col.data_ptr(), d_obj.data_ptr()))
    h.synchronize()
    return d_col.cpu().numpy()
rng = np.random.default_rng(0)
for n, m in [(1, 256), (1, 300), (3, 1000), (8, 5000), (9, 5000), (10, 257), (255, 256), (2, 40000), (33, 301)]:
    w = rng.standard_normal((n, m)) * 0.1
    c = lap(w)
    r, cc = linear_sum_assignment(w, maximize=True)
    print(n, m, bool((c == cc).all()), flush=True)
