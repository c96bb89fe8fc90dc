"""Research helper: synthetic correlation matrices + oracle step residuals, cached under .scratch/ (not shipped)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from macrodna_b200 import synth  # noqa: E402
from oracle import restatement as R  # noqa: E402


def corr_for(name, scale=1.0, gscale=None):
    tag = "%s_s%g_g%s" % (name, scale, gscale)
    p = os.path.join(ROOT, ".scratch", "corr_%s.npy" % tag)
    if os.path.exists(p):
        return np.load(p)
    m, n, g, k = synth.CONFIG_SHAPES[name]
    m, n = int(m * scale), int(n * scale)
    if gscale:
        g = int(g * gscale)
    inst = synth.make_arrays(m, n, g, k, seed=1234 + int(name[1:]))
    c = R.correlation_matrix(inst.rna, inst.dna)
    np.save(p, c)
    return c


def step_blocks(corr):
    """Yield (step, W[n, m]) persons x objects blocks exactly as the product's step loop builds them."""
    from scipy.optimize import linear_sum_assignment

    M, N = corr.shape
    act = np.arange(M)
    s = 0
    while act.size:
        sub = corr[act]
        if act.size > N:
            W = np.ascontiguousarray(sub.T)  # persons = DNA
        else:
            W = sub
        yield s, W
        r, c = linear_sum_assignment(sub, maximize=True)
        keep = np.ones(act.size, bool)
        keep[r] = False
        act = act[keep]
        s += 1


if __name__ == "__main__":
    name = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    c = corr_for(name, scale)
    print(c.shape, c.min(), c.max())
