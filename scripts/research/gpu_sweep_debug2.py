"""Research: concurrent sweep, per-replicate certificate gaps regardless of the call's status."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import bench
from macrodna_b200 import get_handle, synth, _lib

h = get_handle(0)
dev = torch.device("cuda", 0)
M, N, G, clones = bench.SHAPES["C4"]
rna, dna, rc, dc = bench.make_device_instance(torch, M, N, G, clones, 1238, dev)
dch = dc.cpu().numpy()
h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 40
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cols = np.stack([synth.resample_dna_columns(dch, seed=r) for r in range(R)]).astype(np.int32)
nsteps = 5
assign = np.full((R, M), -5, np.int32); step = np.zeros((R, M), np.int32); objs = np.zeros((R, nsteps)); gaps = np.zeros(R)
stats = _lib.McdStats()
st = h.lib.mcd_subinstance_sweep(h.h, R, None, M, cols.ctypes.data, N, assign.ctypes.data, step.ctypes.data, objs.ctypes.data,
                                 gaps.ctypes.data, K, C.byref(stats))
print("status", st, h.lib.mcd_last_error(h.h))
d = stats.as_dict()
print("ms", d["ms_total"], "cert_bad", d["cert_bad"], "gapmax", d["cert_rel_gap"])
for r in range(R):
    a1, s1, o1, st1 = (None,) * 4
    flag = "" if gaps[r] < 1e-9 else "  <<<<<<"
    print(r, "gap %.3e" % gaps[r], "unassigned", int((assign[r] < 0).sum()), "steps", np.bincount(step[r], minlength=6).tolist(), "obj", objs[r].round(6).tolist(), flag)
    if flag:
        a1, s1, o1, st1 = h.subinstance(None, cols[r], M=M, N=N)
        print("   one-at-a-time obj", o1.round(6).tolist(), "same assign", bool((a1 == assign[r]).all()))
