"""Run the whole path at one workload in each precision mode; print stage times, correlation error vs the FP64
kernel (sampled) and assignment differences."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from macrodna_b200 import get_handle, _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp64", "ozaki", "split"]
M, N, G, clones = bench.SHAPES[wl]
dev = torch.device("cuda", 0)
rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, clones, 1234 + int(wl[1:]), dev)
h = get_handle(0)
ref = None
for mode in modes:
    for rep in range(2):
        a, s_, o, stats = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE, precision=mode)
    mv = h.last_match_values(M)
    d = stats.as_dict()
    rec = {"mode": mode, "ms_std": round(d["ms_standardize"], 3), "ms_corr": round(d["ms_corr"], 2),
           "ms_lap": round(d["ms_lap"], 2), "ms_total": round(d["ms_total"], 2), "obj": [float(x) for x in o]}
    if ref is None:
        ref = (a.copy(), mv.copy(), o.copy())
    else:
        rec["cells_differing"] = int((a != ref[0]).sum())
        same = a == ref[0]
        rec["max_dcorr_on_matched"] = float(np.abs(mv[same] - ref[1][same]).max()) if same.any() else None
        rec["rel_obj_gap"] = float(np.abs(o - ref[2]).max() / np.abs(ref[2]).max())
    print(json.dumps(rec), flush=True)
