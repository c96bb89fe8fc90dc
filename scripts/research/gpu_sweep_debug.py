"""Research: find the replicate whose certificate fails (torch C4 instance of bench.py)."""
import os, sys, json
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
bad = []
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    cols = synth.resample_dna_columns(dch, seed=r).astype(np.int32)
    try:
        a, s, o, st = h.subinstance(None, cols, M=M, N=N)
        d = st.as_dict()
        print(r, "ok ms=%.1f rounds=%s aug=%d/%d gap=%.2e" % (d["ms_total"], d["step_rounds"], d["lap_aug_rows"], d["lap_aug_steps"], d["cert_rel_gap"]), flush=True)
    except Exception as e:
        print(r, "FAILED", e, flush=True)
        bad.append(r)
        h.set_option("debug", 1)
        try:
            h.subinstance(None, cols, M=M, N=N)
        except Exception:
            pass
        h.set_option("debug", 0)
print("bad", bad)
