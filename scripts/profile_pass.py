"""One warm pass + one measured pass of the hot path in the given precisions (command profiled under ncu).
usage: profile_pass.py [C5] [fp64,ozaki,split]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from macrodna_b200 import get_handle, _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
M, N, G, clones = bench.SHAPES[wl]
dev = torch.device("cuda", 0)
rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, clones, 1234 + int(wl[1:]), dev)
h = get_handle(0)
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp64", "ozaki", "split"]
for rep in range(2):
    for prec in modes:
        a, s, o, st = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE, precision=prec)
        d = st.as_dict()
        print(wl, prec, "pass", rep, {k: round(d[k], 3) for k in ("ms_standardize", "ms_corr", "ms_lap", "ms_total")}, flush=True)
