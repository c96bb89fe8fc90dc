"""Research: square-step eps schedules with the symmetric cluster tail (C5 shape, two seeds)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from macrodna_b200 import get_handle, _lib
h = get_handle(0)
dev = torch.device("cuda", 0)
scheds = [(0.0, 3.0), (0.037, 5.0), (0.037, 8.0), (0.037, 10.0), (0.02, 6.0), (0.06, 6.0), (0.02, 8.0), (0.01, 10.0)]
M, N, G, clones = bench.SHAPES["C5"]
for seed in (1239, 31):
    rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, clones, seed, dev)
    out = []
    for e0, th in scheds:
        h.set_option("lap.eps0", e0); h.set_option("lap.theta", th)
        for _ in range(2):
            a, s, o, st = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE)
        d = st.as_dict()
        out.append((e0, th, round(d["step_ms"][-1], 1), d["step_rounds"][-1]))
    print(seed, out, flush=True)
    del rna, dna
