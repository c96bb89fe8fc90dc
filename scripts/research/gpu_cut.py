"""Research: early end of the eps-scaling phases (lap.scale_cut) on square last steps of several shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from macrodna_b200 import get_handle, _lib
h = get_handle(0)
dev = torch.device("cuda", 0)
shapes = [(400, 200, 4000, 4, 1), (1000, 500, 4000, 4, 2), (2000, 1000, 4000, 8, 3), (2000, 1000, 4000, 2, 4), (4000, 2000, 4000, 8, 5),
          (4000, 2000, 4000, 16, 6), (8000, 4000, 4000, 8, 7), (8000, 4000, 4000, 32, 8), (12000, 6000, 4000, 16, 9), (5000, 1000, 15000, 8, 21), (5000, 1000, 15000, 8, 22)]
sets = [dict(x.split("=") for x in s.split(",") if x) for s in sys.argv[1:]] or [{}]
for M, N, G, k, seed in shapes:
    rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, k, seed, dev)
    out = []
    ref = None
    for st in sets:
        defaults = {kk: h.get_option(kk) for kk in st}
        for kk, v in st.items():
            h.set_option(kk, float(v))
        for _ in range(2):
            a, s, o, stt = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE)
        d = stt.as_dict()
        if ref is None:
            ref = a.copy()
        out.append((round(d["step_ms"][-1], 2), d["step_rounds"][-1], round(d["ms_lap"], 1), int((a != ref).sum())))
        for kk, v in defaults.items():
            h.set_option(kk, v)
    print((M, N, k), out, flush=True)
    del rna, dna
