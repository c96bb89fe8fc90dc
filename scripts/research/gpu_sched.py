"""Research: square-step eps schedules on several instances (torch C3/C4/C5 seeds, NumPy-synth C3/C4)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from macrodna_b200 import get_handle, synth, _lib
h = get_handle(0)
dev = torch.device("cuda", 0)
scheds = [(0.0, 3.0), (0.037, 6.0), (0.037, 5.0), (0.05, 6.0), (0.025, 6.0)]
def run(tag, rna_p, dna_p, M, N, G, space):
    out = []
    for e0, th in scheds:
        h.set_option("lap.eps0", e0); h.set_option("lap.theta", th)
        for _ in range(2):
            a, s, o, st = h.cell2cell(rna_p, dna_p, M, N, G, in_space=space)
        d = st.as_dict()
        out.append((round(d["step_ms"][-1], 2), d["step_rounds"][-1], round(d["ms_lap"], 1)))
    print(tag, out, flush=True)
for wl, seeds in (("C3", (1237, 11, 12)), ("C4", (1238, 21, 22)), ("C5", (1239, 31))):
    M, N, G, clones = bench.SHAPES[wl]
    for seed in seeds:
        rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, clones, seed, dev)
        run("%s torch seed %d" % (wl, seed), rna.data_ptr(), dna.data_ptr(), M, N, G, _lib.MEM_DEVICE)
        del rna, dna
for wl in ("C3", "C4"):
    inst = synth.make_config_arrays(wl)
    run("%s numpy" % wl, inst.rna, inst.dna, inst.rna.shape[0], inst.dna.shape[0], inst.rna.shape[1], _lib.MEM_HOST)
