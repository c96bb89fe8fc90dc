"""Research: solver cycle counters of the same replicates at concurrency 1 and 16."""
import os, sys, json, time
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
R = 64
cols = np.stack([synth.resample_dna_columns(dch, seed=r) for r in range(R)]).astype(np.int32)
for K in (1, 4, 16):
    h.subinstance_sweep(cols[:max(2, K)], M=M, concurrency=K)
    t0 = time.perf_counter()
    a, s, o, g, st = h.subinstance_sweep(cols, M=M, concurrency=K)
    dt = time.perf_counter() - t0
    d = st.as_dict()
    cyc = d["lap_cycles"]
    print("K", K, "wall s %.2f" % dt, "rounds", d["lap_rounds"], "Gcycles wide[bid,bar1,res,bar2] %s tail[%s]" % (
        [round(c / 1e9, 2) for c in cyc[:4]], [round(c / 1e9, 2) for c in cyc[4:]]), flush=True)
