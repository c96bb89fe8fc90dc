"""Research: sweep throughput vs concurrency."""
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
R = 240
cols = np.stack([synth.resample_dna_columns(dch, seed=r) for r in range(R)]).astype(np.int32)
for K in [int(x) for x in sys.argv[1:]]:
    h.subinstance_sweep(cols[:K], M=M, concurrency=K)
    t0 = time.perf_counter()
    a, s, o, g, st = h.subinstance_sweep(cols, M=M, concurrency=K)
    dt = time.perf_counter() - t0
    print("K", K, "rep/s %.1f" % (R / dt), "device ms", round(st.as_dict()["ms_total"]), "gapmax %.1e" % g.max(), flush=True)
