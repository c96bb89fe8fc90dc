"""Research: per-step solver counters of resampled (tie-heavy) C4 replicates."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from macrodna_b200 import get_handle, synth

h = get_handle(0)
inst = synth.make_config_arrays("C4")
M, G = inst.rna.shape
N = inst.dna.shape[0]
a, s, o, st = h.cell2cell(inst.rna, inst.dna, M, N, G)
print("base", json.dumps({k: st.as_dict()[k] for k in ("ms_lap", "step_ms", "step_rounds", "lap_aug_rows", "lap_aug_steps")}))
h.set_option("debug", 1)
for seed in range(2):
    cols = synth.resample_dna_columns(inst.dna_clone, seed=seed)
    print("replicate", seed, "distinct DNA cells", len(set(cols.tolist())), "of", len(cols), flush=True)
    a, s, o, st = h.subinstance(None, cols, M=M, N=N)
    d = st.as_dict()
    print(json.dumps({k: d[k] for k in ("ms_lap", "step_rounds", "lap_aug_rows", "lap_aug_steps", "lap_rounds")}), flush=True)
