"""Research: per-step solver rounds on the NumPy-synth instances (the ones the CPU simulation uses)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from macrodna_b200 import get_handle, synth

h = get_handle(0)
for wl in sys.argv[1:] or ["C3", "C4"]:
    inst = synth.make_config_arrays(wl)
    M, G = inst.rna.shape
    N = inst.dna.shape[0]
    for rep in range(2):
        a, s, o, st = h.cell2cell(inst.rna, inst.dna, M, N, G)
    d = st.as_dict()
    print(json.dumps({"wl": wl, "ms_lap": d["ms_lap"], "step_ms": [round(x, 2) for x in d["step_ms"]],
                      "rounds": d["step_rounds"], "bids": d["step_bids"], "obj": [float(x) for x in o]}), flush=True)
