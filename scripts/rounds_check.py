import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macrodna_b200 import get_handle, synth
h = get_handle(0)
for name in ("C3", "C4"):
    inst = synth.make_config_arrays(name, ties=True)
    M, G = inst.rna.shape; N = inst.dna.shape[0]
    a, s, o, st = h.cell2cell(inst.rna, inst.dna, M, N, G)
    d = st.as_dict()
    print(name, "rounds", d["step_rounds"], "bids", d["step_bids"], "ms", [round(x, 2) for x in d["step_ms"]], "aug", d["lap_aug_rows"], d["lap_aug_steps"])
