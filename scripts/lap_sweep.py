"""Sweep solver knobs (handle options, e.g. lap.theta=4,lap.aug_nu=64) on a synthetic workload; prints per-step ms / rounds."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from macrodna_b200 import get_handle, _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
M, N, G, clones = bench.SHAPES[wl]
dev = torch.device("cuda", 0)
rna, dna, _, _ = bench.make_device_instance(torch, M, N, G, clones, 1234 + int(wl[1:]), dev)
h = get_handle(0)
settings = [dict(x.split("=") for x in s.split(",") if x) for s in sys.argv[2:]] or [{}]
defaults = {k: h.get_option(k) for st in settings for k in st}
ref = None
for st in settings:
    for k, v in st.items():
        h.set_option(k, float(v))
    for rep in range(2):
        a, s_, o, stats = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE,
                                       precision=os.environ.get("MCD_SWEEP_PRECISION", "ozaki"))
    d = stats.as_dict()
    if ref is None:
        ref = (a.copy(), o.copy())
    same = bool((a == ref[0]).all())
    print(json.dumps({"set": st, "ms_lap": round(d["ms_lap"], 2), "step_ms": [round(x, 2) for x in d["step_ms"]],
                      "rounds": d["step_rounds"], "bids": d["lap_bids"], "aug": [d["lap_aug_rows"], d["lap_aug_steps"]],
                      "same_as_first": same, "obj_rel_diff": [float(abs(x - y) / abs(y)) for x, y in zip(o, ref[1])], "n_diff": int((a != ref[0]).sum()), "cyc_per_round": [round(c / max(1, sum(d["step_rounds"]))) for c in d["lap_cycles"]], "ms_corr": round(d["ms_corr"], 2), "ms_standardize": round(d["ms_standardize"], 3)}), flush=True)
    for k in st:
        h.set_option(k, defaults[k])
