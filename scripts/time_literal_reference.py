"""R0 baseline: the reference file itself, byte-unmodified (oracle/run_reference.py: SciPy-backed
gurobipy stand-in), timed on the synthetic configs it can finish.  /root/reference only exists in the build
container, so this is run THERE and its record committed under profiles/; bench.py never calls it.
usage: python scripts/time_literal_reference.py C2 C3 > profiles/r02_R0_literal_reference.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macrodna_b200 import synth  # noqa: E402
from oracle import restatement as R  # noqa: E402
from oracle import run_reference  # noqa: E402

out = {"what": "literal reference src/MaCroDNA/macrodna.py cell2cell_assignment (Python pair loop macrodna.py:103-107 + "
               "per-step model build, scipy LSA standing in for Gurobi), single thread by construction",
       "host": {"cores": os.cpu_count()}, "runs": []}
for name in sys.argv[1:]:
    inst = synth.make_config_arrays(name)
    rna_df, dna_df, lab = synth.make_frames(inst, extra_rna_genes=0.0)
    t0 = time.perf_counter()
    (res, tagged), log = run_reference.run_reference(rna_df, dna_df, lab)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    corr = R.correlation_matrix(inst.rna, inst.dna)
    t_c = time.perf_counter() - t0
    a, s, o = R.step_loop(corr)
    t_l = time.perf_counter() - t0 - t_c
    dna_ids = list(dna_df.columns)
    same = [dna_ids[j] for j in a] == res["predict_cell"].tolist() and s.tolist() == tagged["step"].tolist()
    out["runs"].append({"config": name, "shape": list(inst.rna.shape[:1]) + list(inst.dna.shape), "R0_literal_s": t_ref,
                        "R1_port_s": t_c + t_l, "R1_corr_s": t_c, "R1_lsa_s": t_l, "R0_equals_R1_assignments": bool(same),
                        "pairs": int(inst.rna.shape[0] * inst.dna.shape[0]),
                        "R0_us_per_pair": 1e6 * t_ref / (inst.rna.shape[0] * inst.dna.shape[0])})
    print(json.dumps(out["runs"][-1]), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
