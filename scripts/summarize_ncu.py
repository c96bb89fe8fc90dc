"""Summarise ncu outputs into profiles/<tag>_ncu_summary.json: launch-list shares and per-kernel metrics of the
`--set full` capture.   usage: summarize_ncu.py <launches.csv> <full.ncu-rep> <tag> [note]"""
import collections
import csv
import json
import os
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
out = {"note": note}


def short(name):
    return name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")


rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9}.get(r[ui], 1)
    agg[short(r[ki])][0] += 1
    agg[short(r[ki])][1] += v
tot = sum(v[1] for v in agg.values())
out["launch_list"] = [{"kernel": k, "launches": v[0], "total_ms": round(v[1] / 1e6, 3), "share": round(v[1] / tot, 4)}
                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__cluster_size": "cluster",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active": "imma_pct_active",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum": "tma_load_bytes",
        "sm__cycles_elapsed.avg.per_second": "sm_clock"}
idx = {h.index(k): v for k, v in want.items() if k in h}
ks = []
for r in rr[2:]:
    d = {"kernel": short(r[h.index("Kernel Name")])}
    for i, name in idx.items():
        d[name] = r[i] + (" " + units[i] if units[i] else "")
    ks.append(d)
out["full_capture"] = ks
json.dump(out, open(os.path.join("profiles", tag + "_ncu_summary.json"), "w"), indent=1)
print(json.dumps(out["launch_list"][:10], indent=0))
for k in ks:
    print(k)
