"""Summarise ncu outputs into profiles/: launch-list shares and per-kernel raw metrics."""
import csv, collections, subprocess, sys, json, os
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out = {}
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9, 'second': 1e9}.get(u, 1)
    name = r[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out["launch_list"] = [{"kernel": k, "launches": v[0], "total_ms": round(v[1] / 1e6, 3), "share": round(v[1] / tot, 4)}
                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]
want = {'gpu__time_duration.sum': 'ms', 'dram__bytes_read.sum': 'read', 'dram__bytes_write.sum': 'write',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct', 'launch__registers_per_thread': 'regs',
        'launch__grid_size': 'grid', 'launch__block_size': 'block', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active': 'tensor_inst_pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64_pct', 'lts__t_bytes.sum': 'l2_bytes',
        'sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active': 'dmma_pct'}
idx = {h.index(k): v for k, v in want.items() if k in h}
units = rr[1]
ks = []
for r in rr[2:]:
    d = {"kernel": r[h.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '')}
    for i, name in idx.items():
        d[name] = r[i] + (" " + units[i] if units[i] else "")
    ks.append(d)
out["full_capture"] = ks
tensor_cols = [c for c in h if 'tensor' in c.lower()][:40]
out["tensor_metric_names_present"] = tensor_cols
json.dump(out, open(os.path.join("profiles", tag + "_ncu_summary.json"), "w"), indent=1)
print(json.dumps(out["launch_list"][:8], indent=0))
for k in ks: print(k)
