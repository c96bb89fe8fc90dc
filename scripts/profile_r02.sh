#!/bin/bash
# Round-2 profiling pass (run under gpurun on one B200): plain run first, then the ncu launch list of the same
# command, one `--set full` capture of the top kernels, and a compute-sanitizer memcheck of the smoke test.
set -x
mkdir -p gpurun_out
timeout 300 python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_C5.csv \
    python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k 'regex:corr_ozaki_kernel|standardize_digits_stream|lap_cert_rows_kernel|lap_auction_kernel|lap_tail_mh_kernel|lap_tail_cluster_kernel' \
    -s 12 -c 14 -o gpurun_out/r02_prof_C5 -f python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -5 gpurun_out/r02_plain.log gpurun_out/r02_ncu_full.log gpurun_out/r02_sanitizer_memcheck.log
