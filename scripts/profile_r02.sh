#!/bin/bash
# Round-2 profiling pass (run under gpurun on one B200): plain run first, then the ncu launch list of the same
# command and `--set full` captures of the top kernels (second pass of scripts/profile_pass.py C5 ozaki).
set -x
mkdir -p gpurun_out
timeout 300 python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_C5.csv \
    python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_ncu_launches.log 2>&1
# second pass: K1 (DNA, RNA), K2c, asynchronous wide kernel + master/helper tail of step 1, asynchronous kernel of step 2
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k 'regex:corr_ozaki_kernel|standardize_digits_stream|lap_async_kernel|lap_tail_mh_kernel|lap_tail_sym_kernel' \
    -s ${MCD_PROF_SKIP:-20} -c 6 -o gpurun_out/r02_prof_C5 -f python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_ncu_full.log 2>&1
# three mid-phase launches of the symmetric cluster tail (square step)
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:lap_tail_sym_kernel' -s 12 -c 3 \
    -o gpurun_out/r02_prof_C5_sym -f python scripts/profile_pass.py C5 ozaki > gpurun_out/r02_ncu_full_sym.log 2>&1
tail -5 gpurun_out/r02_plain.log gpurun_out/r02_ncu_full.log gpurun_out/r02_ncu_full_sym.log
