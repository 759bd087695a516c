#!/bin/bash
# Round-2 profiling pass (one gpurun call, one GPU): launch lists and ncu --set full captures -> gpurun_out/r02_*
# Every command runs plain first (exit code checked) and only then under ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --profile-mode --no-extras"
$B > gpurun_out/r02_plain_dmma.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_cfg2.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trmm_sumsq -s 20 -c 2 -o gpurun_out/r02_trmm_sumsq $B > gpurun_out/r02_ncu_trmm.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kstar_pack -s 20 -c 1 -o gpurun_out/r02_kstar_pack $B > gpurun_out/r02_ncu_kstar.log 2>&1
B8="python bench.py --steps 1 --warmup 3 --profile-mode --no-extras --engine int8"
$B8 > gpurun_out/r02_plain_int8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:oz_sumsq -s 20 -c 1 -o gpurun_out/r02_oz_sumsq $B8 > gpurun_out/r02_ncu_oz.log 2>&1
H="python tools/hvi_pass.py 16000000 300 2"
$H > gpurun_out/r02_hvi_pass.json 2> gpurun_out/r02_hvi_pass.err && \
ncu --set full --clock-control none --import-source on -k regex:acquisition_hvi -c 1 -o gpurun_out/r02_acq_hvi $H > gpurun_out/r02_ncu_hvi.log 2>&1
python tools/hvi_pass.py 4000000 150 3 >> gpurun_out/r02_hvi_pass.json 2>> gpurun_out/r02_hvi_pass.err
F="python tools/fit_time.py 64"
$F > gpurun_out/r02_fit_time.jsonl 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_cfg5_64settings.csv $F > gpurun_out/r02_ncu_cfg5.log 2>&1
ls -la gpurun_out | grep r02_
