mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"
python tools/fit_time.py 256 > gpurun_out/r2k_fit_time.jsonl 2>&1; cat gpurun_out/r2k_fit_time.jsonl
B="python bench.py --steps 2 --warmup 3 --profile-mode --no-extras"
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trmm_sumsq -s 20 -c 2 -o gpurun_out/r02k_trmm_sumsq_grouped $B > gpurun_out/r02k_ncu_trmm.log 2>&1
H="python tools/hvi_pass.py 16000000 300 2"
$H > gpurun_out/r2k_hvi_pass.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:acquisition_hvi -c 1 -o gpurun_out/r02k_acq_hvi $H > gpurun_out/r02k_ncu_hvi.log 2>&1
cat gpurun_out/r2k_hvi_pass.json
