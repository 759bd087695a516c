mkdir -p gpurun_out
python -m pytest tests/test_gpu_hvi.py -m gpu -q > gpurun_out/r2i_hvi_tests.log 2>&1; tail -3 gpurun_out/r2i_hvi_tests.log
python tools/hvi_pass.py 16000000 300 2 > gpurun_out/r2i_hvi_pass.json 2>&1; python tools/hvi_pass.py 16000000 3000 2 >> gpurun_out/r2i_hvi_pass.json 2>&1; cat gpurun_out/r2i_hvi_pass.json
for g in 1 0; do
  echo "BO_TRMM_GROUP=$g"
  BO_TRMM_GROUP=$g BO_I8_GUARD=0 python tools/oz_bench.py 1024 1000000 6 2 5 2>&1 | grep dmma | cut -c1-330
  BO_TRMM_GROUP=$g BO_I8_GUARD=0 python tools/oz_bench.py 2048 500000 8 3 3 2>&1 | grep dmma | cut -c1-330
done > gpurun_out/r2i_trmm_group_ab.txt 2>&1
cat gpurun_out/r2i_trmm_group_ab.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
