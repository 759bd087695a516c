mkdir -p gpurun_out
python -m pytest tests/test_gpu_hvi.py -m gpu -q > gpurun_out/r2j_hvi_tests.log 2>&1; tail -3 gpurun_out/r2j_hvi_tests.log
python tools/hvi_pass.py 16000000 300 2 > gpurun_out/r2j_hvi_pass.json 2>&1; python tools/hvi_pass.py 16000000 3000 2 >> gpurun_out/r2j_hvi_pass.json 2>&1; python tools/hvi_pass.py 4000000 150 3 >> gpurun_out/r2j_hvi_pass.json 2>&1; cat gpurun_out/r2j_hvi_pass.json
python tools/fit_time.py 256 > gpurun_out/r2j_fit_time.jsonl 2>&1; cat gpurun_out/r2j_fit_time.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_full_size.py tests/test_gpu_append.py -m gpu -q > gpurun_out/r2j_tests.log 2>&1; tail -3 gpurun_out/r2j_tests.log
