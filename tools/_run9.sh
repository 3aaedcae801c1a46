python -m pytest tests/test_gpu_parity.py -x -q -k "stream or api or cli or threads or empty" 2>&1 | tail -3 > gpurun_out/t9_pytest.log
LBZ_SPLIT4=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t9_bench_split2.json 2>> gpurun_out/t9_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t9_bench_split4.json 2>> gpurun_out/t9_bench.err
cat gpurun_out/t9_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t9_bench*.json; tail -3 gpurun_out/t9_bench.err
