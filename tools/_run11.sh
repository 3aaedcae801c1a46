python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/t11_pytest.log
LBZ_RLE_V1=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t11_bench_rlev1.json 2>> gpurun_out/t11_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t11_bench_rlev2.json 2>> gpurun_out/t11_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload runs_fib > gpurun_out/t11_bench_runsfib.json 2>> gpurun_out/t11_bench.err
cat gpurun_out/t11_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t11_bench*.json; tail -3 gpurun_out/t11_bench.err
