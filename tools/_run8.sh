python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/t8_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t8_bench.json 2>> gpurun_out/t8_bench.err
LBZ_ROUND_STATS=1 LBZ_LANES=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify 2>&1 | grep "sort depth" | tail -5 > gpurun_out/t8_rounds.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload runs_fib > gpurun_out/t8_bench_runsfib.json 2>> gpurun_out/t8_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload random > gpurun_out/t8_bench_random.json 2>> gpurun_out/t8_bench.err
cat gpurun_out/t8_pytest.log gpurun_out/t8_rounds.log; grep -h -o '"value": [0-9.]*' gpurun_out/t8_bench*.json; tail -3 gpurun_out/t8_bench.err
