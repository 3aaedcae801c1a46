python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/t7_pytest.log
LBZ_L64=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t7_bench_l64.json 2>> gpurun_out/t7_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t7_bench_pair.json 2>> gpurun_out/t7_bench.err
LBZ_ROUND_STATS=1 LBZ_LANES=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify 2>&1 | grep "sort depth" | tail -5 > gpurun_out/t7_rounds.log
cat gpurun_out/t7_pytest.log gpurun_out/t7_rounds.log; grep -h -o '"value": [0-9.]*' gpurun_out/t7_bench_*.json; tail -3 gpurun_out/t7_bench.err
