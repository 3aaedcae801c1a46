python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/t4_pytest.log
LBZ_ROUND_V1=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t4_bench_roundv1.json 2>> gpurun_out/t4_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t4_bench_roundv2.json 2>> gpurun_out/t4_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload random > gpurun_out/t4_bench_random.json 2>> gpurun_out/t4_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload runs_fib > gpurun_out/t4_bench_runsfib.json 2>> gpurun_out/t4_bench.err
df -h /dev/shm | tail -1 > gpurun_out/t4_cli.log; free -g | head -2 >> gpurun_out/t4_cli.log
CLI_CONFIGS='"32 2 64" "32 3 64" "16 4 64"' tools/cli_dropin_bench.sh 4000 >> gpurun_out/t4_cli.log 2>&1
cat gpurun_out/t4_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t4_bench_*.json; cat gpurun_out/t4_cli.log; tail -3 gpurun_out/t4_bench.err
