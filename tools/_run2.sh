python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/t2_pytest.log
LBZ_BWT_K=6 python -m pytest tests/test_gpu_parity.py -x -q -k "bwt or stream" 2>&1 | tail -3 >> gpurun_out/t2_pytest.log
LBZ_BWT_K=5 python -m pytest tests/test_gpu_parity.py -x -q -k "bwt or stream" 2>&1 | tail -3 >> gpurun_out/t2_pytest.log
for k in 8 7 6 5; do
  LBZ_BWT_K=$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t2_bench_k$k.json 2>> gpurun_out/t2_bench.err
done
LBZ_ROUND_STATS=1 LBZ_LANES=1 LBZ_BWT_K=8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify 2>&1 | grep "sort depth" | tail -6 > gpurun_out/t2_rounds_k8.log
LBZ_ROUND_STATS=1 LBZ_LANES=1 LBZ_BWT_K=6 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify 2>&1 | grep "sort depth" | tail -6 > gpurun_out/t2_rounds_k6.log
tools/cli_dropin_bench.sh 2000 > gpurun_out/t2_cli.log 2>&1
cat gpurun_out/t2_pytest.log gpurun_out/t2_rounds_k8.log gpurun_out/t2_rounds_k6.log; grep -h -o '"value": [0-9.]*' gpurun_out/t2_bench_k*.json; cat gpurun_out/t2_cli.log
