python -m pytest tests/test_gpu_parity.py -x -q -k "bwt or stream or rle1" 2>&1 | tail -3 > gpurun_out/t3_pytest.log
LBZ_TP_V1=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t3_bench_v1.json 2>> gpurun_out/t3_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t3_bench_v2.json 2>> gpurun_out/t3_bench.err
LBZ_TP_MINB=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t3_bench_v2_minb2.json 2>> gpurun_out/t3_bench.err
tools/_setup_probe 32 > gpurun_out/t3_probe.log 2>&1
tools/_setup_probe 64 >> gpurun_out/t3_probe.log 2>&1
cat gpurun_out/t3_pytest.log gpurun_out/t3_probe.log; grep -h -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/t3_bench_v*.json
