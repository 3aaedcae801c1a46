python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/t13_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t13_bench.json 2>> gpurun_out/t13_bench.err
LBZ_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rt_|k_rle1" -c 20 --csv --log-file gpurun_out/t13_rle_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/t13_ncu.log 2>&1
cat gpurun_out/t13_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t13_bench.json; tail -2 gpurun_out/t13_bench.err
