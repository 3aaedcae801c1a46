python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/t10_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t10_bench.json 2>> gpurun_out/t10_bench.err
LBZ_LANES=1 ncu --set full --clock-control none --import-source on -k regex:k_text_pass2 -s 9 -c 2 -o gpurun_out/t10_text_pass2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > gpurun_out/t10_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/t10_ncu.log
cat gpurun_out/t10_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t10_bench.json; tail -3 gpurun_out/t10_ncu.log
