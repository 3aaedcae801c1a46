python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t14_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/t14_bench.json 2> gpurun_out/t14_bench.err
LBZ_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t14_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/t14_ncu_list.log 2>&1
LBZ_LANES=1 ncu --set full --clock-control none --import-source on -k regex:k_text_pass2 -s 1 -c 2 -o gpurun_out/t14_text_pass2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-verify > gpurun_out/t14_ncu_full.log 2>&1
cat gpurun_out/t14_pytest.log; grep -h -o '"value": [0-9.]*' gpurun_out/t14_bench.json; tail -2 gpurun_out/t14_bench.err
