LBZ_ROUND_STATS=1 LBZ_LANES=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify 2>&1 | grep "sort depth" | tail -5 > gpurun_out/t5_rounds.log
LBZ_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/t5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/t5_ncu_bench.log 2>&1
cat gpurun_out/t5_rounds.log
