LBZ_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rt_|k_rle1" -c 24 --csv --log-file gpurun_out/t12_rle_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/t12_ncu.log 2>&1
LBZ_RLE_V1=1 LBZ_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rt_|k_rle1" -c 2 --csv --log-file gpurun_out/t12_rle_launches_v1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/t12_ncu1.log 2>&1
tail -2 gpurun_out/t12_ncu.log
