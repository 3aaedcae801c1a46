python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t6_bench_n1.json 2> gpurun_out/t6_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t6_bench_n2.json 2> gpurun_out/t6_bench_n2.err
python -m pytest tests/test_host_taskgraph.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t6_pytest.log
CLI_CONFIGS='"32 2 64"' tools/cli_dropin_bench.sh 4000 2 > gpurun_out/t6_cli_2gpu.log 2>&1
cat gpurun_out/t6_pytest.log; grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/t6_bench_n*.json; grep "lbzip2" gpurun_out/t6_cli_2gpu.log; tail -3 gpurun_out/t6_bench_n2.err
