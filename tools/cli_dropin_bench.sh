#!/bin/sh
# CLI-level throughput on the GPU box, 1 GB of the benchmark text from /dev/shm to /dev/null:
#   oracle/_ref/lbzip2        the unmodified reference, all host cores (pthread CPU path)
#   oracle/_ref/lbzip2_gpu    unmodified reference host code + per-block API of libbz2b200.so
#   oracle/_ref/lbzip2_b200   reference CLI/scheduler + our batch-aware task graph
#                             (lbzip2_b200/host/compress_b200.c) + libbz2b200.so
# and a byte comparison of the three outputs; then `-d` of the reference's output through the reference
# CLI and through our expansion task graph (lbzip2_b200/host/expand_b200.c).  Usage: tools/cli_dropin_bench.sh [MB] [GPUS]
set -e
cd "$(dirname "$0")/.."
MB=${1:-1000}
GPUS=${2:-1}
echo "host: $(nproc) online CPUs, cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null || echo n/a), $(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2)"
python - "$MB" <<'PY'
import sys; sys.path.insert(0, "tests")
import synth
mb = int(sys.argv[1])
base = synth.text(100_000_000)
with open("/dev/shm/lbz_cli_in.raw", "wb") as f:
    left = mb * 1_000_000
    while left > 0:
        f.write(base[:left]); left -= len(base)
PY
IN=/dev/shm/lbz_cli_in.raw
t() { # label, command...
  label=$1; shift
  s=$(date +%s.%N); "$@"; e=$(date +%s.%N)
  echo "$label: $(python -c "print(round($MB/($e-$s),1))") MB/s"
}
head -c 20000000 $IN > /dev/shm/lbz_cli_20.raw
s=$(date +%s.%N); cat $IN > /dev/null; e=$(date +%s.%N)
echo "cat (page cache read): $(python -c "print(round($MB/($e-$s),1))") MB/s"
# first call pays the one-off driver/library page-in of a fresh box: warm up once, untimed
LBZIP2_B200_BATCH=8 LBZIP2_B200_ENGINES=1 oracle/_ref/lbzip2_b200 -9 -n8 -c /dev/shm/lbz_cli_20.raw > /dev/null
eval "set -- ${CLI_CONFIGS:-\"32 2 64\" \"16 4 64\" \"64 2 64\" \"32 3 64\"}"
for cfg in "$@"; do
  set -- $cfg
  for rep in 1 2; do
    t "lbzip2_b200 -9 -n$3 batch=$1 engines=$2 gpus=$GPUS" env LBZIP2_B200_BATCH=$1 LBZIP2_B200_ENGINES=$2 LBZIP2_B200_GPUS=$GPUS LBZIP2_B200_STATS=1 \
       sh -c "oracle/_ref/lbzip2_b200 -9 -n$3 -c $IN > /dev/shm/lbz_cli_b200.bz2"
  done
done
t "lbzip2_gpu  -9 -n64 (per-block API)" env LBZIP2_B200_CONTEXTS=64 sh -c "oracle/_ref/lbzip2_gpu -9 -n64 -c $IN > /dev/shm/lbz_cli_gpu.bz2"
t "lbzip2 (CPU reference, all cores) -9" sh -c "oracle/_ref/lbzip2 -9 -c $IN > /dev/shm/lbz_cli_cpu.bz2"
s=$(date +%s.%N); oracle/_ref/lbzip2 -9 -n1 -c /dev/shm/lbz_cli_20.raw > /dev/null; e=$(date +%s.%N)
echo "lbzip2 (CPU reference, -n1, 20 MB): $(python -c "print(round(20/($e-$s),1))") MB/s"
# decompression of the reference's output: reference CLI (all cores) vs our expansion task graph
for blocks in 320 1200; do
  t "lbzip2_b200 -d -n8 blocks/wave=$blocks" env LBZIP2_B200_DBLOCKS=$blocks LBZIP2_B200_DWAVE_MB=1024 LBZIP2_B200_STATS=1 \
     sh -c "oracle/_ref/lbzip2_b200 -d -n8 -c /dev/shm/lbz_cli_cpu.bz2 > /dev/shm/lbz_cli_b200.raw"
done
t "lbzip2 -d (CPU reference, all cores)" sh -c "oracle/_ref/lbzip2 -d -c /dev/shm/lbz_cli_cpu.bz2 > /dev/null"
cmp /dev/shm/lbz_cli_b200.raw $IN && echo "lbzip2_b200 -d output identical to the input"
rm -f /dev/shm/lbz_cli_b200.raw
cmp /dev/shm/lbz_cli_b200.bz2 /dev/shm/lbz_cli_cpu.bz2 && echo "lbzip2_b200 output identical to the reference's"
cmp /dev/shm/lbz_cli_gpu.bz2 /dev/shm/lbz_cli_cpu.bz2 && echo "lbzip2_gpu output identical to the reference's"
rm -f /dev/shm/lbz_cli_in.raw /dev/shm/lbz_cli_20.raw /dev/shm/lbz_cli_gpu.bz2 /dev/shm/lbz_cli_cpu.bz2 /dev/shm/lbz_cli_b200.bz2
