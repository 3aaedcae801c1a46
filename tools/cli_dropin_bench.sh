#!/bin/sh
# Throughput of the UNMODIFIED reference CLI + scheduler driving the GPU engine through the
# reference-shaped per-block API (oracle/_ref/lbzip2_gpu), next to the all-CPU reference binary.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import synth
open("/dev/shm/lbz_cli_in.raw", "wb").write(synth.text(100_000_000) * 10)
PY
for n in 32 32 64; do
  s=$(date +%s.%N)
  LBZIP2_B200_CONTEXTS=64 oracle/_ref/lbzip2_gpu -9 -n$n -c /dev/shm/lbz_cli_in.raw > /dev/shm/lbz_cli_gpu.bz2
  e=$(date +%s.%N)
  echo "lbzip2_gpu -9 -n$n: $(python -c "print(round(1000/($e-$s),1))") MB/s"
done
s=$(date +%s.%N); oracle/_ref/lbzip2 -9 -c /dev/shm/lbz_cli_in.raw > /dev/shm/lbz_cli_cpu.bz2; e=$(date +%s.%N)
echo "lbzip2 (CPU, all cores) -9: $(python -c "print(round(1000/($e-$s),1))") MB/s"
cmp /dev/shm/lbz_cli_gpu.bz2 /dev/shm/lbz_cli_cpu.bz2 && echo "outputs identical"
rm -f /dev/shm/lbz_cli_in.raw /dev/shm/lbz_cli_gpu.bz2 /dev/shm/lbz_cli_cpu.bz2
