#!/bin/bash
# Run ON THE GPU BOX (via gpurun): GPU tests, then A/B libraries built by tools/ab_build.sh on the RX-SSB-f32 workload.
#   usage: tools/gpu_session.sh <tag> <lib tags for 1024 ch...>   (results under gpurun_out/<tag>_*)
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_env.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; ) > gpurun_out/${TAG}_pytest.log 2>&1
for v in "$@"; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 300 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/${TAG}_rx_$v.json 2>&1
done
tail -n 3 gpurun_out/${TAG}_pytest.log
for v in "$@"; do echo $v; tail -n 1 gpurun_out/${TAG}_rx_$v.json | cut -c1-200; done
