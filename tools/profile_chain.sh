#!/bin/bash
# Run ON THE GPU BOX (via gpurun): one `ncu --set full` capture per chain kernel at a reduced length (profiling only).
#   usage: tools/profile_chain.sh <tag> <which: rx|tx|chan> [seconds]
set -u
TAG=${1:-x}; WHICH=${2:-rx}; SEC=${3:-2}
mkdir -p gpurun_out
case $WHICH in rx) K=rx_ssb_tc;; tx) K=tx_ssb_tc;; chan) K=chan64;; q15) K=q15_tc;; esac
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_${WHICH}_full \
    python tools/bench_chains.py --which $WHICH --steps 1 --seconds $SEC > gpurun_out/${TAG}_${WHICH}_full.log 2>&1
tail -3 gpurun_out/${TAG}_${WHICH}_full.log
