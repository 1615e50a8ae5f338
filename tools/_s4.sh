mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1; rc=$?; tail -3 gpurun_out/s4_smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; nvidia-smi | head -20; exit 1; fi
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s4_pytest.log 2>&1; tail -5 gpurun_out/s4_pytest.log
for v in v3_single v3_pair; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s4_rx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s4_rx_$v.json | cut -c1-200
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --rx-channels 8192 --seconds 4 --steps 5 > gpurun_out/s4_rx8192_$v.json 2>&1; tail -1 gpurun_out/s4_rx8192_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libv3_pair_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/s4_trace_pair.txt 2>&1; tail -5 gpurun_out/s4_trace_pair.txt
