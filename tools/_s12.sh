mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/s12_pytest.log 2>&1; tail -5 gpurun_out/s12_pytest.log
for v in h2c3 h2c4 h2c2 h1c2 h1c3; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s12_rx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s12_rx_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libtrace.so timeout 300 python tools/tc_trace.py > gpurun_out/s12_trace.txt 2>&1; tail -4 gpurun_out/s12_trace.txt
