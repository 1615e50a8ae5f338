mkdir -p gpurun_out
for v in h2c2 h2st h2sl h2hint h2u; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s14_rx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s14_rx_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libtrace.so timeout 300 python tools/tc_trace.py > gpurun_out/s14_trace.txt 2>&1; tail -4 gpurun_out/s14_trace.txt
