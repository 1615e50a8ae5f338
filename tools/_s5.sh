mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s5_pytest.log 2>&1; tail -12 gpurun_out/s5_pytest.log
for v in v4_base v4_ldw8 v4_sets3 v4_sets3_ldw8 v4_sets3_split_ldw8; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s5_rx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s5_rx_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libv4_sets3_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/s5_trace_sets3.txt 2>&1; tail -5 gpurun_out/s5_trace_sets3.txt
