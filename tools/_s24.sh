mkdir -p gpurun_out
for v in q2 q3 q4 q4r; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which q15 --steps 10 > gpurun_out/s24_q15_$v.json 2>&1; echo $v; tail -1 gpurun_out/s24_q15_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libq4r.so timeout 300 python -m pytest tests/test_gpu_rx_ssb_q15.py -m gpu -q 2>&1 | tail -2
