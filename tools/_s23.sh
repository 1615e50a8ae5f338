mkdir -p gpurun_out
for v in txr txu txs4; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which tx --steps 10 > gpurun_out/s23_tx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s23_tx_$v.json | cut -c1-200
done
SELENITE_B200_LIB=build/ab/libtxr.so timeout 300 python -m pytest tests/test_gpu_tx_ssb_f32.py -m gpu -q 2>&1 | tail -2
