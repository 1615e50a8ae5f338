mkdir -p gpurun_out
for v in h1c2 h2c2 h2w h2w3; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s15_rx_$v.json 2>&1; echo $v; tail -1 gpurun_out/s15_rx_$v.json | cut -c1-200
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which rx --steps 5 --rx-channels 8192 --seconds 4 > gpurun_out/s15_rx8192_$v.json 2>&1; tail -1 gpurun_out/s15_rx8192_$v.json | cut -c1-200
done
