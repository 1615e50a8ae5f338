mkdir -p gpurun_out
for v in cbase cnolb cnofft cnoboth; do
  SELENITE_B200_LIB=build/ab/lib$v.so timeout 200 python tools/bench_chains.py --which chan --steps 10 > gpurun_out/s20_chan_$v.json 2>&1; echo $v; tail -1 gpurun_out/s20_chan_$v.json | cut -c1-220
done
