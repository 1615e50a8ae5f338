mkdir -p gpurun_out
for v in h2c3; do
SELENITE_B200_LIB=build/ab/lib$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rx_ssb_tc -s 3 -c 1 -o gpurun_out/s13_${v}_full python tools/bench_chains.py --which rx --steps 1 --seconds 2 > gpurun_out/s13_ncu_$v.log 2>&1; tail -2 gpurun_out/s13_ncu_$v.log
done
