mkdir -p gpurun_out
for v in v4_base v4_sets3_ldw8; do
SELENITE_B200_LIB=build/ab/lib$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rx_ssb_tc -s 3 -c 1 -o gpurun_out/s6_${v}_full python tools/bench_chains.py --which rx --steps 1 --seconds 2 > gpurun_out/s6_ncu_$v.log 2>&1; tail -2 gpurun_out/s6_ncu_$v.log
done
