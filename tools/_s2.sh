tools/gpu_session.sh s2 v1 v1_walk v1_raw3 v1_walk_raw3
timeout 200 tools/microbench/pcie_rate > gpurun_out/s2_pcie_rate.txt 2>&1; cat gpurun_out/s2_pcie_rate.txt
SELENITE_B200_LIB=build/ab/libv1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rx_ssb_tc -s 3 -c 1 -o gpurun_out/s2_v1_rx_full python tools/bench_chains.py --which rx --steps 1 --seconds 2 > gpurun_out/s2_ncu.log 2>&1; tail -3 gpurun_out/s2_ncu.log
