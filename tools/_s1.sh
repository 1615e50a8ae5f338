tools/gpu_session.sh s1 base v1 v1_noepi v1_nomma v1_nost
timeout 120 tools/microbench/umma_cta2 > gpurun_out/s1_umma_cta2.txt 2>&1; tail -30 gpurun_out/s1_umma_cta2.txt
SELENITE_B200_LIB=build/ab/libv1.so timeout 300 python tools/bench_chains.py --which rx --rx-channels 8192 --seconds 4 --steps 5 > gpurun_out/s1_rx8192_v1.json 2>&1; tail -1 gpurun_out/s1_rx8192_v1.json
SELENITE_B200_LIB=build/ab/libv1_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/s1_trace_v1.txt 2>&1; tail -12 gpurun_out/s1_trace_v1.txt
SELENITE_B200_LIB=build/ab/libbase_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/s1_trace_base.txt 2>&1; tail -5 gpurun_out/s1_trace_base.txt
