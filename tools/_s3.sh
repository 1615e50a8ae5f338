tools/gpu_session.sh s3 v2_s0h0 v2_s0h2000 v2_s64h0 v2 v2_s256
SELENITE_B200_LIB=build/ab/libv2_trace.so timeout 300 python tools/tc_trace.py > gpurun_out/s3_trace_v2.txt 2>&1; tail -5 gpurun_out/s3_trace_v2.txt
