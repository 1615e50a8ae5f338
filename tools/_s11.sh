mkdir -p gpurun_out
SELENITE_B200_LIB=build/ab/libtrace.so timeout 300 python tools/tc_trace.py > gpurun_out/s11_trace0.txt 2>&1; tail -4 gpurun_out/s11_trace0.txt; cp gpurun_out/tc_trace.txt gpurun_out/s11_raw0.txt
SELENITE_B200_LIB=build/ab/libtrace1.so timeout 300 python tools/tc_trace.py > gpurun_out/s11_trace1.txt 2>&1; tail -4 gpurun_out/s11_trace1.txt; cp gpurun_out/tc_trace.txt gpurun_out/s11_raw1.txt
