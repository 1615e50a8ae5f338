mkdir -p gpurun_out
timeout 300 python tools/fm_diag.py > gpurun_out/s8_fm_diag.txt 2>&1; cat gpurun_out/s8_fm_diag.txt | tail -30
