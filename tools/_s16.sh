mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_rx_ssb_f32.py tests/test_gpu_rx_ssb_tc.py tests/test_gpu_rx_fm_f32.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/s16_pytest.log 2>&1; tail -6 gpurun_out/s16_pytest.log
echo "---- block 48"
( SLB_TOL_BLOCK=48 timeout 600 python -m pytest tests/test_gpu_rx_ssb_f32.py tests/test_gpu_rx_ssb_tc.py tests/test_gpu_tx_ssb_f32.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s16_pytest48.log 2>&1; tail -14 gpurun_out/s16_pytest48.log
timeout 200 python tools/bench_chains.py --which rx --steps 10 > gpurun_out/s16_rx.json 2>&1; tail -1 gpurun_out/s16_rx.json | cut -c1-200
timeout 200 python tools/bench_chains.py --which rx --steps 5 --rx-channels 8192 --seconds 4 > gpurun_out/s16_rx8192.json 2>&1; tail -1 gpurun_out/s16_rx8192.json | cut -c1-200
