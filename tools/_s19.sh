mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tx_ssb_f32.py tests/test_gpu_ring.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/s19_pytest.log 2>&1; tail -4 gpurun_out/s19_pytest.log
timeout 200 python tools/bench_chains.py --which tx --steps 10 > gpurun_out/s19_tx.json 2>&1; tail -1 gpurun_out/s19_tx.json | cut -c1-220
