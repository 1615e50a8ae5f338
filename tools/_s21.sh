mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_chan64_f32.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/s21_pytest.log 2>&1; tail -4 gpurun_out/s21_pytest.log
timeout 200 python tools/bench_chains.py --which chan --steps 10 > gpurun_out/s21_chan.json 2>&1; tail -1 gpurun_out/s21_chan.json | cut -c1-220
