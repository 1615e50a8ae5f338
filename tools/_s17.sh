mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/s17_pytest.log 2>&1; tail -5 gpurun_out/s17_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err; tail -c 3000 gpurun_out/s17_bench.json; tail -5 gpurun_out/s17_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/s17_bench_ref.json 2> gpurun_out/s17_bench_ref.err; tail -c 1200 gpurun_out/s17_bench_ref.json; tail -4 gpurun_out/s17_bench_ref.err
timeout 600 python tools/bench_feeder.py --channels 1024 --chunks 30 > gpurun_out/s17_feeder.json 2>&1; tail -9 gpurun_out/s17_feeder.json
