mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/s25_pytest.log 2>&1; tail -3 gpurun_out/s25_pytest.log
timeout 300 python tools/bench_chains.py --which tx,chan,rx,q15 --steps 10 > gpurun_out/s25_chains.json 2>&1; cut -c1-200 gpurun_out/s25_chains.json
timeout 200 python tools/bench_chains.py --which q15 --steps 5 --rx-channels 8192 --seconds 4 > gpurun_out/s25_q15_8192.json 2>&1; tail -1 gpurun_out/s25_q15_8192.json | cut -c1-200
for w in tx q15; do bash tools/profile_chain.sh r02 $w 2 > /dev/null 2>&1; done
