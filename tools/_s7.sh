mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/s7_pytest.log 2>&1; tail -30 gpurun_out/s7_pytest.log
