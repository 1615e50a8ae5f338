mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/s18_bench2.json 2> gpurun_out/s18_bench2.err; tail -c 2500 gpurun_out/s18_bench2.json; tail -5 gpurun_out/s18_bench2.err
( timeout 600 python -m pytest tests/test_gpu_sharding_nccl.py -m gpu -q 2>&1 | tail -3 )
