set -x
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_gputests_dist.log 2>&1; tail -8 gpurun_out/r02_gputests_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C3_n2.jsonl 2> gpurun_out/r02_bench_C3_n2.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_C3_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --config C2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C2_n2.jsonl 2> gpurun_out/r02_bench_C2_n2.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_C2_n2.err
