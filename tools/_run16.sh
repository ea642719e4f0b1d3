set -x
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_C3_n8.jsonl 2> gpurun_out/r02_bench_C3_n8.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_C3_n8.err
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --config C4 --steps 2 --warmup 3 > gpurun_out/r02_bench_C4_n8.jsonl 2> gpurun_out/r02_bench_C4_n8.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_C4_n8.err
