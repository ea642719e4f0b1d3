set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_g.log 2>&1; tail -15 gpurun_out/r02_gputests_g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C3_n2.jsonl 2> gpurun_out/r02_bench_C3_n2.err; echo rc=$?; tail -c 1500 gpurun_out/r02_bench_C3_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_bench_C3_ref_n2.jsonl 2> gpurun_out/r02_bench_C3_ref_n2.err; echo rc=$?; tail -c 500 gpurun_out/r02_bench_C3_ref_n2.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_C3_c.jsonl 2> gpurun_out/r02_bench_C3_c.err; echo rc=$?; tail -c 500 gpurun_out/r02_bench_C3_c.err
python __graft_entry__.py smoke 2>&1 | tail -3
