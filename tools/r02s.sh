#!/bin/bash
# 2-GPU run of the bench (sharded forward + cfg3 training step with the native caller backward)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02s_bench_n2.json 2> gpurun_out/r02s_bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/r02s_bench_n2.err
head -c 600 gpurun_out/r02s_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/sharded_check.py 192 --train > gpurun_out/r02s_sharded_check_n2.json 2> gpurun_out/r02s_sharded_check.err; echo "sharded check exit $?"
tail -c 1500 gpurun_out/r02s_sharded_check_n2.json
