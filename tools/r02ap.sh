#!/bin/bash
# 2-GPU box: whole GPU suite (incl. tests/test_gpu_multi.py) + the sharded parity check + a 2-GPU bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ap_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ap_pytest.log
tail -4 gpurun_out/r02ap_pytest.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 29519 tools/sharded_check.py 192 --train" > gpurun_out/r02ap_sharded_check_n2.json 2> gpurun_out/r02ap_sharded_check_n2.err; echo "sharded check exit $?"
tail -c 600 gpurun_out/r02ap_sharded_check_n2.json; echo
timeout 900 bash -c "$(declare -f run); run 29523 bench.py --gpus 2 --steps 10 --warmup 3 --quick" > gpurun_out/r02ap_bench_n2.json 2> gpurun_out/r02ap_bench_n2.err; echo "bench exit $?"
tail -c 400 gpurun_out/r02ap_bench_n2.json
