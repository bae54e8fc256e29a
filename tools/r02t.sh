#!/bin/bash
# 2-GPU: sharded parity with the peer-memory all-reduce (and with NCCL for comparison), then the bench line both ways
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 29519 tools/sharded_check.py 192 --train" > gpurun_out/r02t_sharded_check_n2.json 2> gpurun_out/r02t_sharded_check.err; echo "sharded check (peer) exit $?"
tail -c 1400 gpurun_out/r02t_sharded_check_n2.json; echo
NODE_B200_PEER_REDUCE=0 timeout 600 bash -c "$(declare -f run); run 29521 tools/sharded_check.py 192 --train" > gpurun_out/r02t_sharded_check_n2_nccl.json 2>> gpurun_out/r02t_sharded_check.err; echo "sharded check (nccl) exit $?"
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  e2e %.0f  %.3f ms/step  odeblock %.2f ms  strong %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['odeblock']['ms_per_step'], {k: round(v['ms_per_forward'], 3) for k, v in (d.get('strong_scaling') or {}).items()}))
"; }
for pr in 1 0 1 0; do
  echo "== peer reduce $pr"; NODE_B200_PEER_REDUCE=$pr timeout 600 bash -c "$(declare -f run); run 29523 bench.py --gpus 2 --steps 10 --warmup 3 --train-batch 0" 2>> gpurun_out/r02t_bench.err | show
done
tail -5 gpurun_out/r02t_sharded_check.err
