#!/bin/bash
# N-GPU check (N = $1): sharded parity with the peer-memory all-reduce, then the bench line
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "N=$N; $(declare -f run); run 29519 tools/sharded_check.py $((96 * N)) --train" > gpurun_out/r02y_sharded_check_n$N.json 2> gpurun_out/r02y_sharded_check_n$N.err; echo "sharded check exit $?"
tail -c 1300 gpurun_out/r02y_sharded_check_n$N.json; echo
timeout 900 bash -c "N=$N; $(declare -f run); run 29523 bench.py --gpus $N --steps 10 --warmup 3" > gpurun_out/r02y_bench_n$N.json 2> gpurun_out/r02y_bench_n$N.err; echo "bench exit $?"
python - <<PY
import json
for l in open('gpurun_out/r02y_bench_n$N.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('N', d['n_gpus'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms/step %.3f' % d['ms_per_step'], 'train', d.get('train_step', {}).get('images_per_s'), d.get('train_step', {}).get('ms_per_step'), 'strong', {k: round(v['ms_per_forward'], 3) for k, v in (d.get('strong_scaling') or {}).items()})
PY
tail -3 gpurun_out/r02y_bench_n$N.err
