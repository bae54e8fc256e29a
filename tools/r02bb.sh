#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 900 bash -c "$(declare -f run); run 29527 bench.py --gpus 2 --steps 20 --warmup 5" > gpurun_out/r02bb_bench_n2_full.json 2> gpurun_out/r02bb_bench_n2_full.err; echo "bench exit $?"
tail -3 gpurun_out/r02bb_bench_n2_full.err
python - <<'PY'
import json
for l in open('gpurun_out/r02bb_bench_n2_full.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('N', d['n_gpus'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'train', d.get('train_step', {}).get('images_per_s'), 'strong', {k: round(v['ms_per_forward'], 3) for k, v in (d.get('strong_scaling') or {}).items()})
PY
