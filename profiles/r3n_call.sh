#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29691 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3n_bench8.json 2> gpurun_out/r3n_bench8.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3n_bench8.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], 'e2e', l['e2e']['value'], l['e2e']['how'][:80], l['config']['issue'])
print(json.dumps(l.get('other_configs'))[:1800])
PY
grep -v "Warning\|warn\|run_backward\|^\*\|OMP_NUM" gpurun_out/r3n_bench8.err | tail -6 | cut -c1-300
