#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29682 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r3u_bench2.json 2> gpurun_out/r3u_bench2.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3u_bench2.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step')}, 'e2e', l['e2e'])
PY
grep -v "Warning\|warn\|run_backward\|^\*\|OMP_NUM" gpurun_out/r3u_bench2.err | tail -8 | cut -c1-300
