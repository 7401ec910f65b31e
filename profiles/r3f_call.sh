#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29671 tests/dist_check.py > gpurun_out/r3f_dist4.log 2>&1
grep -n "halo spmm nb\|graphed dist step 3\|DIST_CHECK\|Error\|error" gpurun_out/r3f_dist4.log | head -20
timeout 420 $TR --master-port 29672 bench.py --gpus 4 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r3f_bench4.json 2> gpurun_out/r3f_bench4.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3f_bench4.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['launches_per_step'], (l['roofline'].get('batched') or {}).get('us_per_launch'), 'e2e', l['e2e']['value'], l['config']['issue'])
PY
grep -v "Warning\|warn\|run_backward\|^\*\|OMP_NUM" gpurun_out/r3f_bench4.err | tail -5 | cut -c1-300
