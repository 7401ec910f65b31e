#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_zz7_gpu_multi.py -q -p no:cacheprovider > gpurun_out/r3c_multi_tests.log 2>&1
tail -4 gpurun_out/r3c_multi_tests.log
timeout 420 $TR --master-port 29652 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3c_bench2.json 2> gpurun_out/r3c_bench2.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3c_bench2.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['kernel'][:40], 'e2e', l['e2e']['value'], l['config']['issue'], l['config']['parallelism'][:200])
print(json.dumps(l['roofline'].get('gemm'))[:600])
PY
grep -v "Warning\|warn\|run_backward\|^\*\|OMP_NUM" gpurun_out/r3c_bench2.err | tail -5 | cut -c1-300
