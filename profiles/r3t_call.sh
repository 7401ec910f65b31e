#!/bin/bash
# final validation of the round: exactly what the driver runs at round end on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r3t_tests.log 2>&1
tail -3 gpurun_out/r3t_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r3t_smoke.log
timeout 600 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3t_bench.json 2> gpurun_out/r3t_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3t_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step','clocks')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['frac'], 'e2e', l['e2e']['value'], l['e2e']['h2d_bytes_per_step'])
print({k:(round(v['us_per_launch'],1), round(v['frac'],3)) for k,v in l['roofline']['gemm'].items()})
print(l['cpu_baseline']['value'], (l.get('other_configs') or {}).get('config3',{}).get('ms_per_step'))
PY
tail -2 gpurun_out/r3t_bench.err | cut -c1-200
