#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_pair.py tests/test_gpu_edge_cases.py tests/test_gpu_a2gnn.py tests/test_gpu_graph.py -q -p no:cacheprovider > gpurun_out/r3g_tests.log 2>&1
tail -3 gpurun_out/r3g_tests.log
timeout 600 python bench.py --no-other-configs > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3g_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['frac'], (l['roofline'].get('batched') or {}).get('us_per_launch'), 'e2e', l['e2e']['value'])
g=l['roofline'].get('gemm') or {}
print({k:(v['us_per_launch'], round(v['frac'],3)) for k,v in g.items()})
PY
tail -2 gpurun_out/r3g_bench.err | cut -c1-200
