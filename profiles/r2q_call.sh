#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_xt.py tests/test_gpu_a2gnn.py tests/test_gpu_graphed.py tests/test_zz5_gpu_prefetch.py tests/test_gpu_loader.py -q -p no:cacheprovider > gpurun_out/r2q_tests.log 2>&1
tail -8 gpurun_out/r2q_tests.log
timeout 600 python bench.py --no-other-configs > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2q_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['frac'], 'e2e', l['e2e'])
PY
tail -3 gpurun_out/r2q_bench.err | cut -c1-300
