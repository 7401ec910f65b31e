#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r3a_tests.log 2>&1
tail -6 gpurun_out/r3a_tests.log
timeout 600 python bench.py --no-other-configs > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3a_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['frac'], 'e2e', l['e2e']['value'], l['e2e']['h2d_bytes_per_step'], l['e2e'].get('eager_serial'))
PY
tail -3 gpurun_out/r3a_bench.err | cut -c1-300
