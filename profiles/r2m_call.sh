#!/bin/bash
# session 3 of round 2, first call: the whole GPU suite (no -x), smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1
tail -15 gpurun_out/r2m_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 600 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2m_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','host_issue_ms_per_step','clocks')}, 'agg us', l['roofline']['us_per_launch'], l['roofline']['frac'], 'e2e', l['e2e']['value'])
print(json.dumps(l.get('other_configs'))[:1500])
PY
tail -3 gpurun_out/r2m_bench.err | cut -c1-300
