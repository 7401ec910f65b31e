#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_a2gnn.py tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_tdss.py tests/test_gpu_graphed.py -q -p no:cacheprovider -k "mmd or MMD or forward_model or step or graphed" > gpurun_out/r3o_tests.log 2>&1
tail -3 gpurun_out/r3o_tests.log
timeout 600 python bench.py --no-other-configs --no-cpu-baseline > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3o_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'agg us', l['roofline']['us_per_launch'], 'e2e', l['e2e']['value'])
PY
tail -2 gpurun_out/r3o_bench.err | cut -c1-200
