#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_xt.py tests/test_zz5_gpu_prefetch.py tests/test_gpu_loader.py tests/test_gpu_graphed.py tests/test_gpu_edge_cases.py -q -p no:cacheprovider > gpurun_out/r3s_tests.log 2>&1
tail -4 gpurun_out/r3s_tests.log
timeout 600 python bench.py --no-other-configs --no-cpu-baseline > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3s_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step')}, 'e2e', l['e2e']['value'], l['e2e']['h2d_bytes_per_step'], l['e2e']['eager_serial']['value'], l['e2e']['how'][:60])
PY
tail -2 gpurun_out/r3s_bench.err | cut -c1-200
