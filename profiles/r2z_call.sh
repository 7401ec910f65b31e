#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_xt.py -q -p no:cacheprovider > gpurun_out/r2z_xt_tests.log 2>&1
tail -5 gpurun_out/r2z_xt_tests.log
GDA_XT_MT1=1 timeout 300 python -m pytest tests/test_gpu_gemm_xt.py -q -p no:cacheprovider 2>&1 | tail -2
for m in 0 1; do
GDA_XT_MT1=$m timeout 300 python profiles/bench_gemm_xt.py 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('MT1=$m', 'fwd', round(l['xt_fwd']['us_median'],1), 'dw', round(l['xt_dw']['us_median'],1), 'dense', round(l['dense_fwd']['us_median'],1), round(l['dense_dw']['us_median'],1), 'dw_rel_diff', l['dw_rel_diff'])"
done | tee gpurun_out/r2z_mt.log
