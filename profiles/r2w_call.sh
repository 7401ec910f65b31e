#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
  GDA_XT_DBG=$d timeout 300 python profiles/bench_gemm_xt.py 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg=$d', 'fwd', round(l['xt_fwd']['us_median'],1), 'dw', round(l['xt_dw']['us_median'],1), 'dense', round(l['dense_fwd']['us_median'],1), round(l['dense_dw']['us_median'],1))"
done | tee gpurun_out/r2w_dbg.log
