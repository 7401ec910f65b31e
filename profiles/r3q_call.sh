#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_a2gnn.py tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_tdss.py -q -p no:cacheprovider -k "mmd or MMD or forward_model or step" > gpurun_out/r3q_tests.log 2>&1
tail -2 gpurun_out/r3q_tests.log
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:k_mmd_bwd --csv --log-file gpurun_out/r3q_mmd.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph --no-other-configs > gpurun_out/r3q.log 2>&1
grep "k_mmd_bwd" gpurun_out/r3q_mmd.csv | awk -F'","' '{print $13, $15}' | head -4
timeout 300 python bench.py --no-other-configs --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', l['value'], l['ms_per_step'])"
