#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/profile_config5_host.py > gpurun_out/r3m_config5_host.log 2>&1
head -3 gpurun_out/r3m_config5_host.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r3m_tests.log 2>&1
tail -4 gpurun_out/r3m_tests.log
timeout 300 python bench.py --no-other-configs --no-cuda-graph --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eager', l['value'], l['ms_per_step'], l['host_issue_ms_per_step'])"
