#!/bin/bash
# A/B of the L1 row prefetch in k_spmm_tasks (GDA_SPMM_PF=0/1) on the config-2 target graph.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pair.py tests/test_gpu_spmm.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r1g_tests.log
for pf in 0 1; do
  echo "GDA_SPMM_PF=$pf" | tee -a gpurun_out/r1g_pf.log
  WIDE_ONLY=1 GDA_SPMM_PF=$pf python profiles/bench_spmm.py 2>&1 | tail -2 | tee -a gpurun_out/r1g_pf.log
  WIDE_ONLY=1 GDA_SPMM_PF=$pf GDA_SPMM_TASKS=8 python profiles/bench_spmm.py 2>&1 | tail -2 | tee -a gpurun_out/r1g_pf.log
done
