#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py tests/test_gpu_pair.py tests/test_gpu_loader.py tests/test_zz3_gpu_critic.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2e_tests.log
{ for m in 16 12 0; do GDA_SPMM_UNW=$m python profiles/bench_spmm_structure.py; done; } 2>&1 | tee gpurun_out/r2e_structure.log
