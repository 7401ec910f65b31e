#!/bin/bash
mkdir -p gpurun_out
{ for sg in 64 32 16; do for m in 12 0; do CASES=real,len_cap64 GDA_SEG=$sg GDA_SPMM_UNW=$m python profiles/bench_spmm_structure.py; done; done; } 2>&1 | tee gpurun_out/r2g_seg.log
python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2g_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 3000 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
