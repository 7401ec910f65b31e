#!/bin/bash
mkdir -p gpurun_out
{ for m in 16 12 0; do GDA_SPMM_UNW=$m python profiles/bench_spmm_structure.py; done; } 2>&1 | tee gpurun_out/r2d_structure.log
