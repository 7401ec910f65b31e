#!/bin/bash
# Round 2, call c: factored unit-weight aggregation chain (k_spmm_unw) -- tests, A/B micro-benchmark, bench, ncu.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2c_tests.log
{
  for m in 16 12 0; do GDA_SPMM_UNW=$m python profiles/bench_spmm_unw.py; done
  for m in 16 0; do NB=2 GDA_SPMM_UNW=$m python profiles/bench_spmm_unw.py; done
  for m in 16 0; do N=1000000 E=10000000 GDA_SPMM_UNW=$m python profiles/bench_spmm_unw.py; done
} 2>&1 | tee gpurun_out/r2c_unw_ab.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 5000 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_spmm_unw -s 20 -c 2 -f -o gpurun_out/r2c_unw_full \
    python profiles/bench_spmm_unw.py > gpurun_out/r2c_ncu_unw.log 2>&1
ncu -i gpurun_out/r2c_unw_full.ncu-rep --page raw --csv > gpurun_out/r2c_unw_full.raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
