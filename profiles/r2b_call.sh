#!/bin/bash
# Round 2, call b: full GPU suite after the default-path changes (graphed + staged fit), bench both arms.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2b_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 6000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err
cat gpurun_out/r2b_ref.json; tail -5 gpurun_out/r2b_ref.err
nproc; free -g | head -2
