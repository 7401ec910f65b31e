#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29621 tests/dist_check.py > gpurun_out/r2i_dist.log 2>&1
grep -v "^W\|UserWarning\|warn" gpurun_out/r2i_dist.log | tail -40
timeout 300 python -m pytest tests/test_zz5_gpu_prefetch.py tests/test_gpu_loader.py tests/test_gpu_dense.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -c 1500 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err | cut -c1-400
