#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 tests/dist_check.py > gpurun_out/r2j_dist.log 2>&1
grep -n "spmm k=\|^step\|graphed\|GRADE\|AdaGCN\|DIST_CHECK\|Error" gpurun_out/r2j_dist.log | head -40
timeout 420 $TR --master-port 29633 bench.py --config 4 --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2j_config4_n2.json 2> gpurun_out/r2j_config4_n2.err
cat gpurun_out/r2j_config4_n2.json; tail -3 gpurun_out/r2j_config4_n2.err | cut -c1-300
