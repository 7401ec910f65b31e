#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 tests/dist_check.py > gpurun_out/r2k_dist.log 2>&1
grep -n "spmm k=\|^step\|graphed\|GRADE\|AdaGCN\|DIST_CHECK\|Error\|error" gpurun_out/r2k_dist.log | head -50
timeout 420 $TR --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench2.json 2> gpurun_out/r2k_bench2.err
tail -c 1800 gpurun_out/r2k_bench2.json; grep -v "Warning\|warn\|run_backward" gpurun_out/r2k_bench2.err | tail -5 | cut -c1-300
GDA_DIST_MODE=peer timeout 420 $TR --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r2k_bench2_peer.json 2> gpurun_out/r2k_bench2_peer.err
tail -c 700 gpurun_out/r2k_bench2_peer.json
