#!/bin/bash
# Round 2, call h (2 GPUs): multi-GPU parity inside pytest, the partitioned step as a CUDA graph, configs 3 / 4 / 5.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_zz7_gpu_multi.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2h_tests.log
timeout 420 $TR --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err
tail -c 2500 gpurun_out/r2h_bench2.json; tail -5 gpurun_out/r2h_bench2.err
timeout 420 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cuda-graph > gpurun_out/r2h_bench2_eager.json 2> gpurun_out/r2h_bench2_eager.err
tail -c 600 gpurun_out/r2h_bench2_eager.json
timeout 300 python bench.py --config 3 --steps 5 --warmup 2 > gpurun_out/r2h_config3.json 2> gpurun_out/r2h_config3.err
cat gpurun_out/r2h_config3.json; tail -3 gpurun_out/r2h_config3.err
timeout 420 $TR --master-port 29613 bench.py --config 4 --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2h_config4_n2.json 2> gpurun_out/r2h_config4_n2.err
cat gpurun_out/r2h_config4_n2.json; tail -3 gpurun_out/r2h_config4_n2.err
timeout 420 $TR --master-port 29614 bench.py --config 5 --gpus 2 --steps 10 --warmup 2 > gpurun_out/r2h_config5_n2.json 2> gpurun_out/r2h_config5_n2.err
cat gpurun_out/r2h_config5_n2.json; tail -3 gpurun_out/r2h_config5_n2.err
