#!/bin/bash
# Round-1 session-2 GPU check: new work-list aggregation kernel (A/B against k_spmm_rows), CUDA-graph step.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py tests/test_gpu_graph.py tests/test_gpu_graphed.py tests/test_gpu_a2gnn.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r1e_tests.log
for v in 0 4 12 8; do
  WIDE_ONLY=1 GDA_SPMM_TASKS=$v python profiles/bench_spmm.py 2>&1 | tail -3 | tee -a gpurun_out/r1e_spmm_ab.log
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e 2>gpurun_out/r1e_bench_graph.err | tee gpurun_out/r1e_bench_graph.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-cuda-graph 2>gpurun_out/r1e_bench_eager.err | tee gpurun_out/r1e_bench_eager.json
tail -5 gpurun_out/r1e_bench_graph.err
