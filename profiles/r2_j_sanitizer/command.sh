#!/bin/bash
# hygiene: compute-sanitizer racecheck on the aggregation kernels (segment-reduction counters,
# shared-memory staging), synccheck on the tcgen05 GEMMs (mbarrier / bar.sync use), memcheck on the tile-packed GEMMs
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_pair.py -q -x -p no:cacheprovider -k "not scale and not large and not config" > gpurun_out/r3d_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r3d_racecheck.log; tail -6 gpurun_out/r3d_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm_xt.py tests/test_gpu_gemm_tc.py -q -x -p no:cacheprovider -k "not 20000 and not 4100 and not 19000" > gpurun_out/r3d_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/r3d_synccheck.log; tail -6 gpurun_out/r3d_synccheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gemm_xt.py -q -x -p no:cacheprovider -k "not 20000 and not 4100 and not 19000" > gpurun_out/r3d_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r3d_memcheck.log; tail -6 gpurun_out/r3d_memcheck.log
