#!/bin/bash
# First GPU call of round 2 (one gpurun, one GPU, ~10 min of box time).  Before calling, HERE (no GPU needed):
#   make -C pygda_b200/csrc && make -C pygda_b200/csrc VARIANT=ldcg && make -C pygda_b200/csrc VARIANT=noalloc
#   make -C profiles/probes
# then:  gpurun --timeout 1200 -- 'bash profiles/r2_first_call.sh'
# 1. the GPU suite WITHOUT -x, so the round-1 tests written without a GPU (tests/test_zz_*) all report
# 2. smoke + the bench line (e2e.prefetch is the new key)
# 3. the aggregation A/Bs queued in profiles/probes/README.md: ping-pong vs read-only input, gather cache policy,
#    gather_probe6's feature-by-feature bisection of the probe-vs-kernel gap
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r2a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2a_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 5000 gpurun_out/r2a_bench.json
{
  echo "== default library, ping-pong chain";        WIDE_ONLY=1 python profiles/bench_spmm.py
  echo "== default library, read-only input";        SAME_X=1 WIDE_ONLY=1 python profiles/bench_spmm.py
  for v in ldcg noalloc; do
    if [ -f pygda_b200/libgda_$v.so ]; then
      echo "== libgda_$v.so, ping-pong chain";       GDA_LIB_PATH=$PWD/pygda_b200/libgda_$v.so WIDE_ONLY=1 python profiles/bench_spmm.py
      echo "== libgda_$v.so, read-only input";       GDA_LIB_PATH=$PWD/pygda_b200/libgda_$v.so SAME_X=1 WIDE_ONLY=1 python profiles/bench_spmm.py
    fi
  done
} 2>&1 | tee gpurun_out/r2a_spmm_ab.log
timeout 600 python profiles/bench_strurw.py 2>&1 | tee gpurun_out/r2a_strurw.jsonl
if [ -x profiles/probes/gather_probe6 ]; then
  timeout 300 profiles/probes/gather_probe6 2>&1 | tee gpurun_out/r2a_probe6.log
fi
ls -la gpurun_out
