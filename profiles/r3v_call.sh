#!/bin/bash
# last call of the round: the committed tree (clean rebuild of libgda.so) as the driver runs it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r3v_tests.log 2>&1
tail -2 gpurun_out/r3v_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r3v_bench.json') if x.startswith('{')][-1])
print({k:l.get(k) for k in ('value','ms_per_step')}, 'e2e', l['e2e']['value'], l['e2e']['h2d_bytes_per_step'], 'agg', l['roofline']['us_per_launch'])
PY
