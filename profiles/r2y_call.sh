#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/probe_xt_clocks.py 2>&1 | grep -v Warning | tee gpurun_out/r2y_clocks.log
