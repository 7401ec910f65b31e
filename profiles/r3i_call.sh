#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/profile_config5_host.py > gpurun_out/r3i_config5_host.log 2>&1
head -120 gpurun_out/r3i_config5_host.log | cut -c1-180
