#!/bin/bash
mkdir -p gpurun_out
python tools/run_edge2.py fused > gpurun_out/plain_e2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e2_launches.csv python tools/run_edge2.py fused > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/e2_launches.csv 30 3
python tools/run_edge2.py bf16 > gpurun_out/plain_e2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e2old_launches.csv python tools/run_edge2.py bf16 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/e2old_launches.csv 16 3
