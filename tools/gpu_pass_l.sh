#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/test_edge2_fwd.py 2>&1 | tail -14
timeout 600 python tools/test_edge2_bwd.py > gpurun_out/edge2_bwd.log 2>&1; echo "edge2 rc=$?"; grep -A12 "B=4 N=2048 k=20 train=True" gpurun_out/edge2_bwd.log; grep -A8 "B=2 N=1024 k=40 train=True" gpurun_out/edge2_bwd.log; grep -A8 "B=2 N=256 k=8 train=True" gpurun_out/edge2_bwd.log; tail -3 gpurun_out/edge2_bwd.log
python tools/run_edge2.py fused > gpurun_out/plain_e2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e2_launches.csv python tools/run_edge2.py fused > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/e2_launches.csv 14 3
