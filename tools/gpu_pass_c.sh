#!/bin/bash
mkdir -p gpurun_out
python tools/run_knn_tc.py > gpurun_out/plain_tc.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv python tools/run_knn_tc.py > gpurun_out/ncu_tc1.log 2>&1
python tools/run_knn_tc.py > gpurun_out/plain_tc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_tc_select -s 2 -c 2 -o gpurun_out/r02_tc_select -f python tools/run_knn_tc.py > gpurun_out/ncu_tc2.log 2>&1
ls -la gpurun_out | tail -5
