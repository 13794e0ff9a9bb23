#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -2 gpurun_out/pytest_knn.log
timeout 300 python tools/microbench_knn.py > gpurun_out/mb_knn.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_knn.log
timeout 600 python tools/debug_layers.py > gpurun_out/debug_layers.log 2>&1; echo "debug rc=$?"; cat gpurun_out/debug_layers.log
python tools/run_knn_tc.py > gpurun_out/plain_tc.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tc_launches.csv python tools/run_knn_tc.py > gpurun_out/ncu_tc1.log 2>&1
