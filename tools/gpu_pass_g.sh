#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -2 gpurun_out/pytest_knn.log
timeout 200 python tools/debug_knn_sorted.py 2>&1 | tail -12
python tools/tc_dbg2.py > gpurun_out/plain_dbg.log 2>&1 || exit 1
run() {
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_tc_select --csv --log-file gpurun_out/dbgv.csv python tools/tc_dbg2.py > /dev/null 2>&1
  echo "$* : $(grep gpu__time_duration gpurun_out/dbgv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
}
run FS_TC_DBG=0
run FS_TC_DBG=1
run FS_TC_DBG=15
