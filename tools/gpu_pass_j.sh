#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/test_edge2_bwd.py > gpurun_out/edge2_bwd.log 2>&1; echo "edge2 rc=$?"; tail -75 gpurun_out/edge2_bwd.log
python tools/tc_dbg2.py > gpurun_out/plain_dbg.log 2>&1 || exit 1
run() {
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_tc_select --csv --log-file gpurun_out/dbgv.csv python tools/tc_dbg2.py > /dev/null 2>&1
  echo "$* : $(grep gpu__time_duration gpurun_out/dbgv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
}
run FS_TC_PROBE=0
run FS_TC_PROBE=1
run FS_TC_PROBE=3
