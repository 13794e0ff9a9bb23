#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/test_edge2_fwd.py 2>&1 | tail -20
python tools/tc_dbg2.py > gpurun_out/plain_dbg.log 2>&1 || exit 1
run() {
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_tc_select --csv --log-file gpurun_out/dbgv.csv python tools/tc_dbg2.py > /dev/null 2>&1
  echo "$* : $(grep gpu__time_duration gpurun_out/dbgv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
}
run FS_TC_DBG=15
run FS_TC_DBG=31
run FS_TC_DBG=47
run FS_TC_DBG=63
run FS_TC_DBG=127
