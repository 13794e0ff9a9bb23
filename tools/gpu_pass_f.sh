#!/bin/bash
mkdir -p gpurun_out
python tools/tc_dbg2.py > gpurun_out/plain_dbg.log 2>&1 || exit 1
run() {
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_tc_select --csv --log-file gpurun_out/dbgv.csv python tools/tc_dbg2.py > /dev/null 2>&1
  echo "$* : $(grep gpu__time_duration gpurun_out/dbgv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
}
run FS_TC_DBG=0
run FS_TC_DBG=0 FS_TC_NACC=1
run FS_TC_DBG=0 FS_TC_NACC=2
run FS_TC_DBG=0 FS_TC_STAGES=2
run FS_TC_DBG=0 FS_TC_1CTA=1
run FS_TC_DBG=15 FS_TC_NACC=1
run FS_TC_DBG=15 FS_TC_NACC=2
run FS_TC_DBG=15 FS_TC_1CTA=1
run FS_TC_DBG=1 FS_TC_1CTA=1
