#!/bin/bash
mkdir -p gpurun_out
python tools/tc_dbg2.py > gpurun_out/plain_dbg.log 2>&1 || exit 1
for d in 0 1 8 9 2 3 4 6 7 15; do
  FS_TC_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_tc_select --csv --log-file gpurun_out/dbg_$d.csv python tools/tc_dbg2.py > /dev/null 2>&1
  echo "dbg=$d $(grep gpu__time_duration gpurun_out/dbg_$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
