#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/test_edge2_bwd.py > gpurun_out/edge2_bwd.log 2>&1; echo "edge2 rc=$?"; tail -70 gpurun_out/edge2_bwd.log
