#!/bin/bash
# One measurement pass on a GPU box (gpurun --timeout 420 -- 'bash tools/gpu_pass.sh'): GPU tests, smoke, bench,
# ncu launch list and the two full captures that profiles/ summarises. Every step has its own hard timeout.
mkdir -p gpurun_out
timeout 90 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -2 gpurun_out/pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 120 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:edgeconv_gather_smem -s 1 -c 1 -o gpurun_out/full_gather -f python tools/run_gather.py > gpurun_out/ncu_gather.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:'knn_tc_select|knn_tc_finalize' -s 4 -c 2 -o gpurun_out/full_tc -f python tools/run_knn_tc.py > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out | tail -12
