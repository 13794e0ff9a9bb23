#!/bin/bash
# Round-2 profiles: launch list of the training step (eager: every kernel is a launch) and full captures of the top kernels.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
python tools/run_knn_tc.py > gpurun_out/plain_tc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'knn_tc_select|knn_tc_finalize' -s 4 -c 4 -o gpurun_out/r02_knn_tc -f python tools/run_knn_tc.py > gpurun_out/ncu_tc.log 2>&1
python tools/run_edge2.py fused > gpurun_out/plain_e2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'edge2_fwd|edge2_bwd_kernel' -s 2 -c 2 -o gpurun_out/r02_edge2 -f python tools/run_edge2.py fused > gpurun_out/ncu_e2.log 2>&1
python tools/run_gather.py > gpurun_out/plain_gather.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:edgeconv_gather_smem -s 1 -c 1 -o gpurun_out/r02_gather -f python tools/run_gather.py > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bench_launches.csv
python tools/torch_trace.py 1 > gpurun_out/plain_heads.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'pool_gemm|pool_lin|colsum|pool_stats' -c 8 -o gpurun_out/r02_heads -f python tools/torch_trace.py 1 > gpurun_out/ncu_heads.log 2>&1
ls -la gpurun_out/r02_heads.ncu-rep
