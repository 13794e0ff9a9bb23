set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edgeconv_gather_kernel -s 8 -c 2 -o gpurun_out/full_gather -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_gather.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'knn_tc_select_kernel|knn_tc_finalize_kernel' -s 8 -c 2 -o gpurun_out/full_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'knn3d_regs_kernel|edge3_bwd_kernel|edge_reduce_kernel' -s 12 -c 3 -o gpurun_out/full_misc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_misc.log 2>&1
ls -la gpurun_out
