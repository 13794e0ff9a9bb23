#!/bin/bash
# Round-2 pass B: new tcgen05 kNN (feature + coordinates) - focused tests first, then everything.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -s -x > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 gpurun_out/pytest_knn.log
timeout 300 python tools/microbench_knn.py > gpurun_out/mb_knn.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_knn.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --no-eager-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
for wl in configC static40 inference; do
  timeout 300 python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
done
tail -c 400 gpurun_out/bench.log
