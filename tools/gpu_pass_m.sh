#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
for wl in configC static40 chamfer pointtransformer inference; do
  timeout 300 python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
done
