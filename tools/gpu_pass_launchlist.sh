#!/bin/bash
# launch list of the training step only (eager: every kernel is a launch), after the same command ran clean without ncu
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/r02_bench_launches.csv
