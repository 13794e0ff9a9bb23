#!/bin/bash
# heads pass: dense tests, bench, kernel-level trace, then one ncu capture of the pooled-layer kernels
python -m pytest tests/test_gpu_dense.py tests/test_gpu_measured_configs.py -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'])"
python tools/torch_trace.py 45 2>&1 | grep -A50 "^---- kernels"
ncu --set full --clock-control none --import-source on -k regex:"pool_lin|colsum" -c 5 -o gpurun_out/r02_heads -f python tools/torch_trace.py 1 > gpurun_out/ncu_heads.log 2>&1
tail -3 gpurun_out/ncu_heads.log
