#!/bin/bash
# quick pass: the given test files, the bench line, the kernel-level trace
python -m pytest ${TESTS:-tests/test_gpu_dense.py tests/test_gpu_measured_configs.py} -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 200 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'])"
python tools/torch_trace.py ${NK:-45} 2>&1 | grep -A${NK:-45} "^---- kernels"
