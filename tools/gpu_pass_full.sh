#!/bin/bash
# full pass: all GPU tests, smoke, every bench workload, the reference arm
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in train configC static40 chamfer inference pointtransformer; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "bench $w failed"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d.get("ms_per_step"), d.get("value"), d.get("unit"), "e2e", d.get("e2e", {}).get("value"), "launches", d.get("gpu_launches"))
except Exception as e:
    print("$w: no line", e)
PY
done
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
