"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py per kernel name.
Steps are delimited by `adam_kernel` (the last kernel of a training step); the last `steps` complete steps are averaged.
    python tools/launch_summary.py gpurun_out/r02_bench_launches.csv [steps] > profiles/..._summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
want = int(sys.argv[2]) if len(sys.argv) > 2 else 4
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        us = v / 1000 if u.startswith('n') else (v if u.startswith('u') else v * 1000)
        rows.append((row['Kernel Name'], us))
ends = [i for i, (k, _) in enumerate(rows) if 'adam_kernel' in k]
assert len(ends) > want, "not enough training steps in the launch list"
first, last = ends[-want - 1] + 1, ends[-1] + 1
sel = rows[first:last]


def short(k):
    k = re.sub(r'^void ', '', k)
    k = re.sub(r'<unnamed>::', '', k)
    k = re.sub(r'\(anonymous namespace\)::', '', k)
    m = re.match(r'([A-Za-z0-9_:]+(<[^(]*>)?)', k)
    return (m.group(1) if m else k)[:90]


OWN = re.compile(r'^(knn|tc_|edge|bn_|colstats|pool_|reverse_|morton|adam|cat_cast|split_cast|colsum|multi_copy|softmax_scatter|fps|nn_points|chamfer)')
agg = collections.defaultdict(lambda: [0, 0.0])
cls = collections.defaultdict(lambda: [0, 0.0])
for k, us in sel:
    s = short(k)
    agg[s][0] += 1
    agg[s][1] += us
    c = 'own' if OWN.match(s) else ('gemm' if re.search(r'nvjet|cutlass|gemm|splitKreduce', s) else 'aten')
    cls[c][0] += 1
    cls[c][1] += us
tot = sum(v[1] for v in agg.values())
print('kernels per step: %d, sum of kernel times per step: %.1f us (%d steps averaged)' % (len(sel) / want, tot / want, want))
print("this library's kernels: %d per step, %.1f us (%.1f %%)" % (cls['own'][0] / want, cls['own'][1] / want, 100 * cls['own'][1] / tot))
print('library GEMMs (cuBLAS): %.1f us (%.1f %%); ATen glue: %.1f us (%.1f %%)' % (
    cls['gemm'][1] / want, 100 * cls['gemm'][1] / tot, cls['aten'][1] / want, 100 * cls['aten'][1] / tot))
print()
print(' share    us/step  n/step     avg us  kernel')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%5.1f%% %10.1f %7.1f %10.1f  %s' % (100 * us / tot, us / want, n / want, us / n, k))
