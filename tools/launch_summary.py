"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv [top] [steps]"""
import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
nsteps = float(sys.argv[3]) if len(sys.argv) > 3 else 9
rows=[]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        us = v/1000 if u.startswith('n') else (v if u.startswith('u') else v*1000)
        rows.append((int(row['ID']), row['Kernel Name'], us))
agg=collections.defaultdict(lambda:[0,0.0])
for _,k,us in rows:
    k=re.sub(r'\(.*','',k); k=re.sub(r'^void ','',k); k=re.sub(r'<unnamed>::','',k)[:80]
    agg[k][0]+=1; agg[k][1]+=us
tot=sum(v[1] for v in agg.values())
print('%d launches, total %.1f us, per step %.1f us (%g steps)'%(len(rows), tot, tot/nsteps, nsteps))
for k,(n,us) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print('%5.1f%% %8.1f us/step  n/step=%5.1f  avg %8.1f us  %s'%(100*us/tot, us/nsteps, n/nsteps, us/n, k))
