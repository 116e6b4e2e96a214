#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2500 --csv --log-file gpurun_out/launches_train_b8.csv python bench.py --workload train --batch 8 --steps 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, re, collections
lines=[l for l in open('gpurun_out/launches_train_b8.csv') if not l.startswith('==')]
rows=[]
for r in csv.DictReader(lines):
    if r.get('Metric Name')=='gpu__time_duration.sum':
        v=float(r['Metric Value']); u=r['Metric Unit']
        v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
        rows.append((re.sub(r'\(.*','',r['Kernel Name']).replace('void ','').replace('frcnn::',''), v, r['Grid Size']))
idx=[i for i,r in enumerate(rows) if 'conv_first' in r[0]]
step=rows[idx[-1]:]
step=[r for r in step if 'at::' not in r[0]]
tot=sum(r[1] for r in step)
print('launches', len(step), 'total us %.1f'%tot)
agg=collections.OrderedDict()
for k,v,g in step:
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%-46s %4d %9.1f %5.1f%%'%(k[:46],n,t,100*t/tot))
PY
