#!/bin/bash
# per-kernel device times of the trainable-path ops (tools/head_bench.py) from an ncu launch list
mkdir -p gpurun_out
timeout 100 python tools/head_bench.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/head_launches.csv python tools/head_bench.py > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/head_launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for d in csv.DictReader(lines):
    try: v=float(d['Metric Value'].replace(',',''))
    except: continue
    u=d['Metric Unit']; v=v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    n=re.sub(r'\(.*','',d['Kernel Name']).replace('vlmclip::<unnamed>::','')[:50]
    agg.setdefault(n,[]).append(v)
for n,a in agg.items():
    if n.startswith('void at::'): continue
    a2=sorted(a); print(f"{n:40s} n={len(a):3d} median {a2[len(a2)//2]:7.1f} us  min {a2[0]:7.1f}")
PY
