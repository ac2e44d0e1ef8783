#!/bin/bash
OBVHS_TRACE=1 timeout 300 python scripts/trace_build.py terrain 2>&1 | grep -E "ploc_|build_ploc|reinsertion_optimize|bvh2_to_cwbvh|total" | tail -9
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ploc_|onesweep|gather|morton|leaf_init" -c 200 --csv --log-file gpurun_out/ncu_r2m_ploc.csv python scripts/trace_build.py terrain > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/ncu_r2m_ploc.csv') if not l.startswith('==')]
per=collections.OrderedDict()
for r in csv.DictReader(rows):
    k=(r['ID'], r['Kernel Name'].split('(')[0][-40:])
    v=float(r['Metric Value'].replace(',',''))
    per.setdefault(k,{})[r['Metric Name']]=(v, r['Metric Unit'])
# last build: take the last third
items=list(per.items())
n=len(items)//3
for (i,name),m in items[-n:]:
    t=m['gpu__time_duration.sum']; rd=m.get('dram__bytes_read.sum',(0,'')); wr=m.get('dram__bytes_write.sum',(0,''))
    print(f"{name:42s} {t[0]:10.1f} {t[1]:4s} rd {rd[0]:8.1f} {rd[1]:6s} wr {wr[0]:8.1f} {wr[1]}")
PY
