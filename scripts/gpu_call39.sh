#!/bin/bash
set -u
O=gpurun_out/c39
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csc_|chunk_scan|sparse_grads" -c 60 --csv --log-file $O/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-parity --no-extras > $O/ncu.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(l for l in open("$O/launches_c2.csv") if l.startswith('"')))
h=rows[0]; ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); v = v/1000 if r[ui]=="ns" else (v*1000 if r[ui]=="ms" else v)
    a=agg.setdefault(r[ki][:60],[0,0.0]); a[0]+=1; a[1]+=v
for k,a in agg.items(): print(f"{k:62s} n={a[0]:3d} avg {a[1]/a[0]:8.1f} us")
PY
