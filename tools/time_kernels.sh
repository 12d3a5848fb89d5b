#!/bin/bash
# warm-cache per-kernel durations of one mini-batch (ncu, serialized) -> stdout
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_ -s 130 -c 26 --csv --log-file /tmp/l.csv $B > /tmp/ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('/tmp/l.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
out=[]
for r in rows[14:26]:
    out.append(f"{r['Kernel Name'].split('(')[0][:28]}:{float(r['Metric Value'].replace(',',''))/1e3:.1f}")
print(" | ".join(out))
PY
