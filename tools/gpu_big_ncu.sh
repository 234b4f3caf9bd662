#!/bin/bash
mkdir -p gpurun_out
python tools/big_probe.py 30000 > gpurun_out/big_probe.txt 2>&1; tail -2 gpurun_out/big_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_big.csv python tools/big_probe.py 30000 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_big.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{ms:9.3f} ms  {n:4d}x  {k}")
PY
