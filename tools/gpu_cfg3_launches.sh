mkdir -p gpurun_out
CMD="python tools/cfg3_full.py 100000 31"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg3.csv $CMD > gpurun_out/ncu_cfg3.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg3.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    agg.setdefault(r[ki][:90],[]).append(float(r[vi].replace(',','')))
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:16]: print(f"{sum(v)/1e6:9.3f} ms {100*sum(v)/tot:5.1f}%  x{len(v):4d}  {k}")
PY
