mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_svd2048.csv python tools/kernel_bench.py --svd 2048 > gpurun_out/svdprof.log 2>&1
tail -2 gpurun_out/svdprof.log | cut -c1-300
python - <<PY
import csv, re, collections
lines=[l for l in open("gpurun_out/launches_svd2048.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
by=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    v=float(r["Metric Value"].replace(",",""))
    if r["Metric Unit"]=="ns": v/=1e3
    n=re.sub(r"\(.*","",r["Kernel Name"])[:60]
    by[n][0]+=1; by[n][1]+=v
tot=sum(v[1] for v in by.values())
print("total us", tot)
for k,v in sorted(by.items(), key=lambda kv:-kv[1][1])[:12]: print(k, v[0], round(v[1]), round(v[1]/v[0],1), round(v[1]/tot,3))
PY
