#!/bin/bash
# per-op device timestamps of the mg_cycle_kernel launches (whole cycles or the hybrid path's segments) at the bench size
mkdir -p gpurun_out; rm -f gpurun_out/cycle_trace.csv
FLIPB200_TRACE_CYCLE=gpurun_out/cycle_trace.csv timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err
python - <<'PY'
import csv, collections
rows=[r for r in csv.DictReader(open("gpurun_out/cycle_trace.csv")) if r["k"] != "k"]
names={0:"zero_red",1:"red",2:"black",3:"resid_restrict",4:"prolong",5:"coarse_cg"}
agg=collections.OrderedDict()
for r in rows:
    key=(int(r["level"]), names[int(r["op"])], int(r["leaves"]), int(r.get("dofs",0)), int(r.get("compact",0)))
    a=agg.setdefault(key,[0,0]); a[0]+=1; a[1]+=int(r["ns"])
tot=sum(int(r["ns"]) for r in rows)
print("total us", tot/1e3, "ops", len(rows))
for (lv,op,n,nd,cp),(c,ns) in sorted(agg.items()): print(f"level {lv} ({n} leaves, {nd} dofs, compact={cp}) {op:15s} x{c:4d}  avg {ns/c/1e3:8.2f} us  total {ns/1e3:9.1f} us")
PY
