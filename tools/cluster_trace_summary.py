import csv, collections, sys
lines=open(sys.argv[1]).read().splitlines()
print(lines[0])
rows=list(csv.DictReader(lines[1:]))
print(len(rows), "ops; total us", sum(int(r['ns']) for r in rows)/1e3)
agg=collections.defaultdict(lambda:[0,0,0,0,0,0])
for r in rows:
    k=(r['level'],r['op'],r['leaves'],r['dofs']); agg[k][0]+=1; agg[k][1]+=int(r['ns']); agg[k][2]+=int(r.get('work_ns',0)); agg[k][3]+=int(r.get('cyc_dispatch',0)); agg[k][4]+=int(r.get('cyc_work',0)); agg[k][5]+=int(r.get('cyc_barrier',0))
names={0:'ZR',1:'RED',2:'BLK',3:'RR',4:'PR',5:'CG'}
for k,v in sorted(agg.items()): print(k[0], names[int(k[1])], 'leaves',k[2],'dofs',k[3], 'count',v[0], 'avg us', round(v[1]/v[0]/1e3,2), 'work(rank1) us', round(v[2]/v[0]/1e3,2), 'total', round(v[1]/1e3,1), 'cyc disp/work/barrier', v[3]//v[0], v[4]//v[0], v[5]//v[0])
