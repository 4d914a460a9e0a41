import csv,sys
from collections import defaultdict
lines=open(sys.argv[1]).read().splitlines()
start=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[start:]))
d=defaultdict(list)
for r in rows: d[r['Kernel Name'].split('(')[0]].append(float(r['Metric Value'])/1e6)
tot=sum(sum(v)/len(v) for k,v in d.items() if 'reset' not in k)
for k,v in d.items(): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v):8.3f} ms  {100*sum(v)/len(v)/tot:5.1f}%")
