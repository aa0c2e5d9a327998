import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); im=hdr.index('Metric Name'); iid=hdr.index('ID')
d={}; order=[]
for r in rows[1:]:
    k=r[iid]
    if k not in d: d[k]={'name':r[ik]}; order.append(k)
    d[k][r[im]]=float(r[iv].replace(',',''))
ids=[k for k in order if 'project_fwd' in d[k]['name']]
s=order.index(ids[-2]); e=order.index(ids[-1])
tot=0
for k in order[s:e]:
    x=d[k]
    tot+=x.get('gpu__time_duration.sum',0)/1000
    if 'at::' in x['name']: continue
    print(f"{x.get('gpu__time_duration.sum',0)/1000:8.1f} us {x.get('smsp__inst_executed.sum',0)/1e6:7.2f} Minst  {x['name'][:90]}")
print("total",tot)
