import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
data=rows[1:]
names=[r[ik] for r in data]
idx=[i for i,n in enumerate(names) if 'project_fwd' in n]
which=int(sys.argv[2]) if len(sys.argv)>2 else -1
s=idx[which]; e=idx[which+1] if which+1<0 and which+1 < len(idx) else len(data)
tot=0
for r in data[s:e]:
    v=float(r[iv].replace(',',''))/1000; tot+=v
    print(f"{v:9.1f}  {r[ik][:100]}")
print("total",tot)
