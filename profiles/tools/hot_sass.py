import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.006
want=sys.argv[3] if len(sys.argv)>3 else ""
i=0
while i < len(rows):
    if rows[i] and rows[i][0]=="Kernel Name":
        name=rows[i][1]; h=rows[i+1]; j=i+2
        data=[]
        while j<len(rows) and not (rows[j] and rows[j][0]=="Kernel Name"):
            if len(rows[j])>10: data.append(rows[j])
            j+=1
        if want in name:
            ia=h.index('Instructions Executed'); isrc=h.index('Source'); ith=h.index('Avg. Threads Executed')
            tot=sum(int(r[ia]) for r in data)
            print("==",name[:80]); print("total",tot,"sass lines",len(data))
            for idx,r in enumerate(data):
                c=int(r[ia])
                if c>tot*thr: print(f"{idx:4d} {c:9d} {c/tot*100:5.1f}% thr={r[ith]:>5s}  {r[isrc][:100]}")
        i=j
    else: i+=1
