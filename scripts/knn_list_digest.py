import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; order=[]
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d["Metric Name"]!="gpu__time_duration.sum": continue
        v=float(d["Metric Value"].replace(",","")); u=d["Metric Unit"]
        v = v/1e3 if u in ("ns","nsecond") else v
        order.append((d["Kernel Name"].split("(")[0][:40], v))
idx=[i for i,(n,v) in enumerate(order) if "knn_tc_scan" in n and v>300]
i=idx[-1]
print(" | ".join(f"{n.replace('void ','')[:18]} {v:.0f}" for n,v in order[i-1:i+2]))
