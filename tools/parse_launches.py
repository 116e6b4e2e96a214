import csv, re, sys
def parse(f, marker="conv_first"):
    lines=[l for l in open(f) if not l.startswith("==")]
    rows=[]
    for r in csv.DictReader(lines):
        if r.get("Metric Name")=="gpu__time_duration.sum":
            v=float(r["Metric Value"]); u=r["Metric Unit"]
            v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
            rows.append((int(r["ID"]), re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","").replace("frcnn::",""), v, r["Grid Size"]))
    idx=[i for i,r in enumerate(rows) if marker in r[1]]
    step=rows[idx[-3]:idx[-2]]
    step=[r for r in step if "at::" not in r[1]]
    return step
if __name__=="__main__":
    step=parse(sys.argv[1], sys.argv[2] if len(sys.argv)>2 else "conv_first")
    tot=sum(r[2] for r in step)
    print("launches",len(step),"total us %.1f"%tot)
    for r in step: print("%-50s %8.1f  %s"%(r[1][:50],r[2],r[3]))
