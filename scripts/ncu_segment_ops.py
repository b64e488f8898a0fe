import csv, io, re, subprocess, sys, collections
rep, pat, seg = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--kernel-name-base", "demangled", "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[1]; ix = {k: i for i, k in enumerate(h)}
seen=set(); data=[]
for r in rows[2:]:
    if r and r[0]=="Kernel Name": break
    a=r[ix["Address"]]
    if a in seen: continue
    seen.add(a); data.append(r)
cur=0; ops=collections.Counter(); ex=collections.Counter()
for r in data:
    src=r[ix["Source"]]
    if cur==seg:
        m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src); op=m.group(2).split('.')[0] if m else '?'
        ops[op]+=1
        try: ex[op]+=float(r[ix["Instructions Executed"]])
        except: pass
    if re.search(r"\bBAR\.", src): cur+=1
print(sum(ops.values()), ops.most_common(25))
