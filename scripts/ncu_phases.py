"""Per-phase share of warp-stall samples and executed instructions of one kernel in an ncu report:
the kernels are straight-line code per tile, phases are separated by BAR instructions."""
import csv, io, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--kernel-name-base", "demangled", "--page", "source", "--csv", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
seen = set(); data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    a = r[ix["Address"]]
    if a in seen: continue
    seen.add(a); data.append(r)
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
phases = [[0, 0.0, 0.0, {}]]
for r in data:
    src = r[ix["Source"]]
    ph = phases[-1]
    ph[0] += 1; ph[1] += f(r, "# Samples"); ph[2] += f(r, "Instructions Executed")
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            ph[3][k] = ph[3].get(k, 0) + f(r, k)
    if re.search(r"\bBAR\.", src): phases.append([0, 0.0, 0.0, {}])
ts = sum(p[1] for p in phases); ti = sum(p[2] for p in phases)
for n, p in enumerate(phases):
    top = sorted(p[3].items(), key=lambda kv: -kv[1])[:4]
    print(f"seg {n:2d}: static {p[0]:5d}  samples {100*p[1]/ts:5.1f}%  executed {100*p[2]/ti:5.1f}%  " +
          ", ".join(f"{k[6:]} {100*v/max(p[1],1):.0f}%" for k, v in top))
