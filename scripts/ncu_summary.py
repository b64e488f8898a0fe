"""Summarise an ncu report: per kernel duration / occupancy / dram bytes, stall mix and the
hottest SASS instructions (uses `ncu --page raw/source --csv`; works without a GPU)."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 8


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, "--kernel-name-base", "demangled", *a], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, rows = raw[0], raw[1], raw[2:]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for r in rows:
    name = r[hdr.index("Kernel Name")]
    if pat and not re.search(pat, name):
        continue
    print("=====", re.sub(r"void fcd::rt::fcd_kernel<fcd::|\(T1::Params.*", "", name))
    for k in want:
        if k in hdr:
            print(f"   {k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")

src_args = ["--page", "source", "--csv"] + (["-k", f"regex:{pat}"] if pat else [])
src = list(csv.reader(io.StringIO(ncu(*src_args))))
# split per kernel blocks ("Kernel Name" rows)
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h, data = b["rows"][0], b["rows"][1:]
    ix = {k: i for i, k in enumerate(h)}
    seen, uniq = set(), []
    for r in data:           # the page lists every instruction twice
        key = r[ix["Address"]]
        if key in seen:
            continue
        seen.add(key)
        uniq.append(r)
    data = uniq

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = collections.Counter({k: sum(f(r, k) for r in data) for k in stalls})
    ns = sum(f(r, "# Samples") for r in data)
    ni = sum(f(r, "Instructions Executed") for r in data)
    print("=====", re.sub(r"void fcd::rt::fcd_kernel<fcd::|\(T1::Params.*", "", b["name"]), "static", len(data), "samples", int(ns), "warp-inst", int(ni))
    print("   stalls:", ", ".join(f"{k[6:]} {100 * v / max(ns, 1):.0f}%" for k, v in tot.most_common(9)))
    byop = collections.Counter()
    ex = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2) if m else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "MUFU", "BAR")) else op.split(".")[0]
        byop[op] += f(r, "# Samples")
        ex[op] += f(r, "Instructions Executed")
    print("   by opcode (samples% / executed%):", ", ".join(f"{o} {100 * v / max(ns, 1):.0f}/{100 * ex[o] / max(ni, 1):.0f}" for o, v in byop.most_common(12)))
    hot = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:ntop]
    for i in hot:
        r = data[i]
        top = max(stalls, key=lambda k: f(r, k))
        print(f"   hot[{i}] {int(f(r, '# Samples')):6d} {top[6:]:10s} {r[ix['Source']][:60]}")
        for j in range(max(0, i - 3), i):
            print(f"            .. {data[j][ix['Source']][:60]}")
