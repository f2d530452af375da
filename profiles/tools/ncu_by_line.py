"""Joins an `ncu --page source --csv` export (per SASS instruction) with `nvdisasm -g` line info of the same cubin and
prints the hottest source lines: warp instructions executed, lanes per instruction, stall samples by reason.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all build/ga_kernels.cu.o && nvdisasm -g -c ga_kernels.sm_100a.cubin > lines.sass
    python profiles/tools/ncu_by_line.py src.csv lines.sass <mangled kernel name substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, lines_sass, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
addr2line = {}
cur = None
inside = False
for ln in open(lines_sass, errors="replace"):
    if ln.startswith("//-----") and ".text." in ln:
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    line = addr2line.get(a - base, ("?", 0))
    for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_lg", "stall_mio", "stall_math", "stall_no_inst", "stall_selected", "stall_dispatch"):
        v = float(r[col[k]] or 0)
        agg[line][k] += v
        tot[k] += v
print("total warp instr %.3g  thread instr %.3g  samples %d" % (tot["Instructions Executed"], tot["Thread Instructions Executed"], tot["# Samples"]))
print("%-22s %7s %6s %6s | %5s %5s %5s %5s %5s %5s %5s" % ("line", "inst%", "lanes", "smpl%", "long", "short", "wait", "brnch", "lg", "mio", "noins"))
for line, d in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    ie = d["Instructions Executed"]
    s = max(d["# Samples"], 1)
    print("%-22s %6.2f%% %6.1f %5.2f%% | %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f" % ("%s:%d" % line, 100 * ie / tot["Instructions Executed"], d["Thread Instructions Executed"] / max(ie, 1),
          100 * d["# Samples"] / tot["# Samples"], 100 * d["stall_long_sb"] / s, 100 * d["stall_short_sb"] / s, 100 * d["stall_wait"] / s, 100 * d["stall_branch_resolving"] / s, 100 * d["stall_lg"] / s,
          100 * d["stall_mio"] / s, 100 * d["stall_no_inst"] / s))
if len(sys.argv) > 5:
    # coarse view: instruction and sample shares per (file, line range) given as file:lo-hi:label ...
    print()
    for spec in sys.argv[5:]:
        f, rng, label = spec.split(":")
        lo, hi = [int(x) for x in rng.split("-")]
        ie = sum(d["Instructions Executed"] for (ff, l), d in agg.items() if ff == f and lo <= l <= hi)
        te = sum(d["Thread Instructions Executed"] for (ff, l), d in agg.items() if ff == f and lo <= l <= hi)
        sm = sum(d["# Samples"] for (ff, l), d in agg.items() if ff == f and lo <= l <= hi)
        print("%-28s inst %5.1f%%  lanes %5.1f  samples %5.1f%%" % (label, 100 * ie / tot["Instructions Executed"], te / max(ie, 1), 100 * sm / tot["# Samples"]))
