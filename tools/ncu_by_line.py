#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` (SASS view) export per CUDA source line, using the line table of the
cubin that ran (nvdisasm -g).  usage: ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <mangled kernel> [top]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line) from the disassembly of this kernel
addr2line = {}
cur = None
inside = False
for ln in open(dis):
    if ln.startswith(".text." + kern + ":"):
        inside = True
        continue
    if inside and ln.startswith("//-----"):
        break
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
fields = ["# Samples", "Instructions Executed", "Thread Instructions Executed", "L1 Tag Requests Global", "L2 Theoretical Sectors Global",
          "stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_no_inst", "stall_math", "stall_branch_resolving",
          "stall_selected", "stall_not_selected", "L1 Wavefronts Shared"]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") or re.fullmatch(r"[0-9a-f]+", r[col["Address"]]) else None
    if a is None:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    for f in fields:
        try:
            v = float(r[col[f]] or 0)
        except ValueError:
            v = 0
        agg[key][f] += v
        tot[f] += v
print("totals:", {f: int(tot[f]) for f in fields})
print(f"{'line':34s} {'samp%':>6s} {'inst%':>6s} {'thr/inst':>8s} {'L1req%':>6s} {'long':>6s} {'bar':>6s} {'wait':>6s} {'short':>6s} {'noinst':>6s} {'sel':>6s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    ie = v["Instructions Executed"]
    print(f"{key[0][:26]:26s}:{key[1]:<6d} {100*v['# Samples']/tot['# Samples']:6.2f} {100*ie/tot['Instructions Executed']:6.2f} "
          f"{(v['Thread Instructions Executed']/ie if ie else 0):8.1f} {100*v['L1 Tag Requests Global']/max(1,tot['L1 Tag Requests Global']):6.2f} "
          f"{100*v['stall_long_sb']/tot['# Samples']:6.2f} {100*v['stall_barrier']/tot['# Samples']:6.2f} {100*v['stall_wait']/tot['# Samples']:6.2f} "
          f"{100*v['stall_short_sb']/tot['# Samples']:6.2f} {100*v['stall_no_inst']/tot['# Samples']:6.2f} {100*v['stall_selected']/tot['# Samples']:6.2f}")
