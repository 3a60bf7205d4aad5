#!/usr/bin/env python3
"""gpurun_out/<tag>_traffic.csv (tools/profile.sh traffic) -> profiles/march_ncu_summary.json: DRAM and L2 bytes of ONE march
launch (every k_gather_round / k_mlp_round / k_march_ws of the last 512-candidate chunk), stamped with the hash of the kernel
sources so that bench.py refuses to quote a capture of other code.  usage: march_traffic.py <traffic.csv> [candidates] [res] [scene]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path = sys.argv[1]
cands = int(sys.argv[2]) if len(sys.argv) > 2 else 512
res = int(sys.argv[3]) if len(sys.argv) > 3 else 800
scene = sys.argv[4] if len(sys.argv) > 4 else "shopping"
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {h: i for i, h in enumerate(rows[hi])}
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= col["Metric Value"]:
        continue
    d = launches.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()})
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name == "gpu__time_duration.sum":
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v      # -> ms
    elif unit in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    d[name] = v
ids = sorted(launches)
# the last chunk = everything after the last k_classify launch (+ that classify)
last_classify = max(i for i in ids if launches[i]["kernel"] == "k_classify")
chunk = [launches[i] for i in ids if i >= last_classify]
agg = {}
for l in chunk:
    a = agg.setdefault(l["kernel"], {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "lts_bytes": 0.0})
    a["launches"] += 1
    a["ms"] += l.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    a["lts_bytes"] += l.get("lts__t_bytes.sum", 0.0)
march = [k for k in agg if k in ("k_gather_round", "k_mlp_round", "k_march_ws")]
out = {
    "what": "ncu (dram__bytes_read/write.sum, lts__t_bytes.sum, gpu__time_duration.sum) over every march kernel of ONE launch: "
            f"{cands} candidates of the {scene} stand-in at {res}x{res}, 2^19-entry tables (tools/profile.sh traffic); cold-cache serialised times",
    "scene": scene, "candidates_per_launch": cands, "resolution": res, "source_sha1": bench.march_sources_sha1(),
    "kernel_launches": sum(agg[k]["launches"] for k in march),
    "dram_bytes_per_launch": sum(agg[k]["dram_bytes"] for k in march),
    "lts_bytes_per_launch": sum(agg[k]["lts_bytes"] for k in march),
    "march_ms": sum(agg[k]["ms"] for k in march),
    "per_kernel": agg,
}
json.dump(out, open(os.path.join(ROOT, "profiles", "march_ncu_summary.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
