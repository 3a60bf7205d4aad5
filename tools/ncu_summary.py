#!/usr/bin/env python3
"""Summarise an ncu report (ncu -i X.ncu-rep --page raw --csv) into a small JSON: one object per profiled launch with the
metrics the roofline discussion uses.  usage: tools/ncu_summary.py <report.ncu-rep> [--out summary.json] [--extra k=v ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_requests",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_lsu_wavefront_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def main():
    rep = sys.argv[1]
    out = None
    extra = {}
    a = sys.argv[2:]
    while a:
        if a[0] == "--out":
            out = a[1]; a = a[2:]
        elif a[0] == "--extra":
            k, v = a[1].split("=", 1)
            try:
                v = json.loads(v)
            except Exception:
                pass
            extra[k] = v; a = a[2:]
        else:
            a = a[1:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS and v != "":
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                name = KEYS[h]
                if u in UNIT and name in ("duration", "dram_read", "dram_write", "l2_bytes"):
                    x *= UNIT[u]
                    name += "_s" if name == "duration" else "_bytes" if not name.endswith("bytes") else ""
                d[name] = x
        for h, u, v in zip(hdr, units, r):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v not in ("", "0"):
                d.setdefault("stall_per_issue", {})[h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = float(v)
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        launches.append(d)
    res = {"report": rep.split("/")[-1], "launches": launches, **extra}
    s = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(s + "\n")
    print(s[:3000])


if __name__ == "__main__":
    main()
