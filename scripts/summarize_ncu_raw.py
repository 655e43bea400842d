#!/usr/bin/env python
"""One line per launch of an `ncu --page raw --csv` export (scripts/ncu_step_capture.sh): duration, DRAM bytes and
throughput, L2 and L1 throughput, tensor-pipe activity, occupancy limits and the largest warp-stall reason.
usage: summarize_ncu_raw.py raw.csv [--by-kernel]"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
by_kernel = "--by-kernel" in sys.argv
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def to_mb(r, name):
    v, u = f(r, name), units[col[name]] if name in col else ""
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)


def to_us(r, name):
    v, u = f(r, name), units[col[name]] if name in col else ""
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("dfb::", "").replace("(anonymous namespace)::", "")


out = []
for r in data:
    if len(r) < len(hdr) // 2:
        continue
    stalls = sorted(((f(r, c), c) for c in stall_cols), reverse=True)
    top = re.sub(r"^smsp__average_warps?(_latency)?_issue_stalled_|_per_issue_active\.ratio$|\.ratio$", "", stalls[0][1]) if stalls else "-"
    out.append(OrderedDict(
        kernel=short(r[col["Kernel Name"]]), grid=int(f(r, "launch__grid_size")), block=int(f(r, "launch__block_size")),
        us=to_us(r, "gpu__time_duration.sum"),
        dram_mb=to_mb(r, "dram__bytes_read.sum") + to_mb(r, "dram__bytes_write.sum"),
        dram_pct=0.0,
        l2_pct=f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        l1_pct=f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        tensor_pct=f(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
                     f(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")),
        regs=int(f(r, "launch__registers_per_thread")), smem_kb=f(r, "launch__shared_mem_per_block_dynamic"),
        warps_pct=f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), stall=top))
    o = out[-1]   # DRAM bytes over the duration against the measured copy bandwidth (MEASURED_PEAKS.json: 6454 GB/s)
    o["dram_pct"] = 100.0 * (o["dram_mb"] * 1e6 / max(o["us"] * 1e-6, 1e-12)) / 6454e9

if by_kernel:
    agg = OrderedDict()
    for o in out:
        a = agg.setdefault(o["kernel"], dict(n=0, us=0.0, dram_mb=0.0, dram_pct=0.0, l2_pct=0.0, tensor_pct=0.0, regs=o["regs"], smem_kb=o["smem_kb"], stall={}))
        a["n"] += 1
        for k in ("us", "dram_mb", "dram_pct", "l2_pct", "tensor_pct"):
            a[k] += o[k]
        a["stall"][o["stall"]] = a["stall"].get(o["stall"], 0) + 1
    print("%-78s %3s %8s %9s %6s %6s %6s %4s %6s  %s" % ("kernel", "n", "avg us", "DRAM MB", "DRAM%", "L2%", "TC%", "regs", "smemKB", "top stall"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        print("%-78s %3d %8.2f %9.3f %6.1f %6.1f %6.1f %4d %6.1f  %s" % (k[:78], n, a["us"] / n, a["dram_mb"] / n, a["dram_pct"] / n, a["l2_pct"] / n,
                                                                  a["tensor_pct"] / n, a["regs"], a["smem_kb"], max(a["stall"], key=a["stall"].get)))
else:
    print("%3s %-70s %6s %8s %9s %6s %6s %6s %6s  %s" % ("#", "kernel", "grid", "us", "DRAM MB", "DRAM%", "L2%", "L1%", "TC%", "top stall"))
    for i, o in enumerate(out):
        print("%3d %-70s %6d %8.2f %9.3f %6.1f %6.1f %6.1f %6.1f  %s" % (i, o["kernel"][:70], o["grid"], o["us"], o["dram_mb"], o["dram_pct"], o["l2_pct"],
                                                                    o["l1_pct"], o["tensor_pct"], o["stall"]))
