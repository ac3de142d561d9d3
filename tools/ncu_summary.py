"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.json
      per-kernel count / total / mean duration of an `ncu --metrics gpu__time_duration.sum` launch list
  python tools/ncu_summary.py full gpurun_out/r1_decode_mega.ncu-rep decode_mega_kernel profiles/r1_ncu_full_decode_mega_kernel.json [ctx]
      key metrics of one `ncu --set full` capture (read with `ncu -i ... --page raw --csv`)
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    m = re.search(r"(\w+)\s*\(", name)
    return m.group(1) if m else name


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, order = OrderedDict(), []
    for r in rows[1:]:
        n = short(r[k])
        ns = float(r[v].replace(",", ""))
        a = agg.setdefault(n, {"count": 0, "total_us": 0.0})
        a["count"] += 1
        a["total_us"] += ns / 1e3
        order.append(n)
    total = sum(a["total_us"] for a in agg.values())
    for a in agg.values():
        a["mean_us"] = a["total_us"] / a["count"]
        a["share"] = a["total_us"] / total
    out = {"source": src, "launches": len(order), "total_us": total,
           "note": "ncu-serialised, cold-cache per-launch times: shares are meaningful, absolutes are not",
           "kernels": OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["total_us"]))}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k2: (v2["count"], round(v2["total_us"], 1), round(v2["share"], 3)) for k2, v2 in out["kernels"].items()}, indent=0))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def full(rep, kernel, dst, ctx=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if kernel in r[hdr.index("Kernel Name")])
    m = {}
    for h, u, v in zip(hdr, units, row):
        if h in KEYS:
            try:
                m[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                m[h] = {"value": v, "unit": u}

    def to(h, table):
        e = m[h]
        return e["value"] * table.get(e["unit"].split("/")[0], 1)
    out = {"source": rep, "kernel": kernel, "ctx": ctx,
           "duration_us_under_ncu": to("gpu__time_duration.sum", SCALE),
           "dram_bytes_read": to("dram__bytes_read.sum", SCALE), "dram_bytes_write": to("dram__bytes_write.sum", SCALE),
           "metrics": m}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("duration_us_under_ncu", "dram_bytes_read", "dram_bytes_write")}))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else None)
