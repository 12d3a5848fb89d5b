#!/usr/bin/env python
"""Builds profiles/rNN_summary.md (+ the copied launch lists / raw metric CSVs) from gpurun_out/.
   python tools/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("spp::", ""),
                                    "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(by.values())


def share_table(rows, skip_frac=0.45):
    rows = rows[int(len(rows) * skip_frac):]  # steady-state part (after warm-up batches)
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["gpu__time_duration.sum"]
        a[2] += r.get("dram__bytes_read.sum", 0.0)
        a[3] += r.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total us | mean us | share | DRAM rd MB/launch | DRAM wr MB/launch |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {100 * a[1] / tot:.1f}% | "
                   f"{a[2] / a[0] / 1e6:.2f} | {a[3] / a[0] / 1e6:.2f} |")
    return "\n".join(out), agg


md = [f"# Profile summary {tag} (B200, ogbn-products-shaped workload, fanout (15,10,5), batch 1024)", "",
      "Command profiled: `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (the plain run exited 0 first).",
      "ncu launch lists are serialised; the cold list flushes caches between kernels, the warm one "
      "(`--cache-control none`) keeps L2 contents like the real pipeline. Compare SHARES, not absolutes.", ""]
for kind in ("cold", "warm"):
    src = os.path.join(G, f"{tag}_launches_{kind}.csv")
    if not os.path.exists(src):
        continue
    shutil.copy(src, os.path.join(P, f"{tag}_launches_{kind}.csv"))
    rows = launches(src)
    table, agg = share_table(rows)
    md += [f"## Launch list ({kind} cache): `profiles/{tag}_launches_{kind}.csv`", "", table, ""]
    if kind == "warm":
        per_batch = [r for r in rows[int(len(rows) * 0.45):]]
        # one mini-batch = the kernels between two k_seeds_init
        idx = [i for i, r in enumerate(per_batch) if r["name"].startswith("k_seeds_init")]
        if len(idx) >= 2:
            one = per_batch[idx[0]:idx[1]]
            md += ["One mini-batch, in launch order (warm):", "", "| kernel | grid | us | DRAM rd MB | DRAM wr MB |", "|---|---|---|---|---|"]
            for r in one:
                md.append(f"| `{r['name']}` | {r['grid']} | {r['gpu__time_duration.sum'] / 1e3:.2f} | "
                          f"{r.get('dram__bytes_read.sum', 0) / 1e6:.2f} | {r.get('dram__bytes_write.sum', 0) / 1e6:.2f} |")
            md += ["", f"Sum of the serialised kernel times of one mini-batch: "
                       f"{sum(r['gpu__time_duration.sum'] for r in one) / 1e3:.1f} us.", ""]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed.sum"]
traffic = {}
for rep in sorted(f for f in os.listdir(G) if f.startswith(tag) and f.endswith(".ncu-rep")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, rep.replace(".ncu-rep", "_raw.csv")), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    md += [f"## `{rep}` (ncu --set full; raw metrics in `profiles/{rep.replace('.ncu-rep', '_raw.csv')}`)", ""]
    for r in rows[2:]:
        md.append(f"Kernel `{r[ix['Kernel Name']]}` grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}:")
        md.append("")
        for k in KEYS:
            if k in ix:
                md.append(f"* `{k}` = {r[ix[k]]} {units[ix[k]]}")
        md.append("")
        if "gather" in rep:
            def num(k):
                v, u = float(r[ix[k]].replace(",", "")), units[ix[k]].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
            traffic = {"kernel": r[ix["Kernel Name"]], "dram_bytes_read": num("dram__bytes_read.sum"),
                       "dram_bytes_write": num("dram__bytes_write.sum"),
                       "duration_us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")),
                       "source": f"profiles/{rep.replace('.ncu-rep', '_raw.csv')}"}
    top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top_lines.py"), os.path.join(G, rep), "12"],
                         capture_output=True, text=True).stdout
    md += ["Top stall source lines (warp-state samples):", "", "```", top.rstrip(), "```", ""]
if traffic:
    traffic["traffic_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
    json.dump(traffic, open(os.path.join(P, f"{tag}_gather_traffic.json"), "w"), indent=1)
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md)[:6000])
