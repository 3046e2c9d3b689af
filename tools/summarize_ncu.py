#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches.csv into the small text summaries committed under profiles/.
Usage: python tools/summarize_ncu.py <tag>   (reads gpurun_out/, writes profiles/<tag>_*.md)"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        res.append((d, dict(zip(hdr, units))))
    return res


def launches(tag):
    p = os.path.join(SRC, "launches.csv")
    if not os.path.isfile(p):
        return
    rows = list(csv.reader(open(p)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d.get("Metric Unit", "ns")
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            agg[d["Kernel Name"][:90]][0] += 1
            agg[d["Kernel Name"][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised -> compare SHARES)\n\n")
        f.write(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write(f"| {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k}` |\n")


def kernel_report(tag, name):
    rep = os.path.join(SRC, f"{name}.ncu-rep")
    if not os.path.isfile(rep):
        return
    with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none ({name}.ncu-rep)\n\n")
        for d, u in raw_page(rep):
            f.write(f"## {d.get('Kernel Name', '?')[:120]}\n\n| metric | unit | value |\n|---|---|---:|\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"| {k} | {u.get(k, '')} | {d[k]} |\n")
            f.write("\n")
        # warp-stall sampling totals (SASS level)
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) > 2:
            # first kernel only
            hdr = rows[1]
            idx = {h: i for i, h in enumerate(hdr)}
            cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot = collections.Counter()
            top = []
            for r in rows[2:]:
                if len(r) < len(hdr) or r[0] == "Kernel Name" or not r[idx["# Samples"]].isdigit():
                    if r and r[0] == "Kernel Name":
                        break
                    continue
                for h in cols:
                    tot[h] += int(r[idx[h]] or 0)
                top.append((int(r[idx["# Samples"]]), r[idx["Source"]].strip()))
            T = sum(tot.values()) or 1
            f.write("## warp stall sampling (first kernel), share of samples\n\n")
            for h, v in tot.most_common(8):
                f.write(f"- {h}: {100 * v / T:.1f}%\n")
            f.write("\nhottest SASS instructions:\n\n")
            for n, s in sorted(top, reverse=True)[:12]:
                f.write(f"- {100 * n / T:.1f}%  `{s[:90]}`\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    for n in ("prof_decode", "prof_gemm", "prof_attn", "prof_attn_t5"):
        kernel_report(tag, n)
    print(os.listdir(OUT))
