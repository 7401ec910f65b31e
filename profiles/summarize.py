#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, *_full.ncu-rep, bench.json} into a committed summary.

    python profiles/summarize.py <tag>      -> profiles/<tag>/{summary.md, launches_top.csv, bench.json, *.metrics.csv}
"""
import csv
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def launches_summary(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")[:90]
        tot[name] += v / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    out = [("kernel", "launches", "total_us", "avg_us", "share")]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append((k, cnt[k], f"{v:.1f}", f"{v / cnt[k]:.1f}", f"{v / total:.3f}"))
    return out, total, sum(cnt.values())


def rep_metrics(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:80]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        out.append(d)
    return out


def main():
    tag = sys.argv[1]
    dst = os.path.join(ROOT, "profiles", tag)
    os.makedirs(dst, exist_ok=True)
    md = [f"# {tag}\n"]
    bj = os.path.join(SRC, "bench.json")
    if os.path.exists(bj):
        shutil.copy(bj, os.path.join(dst, "bench.json"))
        md.append("## bench.py line\n\n```\n" + open(bj).read().strip()[-4000:] + "\n```\n")
    lc = os.path.join(SRC, "launches.csv")
    if os.path.exists(lc):
        table, total, n = launches_summary(lc)
        with open(os.path.join(dst, "launches_top.csv"), "w", newline="") as f:
            csv.writer(f).writerows(table)
        md.append(f"## ncu launch list (gpu__time_duration.sum, --clock-control none): {n} launches, "
                  f"{total / 1e3:.2f} ms total -- cold-cache, serialised: compare SHARES\n")
        md.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
        for row in table[1:16]:
            md.append("| " + " | ".join(str(c) for c in row) + " |")
        md.append("")
    for f in sorted(os.listdir(SRC)):
        if f.endswith(".ncu-rep"):
            mets = rep_metrics(os.path.join(SRC, f))
            with open(os.path.join(dst, f.replace(".ncu-rep", ".metrics.csv")), "w", newline="") as out:
                w = csv.writer(out)
                w.writerow(["kernel"] + KEYS)
                for d in mets:
                    w.writerow([d.get("kernel")] + [d.get(k, "") for k in KEYS])
            md.append(f"## {f} (ncu --set full)\n")
            for d in mets:
                md.append("* `" + d["kernel"] + "`")
                for k in KEYS:
                    if k in d:
                        md.append(f"  * {k}: {d[k]}")
            md.append("")
    # per-launch DRAM traffic of the dominant aggregation kernel for bench.py's roofline.traffic
    sp = os.path.join(SRC, "spmm_full.ncu-rep")
    if os.path.exists(sp):
        import json
        for d in rep_metrics(sp):
            if "k_spmm_rows<float, 4" in d["kernel"] or "k_spmm_fast<float, 4, 32" in d["kernel"]:
                def mb(v):
                    num, unit = v.split()[:2]
                    return float(num) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                tot = mb(d["dram__bytes_read.sum"]) + mb(d["dram__bytes_write.sum"])
                json.dump({"bytes_per_launch": tot, "source": f"profiles/{tag}/spmm_full.metrics.csv "
                           "(dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"},
                          open(os.path.join(ROOT, "profiles", "spmm_traffic.json"), "w"))
                break
    open(os.path.join(dst, "summary.md"), "w").write("\n".join(md) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
