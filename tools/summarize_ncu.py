#!/usr/bin/env python
"""Summarise ncu CSV output (launch list from `--metrics gpu__time_duration.sum`, or `--page raw --csv` of a
--set full report) into a small markdown table for profiles/."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:90]


def launch_list(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    k, m, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    u = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= v or r[m] != "gpu__time_duration.sum":
            continue
        t = float(r[v].replace(",", ""))
        unit = r[u]
        t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
        a = agg.setdefault(short(r[k]), [0, 0.0])
        a[0] += 1
        a[1] += t_us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches (ncu per-launch times: cold cache, serialised)")


def raw_page(path, metrics):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = [(m, hdr.index(m)) for m in metrics if m in hdr]
    print("| kernel | " + " | ".join(f"{m} [{units[i]}]" for m, i in idx) + " |")
    print("|---|" + "---:|" * len(idx))
    for r in rows[2:]:
        print(f"| `{short(r[hdr.index('Kernel Name')])}` | " + " | ".join(r[i] for _, i in idx) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_list(sys.argv[2])
    else:
        raw_page(sys.argv[2], ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                               "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                               "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
                               "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                               "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"])
