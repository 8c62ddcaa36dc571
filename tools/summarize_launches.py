#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel averages and shares.
usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1000 if d["Metric Unit"] == "ns" else v
            name = d["Kernel Name"].replace("hv::<unnamed>::", "").split("(")[0]
            agg[(name, d["Grid Size"], d["Block Size"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | grid | block | launches | avg us | share |", "|---|---|---|---:|---:|---:|"]
    for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {g} | {b} | {len(v)} | {sum(v) / len(v):.2f} | {sum(v) / tot * 100:.1f}% |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
