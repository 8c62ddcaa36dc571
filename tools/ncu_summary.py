#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (`ncu --set full`) into markdown + a small JSON with the DRAM traffic.
usage: ncu_summary.py report.ncu-rep out.md [traffic.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration (under ncu: cold caches, serialised)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes.sum.per_second", "DRAM bytes/s"),
    ("lts__t_sectors_op_write.sum", "L2 write sectors"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "SM -> L2 write bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes (L2 -> SM)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.sum", "warp instructions executed"),
    ("smsp__inst_executed.sum", "warp instructions executed (smsp)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
]


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"### `{d['Kernel Name'][:110]}`\n")
        lines.append("| metric | value | unit |\n|---|---:|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                lines.append(f"| {label} (`{k}`) | {d[k]} | {u[k]} |")

        def to_bytes(k):
            v = float(d[k].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]

        if "dram__bytes_read.sum" in d:
            traffic = {"kernel": d["Kernel Name"][:80], "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
                       "dram_bytes_write": to_bytes("dram__bytes_write.sum")}
            traffic["dram_bytes_per_launch"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
        lines.append("")
    open(out_md, "w").write(f"ncu `--set full --clock-control none` summary of `{rep.split('/')[-1]}`\n\n" + "\n".join(lines))
    if len(sys.argv) > 3 and traffic:
        traffic["source"] = rep.split("/")[-1]
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
