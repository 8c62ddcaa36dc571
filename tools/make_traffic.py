#!/usr/bin/env python3
"""profiles/traffic.json from `ncu --set full` captures of the headline step (K1 + per-frame CCL kernel), stamped with the
fingerprint of the kernel sources they were taken from: bench.py quotes the DRAM traffic only when the fingerprint matches
the code it runs.  Also writes the markdown summaries.

usage: make_traffic.py TAG plain.ncu-rep [morph.ncu-rep]        (run on the CPU box; reads the reports with `ncu -i`)

capture (one gpurun call):
  python tools/sweep_k1.py 40 && ncu --set full --clock-control none --import-source on \
      -k regex:"k_preprocess_tma|k_ccl_frame" -s 60 -c 4 -o gpurun_out/TAG_k1_ccl python tools/sweep_k1.py 40
  (SWEEP_MORPH=3 ... -o gpurun_out/TAG_k1m_ccl for the pipeline with morphology)
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha16():
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "heimdall-vision_b200", "csrc")
    for f in ("expand_tile.cuh", "hv_common.cuh", "k_ccl_frame.cu", "k_preprocess.cu", "score_device.cuh"):  # = bench.py
        hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


def dram(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    acc = {}
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))

        def tob(k):
            return float(d[k].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
        kind = "k1" if "k_preprocess_tma" in d["Kernel Name"] else "ccl"
        acc.setdefault(kind, []).append(tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum"))
    return {k: int(sum(v) / len(v)) for k, v in acc.items()}


def main():
    tag, plain = sys.argv[1], sys.argv[2]
    morph = sys.argv[3] if len(sys.argv) > 3 else None
    p = dram(plain)
    t = {"csrc_sha16": csrc_sha16(),
         "source": f"profiles/{tag}_k1_ccl_ncu.md (ncu --set full of tools/sweep_k1.py: headline batch, shipped configuration)",
         "k1_dram_bytes_per_launch": p["k1"], "ccl_dram_bytes_per_launch": p["ccl"], "algorithmic_bytes_per_launch": 196608000}
    summ = os.path.join(ROOT, "tools", "ncu_summary.py")
    subprocess.run(["python", summ, plain, os.path.join(ROOT, "profiles", f"{tag}_k1_ccl_ncu.md")], capture_output=True)
    if morph:
        m = dram(morph)
        t["morph_k1_dram_bytes_per_launch"], t["morph_ccl_dram_bytes_per_launch"] = m["k1"], m["ccl"]
        subprocess.run(["python", summ, morph, os.path.join(ROOT, "profiles", f"{tag}_k1m_ccl_ncu.md")], capture_output=True)
    json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(t)


if __name__ == "__main__":
    main()
