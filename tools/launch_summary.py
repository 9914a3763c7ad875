#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: total us, share, launches.

    python tools/launch_summary.py profiles/r01a_launches_fused_step.csv > profiles/r01a_launches_fused_step.summary.txt
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v / 1e3)
        name = r[ki]
        name = name.split("(")[0][:90] if name.startswith("void ") else name[:90]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for n, v in tot.items() if any(ns in n for ns in ("tmx::", "k10::", "k3::", "xattn")))
    print(f"# {path}: {sum(cnt.values())} launches, {total:.1f} us serialised (cold-cache; compare shares, not absolutes)")
    print(f"# hand-written kernels (tmx:: / k10:: / k3:: / xattn*::): {ours:.1f} us = {100 * ours / total:.1f} % ; library (cuBLAS nvjet / cuDNN cutlass / ATen): {100 - 100 * ours / total:.1f} %")
    for n, v in tot.most_common():
        print(f"{v:10.1f} us {100 * v / total:5.1f}%  x{cnt[n]:4d}  {n}")


if __name__ == "__main__":
    main(sys.argv[1])
