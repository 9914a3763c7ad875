#!/usr/bin/env python
"""Aggregate the ncu source page (SASS view) of a report: stall-reason totals, the top stalled
instructions, and the instruction mix by opcode.   python tools/ncu_stalls.py rep.ncu-rep [topN]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    for r in body:
        for n in stall_cols:
            tot[n] += int(r[col[n]] or 0)
    samples = sum(int(r[col["# Samples"]] or 0) for r in body)
    print(f"# {path}\n# {len(body)} SASS instructions, {samples} samples")
    print("stall reasons (all samples):")
    for n, v in tot.most_common():
        if v:
            print(f"   {n:28s} {v:8d} {100 * v / max(samples, 1):5.1f}%")
    ops = collections.Counter()
    for r in body:
        op = r[col["Source"]].split()
        op = [t for t in op if not t.startswith("@")]
        ops[op[0].split(".")[0] if op else "?"] += int(r[col["Instructions Executed"]] or 0)
    n_inst = sum(ops.values())
    print(f"warp instructions executed: {n_inst}; mix:")
    for o, v in ops.most_common(25):
        print(f"   {o:12s} {v:12d} {100 * v / n_inst:5.1f}%")
    print(f"top {top} instructions by samples:")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][col['# Samples']] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        reasons = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:3]
        rs = " ".join(f"{n}:{v}" for v, n in reasons if v)
        print(f"   #{i:5d} {int(r[col['# Samples']]):6d}  x{int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']].strip()[:70]:70s} {rs}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
