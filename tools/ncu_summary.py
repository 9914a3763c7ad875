#!/usr/bin/env python
"""Extract the judged metrics from an .ncu-rep (raw page) into a short text summary.

    python tools/ncu_summary.py gpurun_out/r01a_ncu/attn_n4096.ncu-rep [more.ncu-rep ...] > profiles/xxx.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr = i
            break
    if hdr is None:
        return []
    names, units = rows[hdr], rows[hdr + 1]
    res = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        res.append({n: (v, u) for n, u, v in zip(names, units, r)})
    return res


def main(paths, extra):
    for p in paths:
        for k in raw(p):
            print(f"== {p} :: {k['Kernel Name'][0][:110]}  grid {k.get('Grid Size', ('?',))[0]} block {k.get('Block Size', ('?',))[0]}")
            for name in sorted(k):
                if any(name == key or (key in name and key in extra) for key in KEYS + extra) or any(s in name for s in ("stall", "warp_issue_stalled")) and "pct" in name:
                    v, u = k[name]
                    if v not in ("", "0", "n/a"):
                        print(f"   {name:95s} {v:>18s} {u}")
            try:
                rd = float(k["dram__bytes_read.sum"][0].replace(",", "")); wr = float(k["dram__bytes_write.sum"][0].replace(",", ""))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = rd * scale.get(k["dram__bytes_read.sum"][1], 1) + wr * scale.get(k["dram__bytes_write.sum"][1], 1)
                print(f"   traffic (dram read+write) = {tr / 1e6:.2f} MB")
            except Exception:
                pass


if __name__ == "__main__":
    args = sys.argv[1:]
    extra = [a[2:] for a in args if a.startswith("--")]
    main([a for a in args if not a.startswith("--")], extra)
