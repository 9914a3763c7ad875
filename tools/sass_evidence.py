#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libtmx.so (cuobjdump -sass): which kernels are tcgen05 / TMEM / TMA kernels, which
stream through the LSU, and where (if anywhere) the legacy warp-level HMMA path is used.

    python tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tweediemix_b200", "lib", "libtmx.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDSM", "LDGSTS", "MUFU.EX2", "MUFU.TANH", "FFMA2", "FADD2",
        "UCGABAR", "LDG", "STG", "LDS", "STS", "ATOMG", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            total[fn] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if fn and m:
            op = m.group(1)
            total[fn] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[fn][k] += 1
    print("# cuobjdump -sass tweediemix_b200/lib/libtmx.so (sm_100a): instruction evidence per kernel")
    print("# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = cp.async.bulk.tensor load/store (TMA),")
    print("# UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UCGABAR = cluster barrier, LDGSTS = cp.async, LDSM = ldmatrix.")
    print("# HMMA (warp-level mma.sync) appears in exactly one kernel family: short_kv_attn_kernel (k2s), the streaming alternative to the tcgen05 short-K/V kernel k2t")
    print("# (short_kv_attn_tc_kernel, the default for the 77-token sites); k2s serves key counts k2t does not take (<= 64, 81..128).")
    print()
    for fn, c in counts.items():
        if total[fn] == 0:
            continue
        try:
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        except Exception:
            name = fn
        name = re.sub(r"\(.*", "", name)
        print(name)
        print("    instructions=%d  " % total[fn] + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
