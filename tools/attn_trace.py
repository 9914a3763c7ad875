#!/usr/bin/env python
"""Print the per-event timeline of CTA 0 of the attention kernel (needs a -DTMX_ATTN_TRACE build).
    TMX_NVCC_EXTRA=-DTMX_ATTN_TRACE python tools/attn_trace.py [N] [first_event] [count]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tweediemix_b200 import build, ops, _lib  # noqa: E402

NAMES = {0: "tma  k_empty ok", 1: "tma  v_empty ok", 10: "mma  k_full(next) ok", 11: "mma  s_free ok", 12: "mma  QK issued",
         13: "pv   v_full ok", 16: "mma    4 QK MMAs issued", 17: "mma    commit 1 issued", 18: "mma    elected", 19: "mma    8 PV MMAs issued", 14: "pv   p_full ok", 15: "pv   PV issued", 20: "smx  s_full ok", 21: "smx  S loaded, s_free sent",
         22: "smx  exp done (full tile)", 23: "smx  pv_done ok", 24: "smx  exp done", 25: "smx  P stored, p_full sent",
         30: "smx  EPILOGUE pv_done ok", 31: "smx  EPILOGUE barrier A passed", 32: "smx  EPILOGUE O staged, barrier B passed", 33: "smx  EPILOGUE store issued + read"}


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    count = int(sys.argv[3]) if len(sys.argv) > 3 else 90
    build.build()
    lib = _lib.load()
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    B = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    q = torch.randn(B, N, H * 64, device="cuda", dtype=torch.bfloat16)
    k = torch.randn_like(q); v = torch.randn_like(q)
    for _ in range(2):
        ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    ev = []
    for role in range(4):
        buf = (ctypes.c_longlong * 4096)()
        lib.tmx_attn_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
        assert lib.tmx_attn_debug_trace(buf, role) == 0
        a = list(buf)
        for i in range(0, 4096, 2):
            if a[i + 1] == 0:
                break
            ev.append((a[i + 1], a[i]))
    ev.sort()
    t0 = ev[0][0]
    prev = {}
    for t, tag in ev[first:first + count]:
        role = NAMES[tag][:3]
        d = t - prev.get(role, t)
        prev[role] = t
        print(f"{t - t0:9d}  (+{d:5d})  {NAMES[tag]}")


if __name__ == "__main__":
    main()
