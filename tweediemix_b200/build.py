"""Build libtmx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tweediemix_b200.build [--force] [--verbose]

The .so lands in tweediemix_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libtmx.so")
STAMP = os.path.join(LIBDIR, "libtmx.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _flags():
    """NVCC_FLAGS plus optional tuning defines from $TMX_NVCC_EXTRA (e.g. "-DTMX_ATTN_POLY_EVERY=4")."""
    return NVCC_FLAGS + os.environ.get("TMX_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    deps.append(os.path.join(INCLUDE, "tmx.h"))
    for p in deps:
        h.update(p.encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """Compile csrc/*.cu and link libtmx.so (or `out`, a variant library built with $TMX_NVCC_EXTRA defines).  With
    $TMX_LIB_PATH set the caller has chosen a prebuilt variant: nothing is rebuilt."""
    if os.environ.get("TMX_LIB_PATH") and out is None:
        path = os.environ["TMX_LIB_PATH"]
        if not os.path.exists(path):
            raise RuntimeError(f"TMX_LIB_PATH={path} does not exist")
        return path
    if out is not None:
        return _build_variant(out, verbose)
    os.makedirs(LIBDIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), *_flags(), "-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {os.path.basename(src)}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log), file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed; see tweediemix_b200/lib/build.log")
    link = [_nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout, file=sys.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


def _build_variant(out: str, verbose: bool) -> str:
    """A side build (objects under lib/<name>.objs/) that leaves libtmx.so and its stamp alone."""
    objdir = out + ".objs"
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        procs.append((src, subprocess.Popen([_nvcc(), *_flags(), "-I", INCLUDE, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(o, file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    r = subprocess.run([_nvcc(), "-shared", "-o", out, *objs, "-cudart", "static", "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout, file=sys.stderr)
        raise RuntimeError("link failed")
    shutil.rmtree(objdir, ignore_errors=True)
    return out


if __name__ == "__main__":
    if "--out" in sys.argv:
        print(build(out=os.path.abspath(sys.argv[sys.argv.index("--out") + 1]), verbose="--verbose" in sys.argv))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
