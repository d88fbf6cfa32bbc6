"""Build libsfd2_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m sfd2_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/sfd2_b200.h); nvcc
cross-compiles it on a box without a GPU.  Objects are cached by source mtime.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsfd2_b200.so")
OBJ = os.path.join(HERE, "_obj")
SOURCES = ["api.cu", "simt_conv.cu", "tc_conv.cu", "tc_conv1a.cu", "tc_desc_sparse.cu", "tc_match.cu", "post.cu", "match.cu", "preprocess.cu"]
# hardware probes (tools/umma_probe.py, tools/mma_rate_probe.py): measurement scaffolding, only on request
if os.environ.get("SFD2_WITH_PROBES") == "1":
    SOURCES.append("umma_probe.cu")
# every header under csrc/ (a stale api.o once ran against a changed TcMatchArgs layout: tc_match.cuh was missing here)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "sfd2_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _mtime(obj) < max(_mtime(src), hdr_t):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- nvcc {s}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
