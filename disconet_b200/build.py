"""Build libdisco_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python -m disconet_b200.build [--force]

The shared library sits next to this file so that it travels with a repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdisco_b200.so")
STAMP = os.path.join(HERE, ".libdisco_b200.stamp")
SOURCES = ["conv_tc.cu", "conv_ref.cu", "misc.cu", "fusion.cu", "train.cu", "wgrad.cu", "fusion_train.cu", "seg.cu", "post.cu", "capi.cu"]
HEADERS = ["common.cuh", "conv.h", "ops.h", "train.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(HERE, "build")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_digest()


def _file_digest(name: str) -> str:
    h = hashlib.sha256()
    for n in [name] + HEADERS:
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(n.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every .cu to an object (in parallel, skipped when source + headers are unchanged), then link."""
    if not force and is_fresh():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        stamp = obj + ".stamp"
        objs.append(obj)
        dig = _file_digest(src)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print("[disconet_b200.build]", " ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd, cwd=CSRC), cmd, stamp, dig))
    for p, cmd, stamp, dig in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
        with open(stamp, "w") as f:
            f.write(dig)
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", LIB]
    if verbose:
        print("[disconet_b200.build]", " ".join(link), flush=True)
    subprocess.run(link, check=True, cwd=CSRC)
    with open(STAMP, "w") as f:
        f.write(source_digest())
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
