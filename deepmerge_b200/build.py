"""Build libdeepmerge_b200.so in-tree with nvcc for sm_100a.

    python -m deepmerge_b200.build [--force]

Sources: deepmerge_b200/csrc/*.cu -> deepmerge_b200/libdeepmerge_b200.so (git-ignored,
travels to the GPU box with the snapshot).  No torch involvement: the library's ABI is
plain C (include/deepmerge_b200.h).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdeepmerge_b200.so")
SYNTH_LIB = os.path.join(HERE, "libdeepmerge_b200_synth.so")      # scene generator: bench / test utility, its own library
SYNTH_SRC = "synth.cu"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=default"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stamp():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
            os.path.join(os.path.dirname(HERE), "include", "deepmerge_b200.h"),
            os.path.join(os.path.dirname(HERE), "include", "deepmerge_b200_synth.h"), os.path.abspath(__file__)]:
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def needs_build():
    try:
        with open(os.path.join(OBJ, "stamp")) as f:
            return f.read().strip() != _stamp() or not os.path.exists(LIB) or not os.path.exists(SYNTH_LIB)
    except OSError:
        return True


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(_compile, _sources()))
    synth_obj = os.path.join(OBJ, SYNTH_SRC[:-3] + ".o")
    for lib, members in ((LIB, [o for o in objs if o != synth_obj]), (SYNTH_LIB, [synth_obj])):
        r = subprocess.run([NVCC, "-shared", "-o", lib, *members, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, "stamp"), "w") as f:
        f.write(_stamp())
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
