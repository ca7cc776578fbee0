"""In-tree build of libtfin.so (nvcc, sm_100a only).  Used by ``__graft_entry__.build()`` and by hand:
``python -m bayesianinferencedl_b200._build [--force] [-v]``.

The heavily unrolled PCG kernel variants are instantiated in separate translation units (pcg_inst.cu compiled
once per (group, nodal) pair) so that they build in parallel; objects land in csrc/_obj (git-ignored)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libtfin.so")
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "tfin.h")]
PCG_GROUPS = 5
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libtfin cannot be built")
    return exe


def _units():
    """(source, object, extra flags) for every translation unit."""
    units = [("tfin_api.cu", "tfin_api.o", [])]
    for g in range(PCG_GROUPS):
        for nodal in (0, 1):
            units.append(("pcg_inst.cu", f"pcg_inst_g{g}_n{nodal}.o", [f"-DPCG_GROUP={g}", f"-DPCG_NODAL={nodal}"]))
    return units


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in ["tfin_api.cu", "pcg_inst.cu"] + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, jobs=None):
    """Compile csrc/*.cu into bayesianinferencedl_b200/libtfin.so; returns the library path."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    logs = []

    def compile_one(unit):
        src, obj, extra = unit
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + \
              ["-c", "-o", os.path.join(OBJ, obj), os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {obj}:\n{proc.stdout}{proc.stderr}")
        logs.append(f"== {obj}\n{proc.stderr}")

    units = _units()
    with ThreadPoolExecutor(max_workers=jobs or min(len(units), os.cpu_count() or 4)) as ex:
        list(ex.map(compile_one, units))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + \
           [os.path.join(OBJ, u[1]) for u in units]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
