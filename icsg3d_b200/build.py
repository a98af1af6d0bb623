"""In-tree build of libicsg3d.so (sm_100a) with nvcc — no torch, no JIT cache.

`python -m icsg3d_b200.build` (or `__graft_entry__.build()`) compiles every csrc/*.cu to an object
file (in parallel, only when stale) and links `icsg3d_b200/libicsg3d.so`.  nvcc cross-compiles for
sm_100a without a GPU, so this runs on the CPU dev box; the .so then travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_build"
LIB = PKG / "libicsg3d.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--use_fast_math",
    "-Xptxas", "-v",
]
# The voxeliser's species predicate must be evaluated exactly like numpy/scipy fp64 (no FMA contraction,
# no fast-math): it gets its own flag set.
STRICT_FP = {"voxelize.cu"}
NVCC_FLAGS_STRICT = [f for f in NVCC_FLAGS if f != "--use_fast_math"] + ["--fmad=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libicsg3d cannot be built (there is no CPU fallback)")


def _stale(src: Path, obj: Path, deps: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in [src, *deps])


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h")) + [Path(__file__)]
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(src, obj, headers):
            flags = NVCC_FLAGS_STRICT if src.name in STRICT_FP else NVCC_FLAGS
            jobs.append((src, obj, [nvcc, *flags, "-c", str(src), "-o", str(obj)]))

    def run(job):
        src, obj, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        (OBJ / (src.stem + ".ptxas.txt")).write_text(r.stderr)
        if verbose:
            print(f"[build] {src.name}\n{r.stderr}", file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if jobs or not LIB.exists() or force:
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
