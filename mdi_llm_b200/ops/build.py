"""Build the sm_100a kernel library in-tree.

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` over ``csrc/*.cu`` into one shared
object ``mdi_llm_b200/ops/_mdi_ops.so`` (plain C ABI, loaded with ctypes — no torch headers, so a
full rebuild takes seconds and the binary does not depend on the torch ABI).  The ``.so`` is
git-ignored but travels with the ``gpurun`` snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path
from typing import List, Optional

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "_mdi_ops.so"
STAMP = HERE / "_mdi_ops.stamp"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> Optional[str]:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def sources() -> List[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


_MARK = b"MDI_BUILD_DIGEST:"


def embedded_digest(path: Path = LIB) -> Optional[str]:
    """Digest of the sources a built library was compiled from, read out of the binary itself (the string is
    compiled into ``runtime.cu``), so a stale ``.so`` can never pass for a fresh one whatever the state of the
    work tree (a separate stamp file could — and did — get out of sync with a git-ignored binary)."""
    try:
        blob = path.read_bytes()
    except OSError:
        return None
    i = blob.find(_MARK)
    if i < 0:
        return None
    return blob[i + len(_MARK): i + len(_MARK) + 64].decode("ascii", "replace")


def is_fresh() -> bool:
    return LIB.exists() and embedded_digest() == _digest()


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    """Compile every ``.cu`` (parallel object builds) and link ``_mdi_ops.so``."""
    if not force and is_fresh():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    digest = _digest()
    common = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH_FLAGS, "-Xcompiler", "-fPIC",
              "--use_fast_math", "-I", str(CSRC), f'-DMDI_BUILD_DIGEST_STR="{digest}"']
    if ptxas_info:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = common + ["-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] {src.name} FAILED\n{out}\n")
        elif (verbose or ptxas_info) and out.strip():
            print(f"[build] {src.name}\n{out}")
    if failed:
        raise RuntimeError("nvcc failed (see stderr)")
    link = [nvcc, "-shared", *ARCH_FLAGS, "-o", str(LIB), *map(str, objs), "-lcudart", "-lcuda"]
    if verbose:
        print(" ".join(link))
    subprocess.run(link, check=True)
    if embedded_digest() != digest:
        raise RuntimeError("built library does not carry the source digest (runtime.cu: mdi_build_digest)")
    if STAMP.exists():
        STAMP.unlink()  # superseded by the embedded digest
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(path)
