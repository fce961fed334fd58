"""Build A/B variants of the kernel library for same-box comparisons (``MDI_OPS_LIB=<so> python bench.py``).

usage: python scripts/build_ab.py NAME [file.cu=GITREV ...]
Builds mdi_llm_b200/ops/build/ab_NAME.so from the working tree, with the listed sources taken from a
git revision instead.
"""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "mdi_llm_b200/ops/csrc"
OUT = ROOT / "mdi_llm_b200/ops/build"


def main() -> None:
    name = sys.argv[1]
    over = dict(a.split("=") for a in sys.argv[2:])
    OUT.mkdir(exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for f in CSRC.iterdir():
            rev = over.get(f.name) or over.get("*")
            if rev:
                data = subprocess.run(["git", "show", f"{rev}:mdi_llm_b200/ops/csrc/{f.name}"], cwd=ROOT,
                                      capture_output=True, check=True).stdout
                (td / f.name).write_bytes(data)
            else:
                (td / f.name).write_bytes(f.read_bytes())
        so = OUT / f"ab_{name}.so"
        cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
               "-fPIC", "--use_fast_math", "-I", str(td), "-shared", "-o", str(so),
               *map(str, sorted(td.glob("*.cu"))), "-lcudart", "-lcuda"]
        subprocess.run(cmd, check=True)
        print(so)


if __name__ == "__main__":
    main()
