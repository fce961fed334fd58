#!/usr/bin/env python3
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels with gpurun).

The reference ships no packaging metadata (no setup.py / pyproject), so a plain
``pip install /root/reference`` cannot work.  As the task allows, we install from a copy under
/tmp to which ONLY a minimal ``setup.py`` is added (packages = the reference's own ``sub`` tree plus
its top-level scripts as py_modules); source files are byte-identical.  Dependencies that are not
in the offline wheelhouse (cherrypy, accelerate, matplotlib) are provided at run time by the
environment shims in baseline/shims/ — see DESIGN.md.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MDI_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

SETUP = '''
from setuptools import setup, find_packages
setup(name="mdi-llm-reference", version="0.0.0", package_dir={"": "src"},
      packages=find_packages("src"), py_modules=["starter", "secondary", "sample"],
      package_data={"": ["*.json", "*.txt"]}, include_package_data=True)
'''


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"reference source {SRC} not found", file=sys.stderr)
        return 2
    tmp = tempfile.mkdtemp(prefix="mdi_ref_")
    work = os.path.join(tmp, "reference")
    shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "assets", "old"))
    with open(os.path.join(work, "setup.py"), "w") as f:
        f.write(SETUP)
    shutil.rmtree(DST, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", DST, work]
    print(" ".join(cmd))
    rc = subprocess.call(cmd)
    # settings_distr/*.json and prompts are data files next to the scripts: copy verbatim
    for d in ("settings_distr", "prompts"):
        s = os.path.join(SRC, "src", d)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DST, d), dirs_exist_ok=True)
    shutil.rmtree(tmp, ignore_errors=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
