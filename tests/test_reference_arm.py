"""``bench.py --impl reference`` plumbing: the unmodified reference (baseline/_ref) runs through its
own GPTDistributed API and stdout carries exactly one JSON line (CPU, small model)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not (ROOT / "baseline/_ref/sub/model_dist.py").is_file(), reason="baseline/_ref not installed")
def test_reference_arm_cpu(tmp_path):
    from conftest import free_ports

    env = dict(os.environ, MDI_REF_DEVICE="cpu", MDI_REF_MODEL="pythia-160m", MDI_REF_CKPT_DIR=str(tmp_path),
               MASTER_PORT=str(free_ports(1)[0]))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                        "--prompt-len", "8"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    out = json.loads(lines[0])
    assert out["impl"] == "reference"
    assert "unavailable" in out or out["value"] > 0


def test_reference_arm_unavailable_is_clean(tmp_path):
    """A node count the reference has no partition for reports `unavailable` and exits 0."""
    if not (ROOT / "baseline/_ref/sub/model_dist.py").is_file():
        pytest.skip("baseline/_ref not installed")
    env = dict(os.environ, MDI_REF_DEVICE="cpu", MDI_REF_MODEL="pythia-14m", MDI_REF_CKPT_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["impl"] == "reference" and "unavailable" in out
