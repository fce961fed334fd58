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


@pytest.mark.skipif(not (ROOT / "baseline/_ref/sub/model_dist.py").is_file(), reason="baseline/_ref not installed")
def test_reference_arm_two_nodes_cpu(tmp_path):
    """Two reference nodes (starter + secondary) as two torchrun ranks on CPU: the checkpoint is split with the
    reference's own split_and_store, the nodes talk over its loopback sockets, rank 0 prints the one JSON line."""
    from conftest import free_ports

    env = dict(os.environ, MDI_REF_DEVICE="cpu", MDI_REF_MODEL="pythia-160m", MDI_REF_CKPT_DIR=str(tmp_path))
    (port,) = free_ports(1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3",
           "--prompt-len", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["value"] > 0, out
    chunk_dir = next(tmp_path.glob("custom/*/chunks/2nodes"))
    assert (chunk_dir / ".complete").is_file() and (chunk_dir / "model_secondary0.pth").is_file()


@pytest.mark.skipif(not (ROOT / "baseline/_ref/sub/model_dist.py").is_file(), reason="baseline/_ref not installed")
@pytest.mark.parametrize("impl,world", [("reference-nccl", 2), ("reference-table", 4)])
def test_reference_variants_cpu(tmp_path, impl, world):
    """The two comparator variants of the reference arm on CPU (gloo stands in for NCCL): ``reference-nccl`` swaps
    ``sub.connections`` for torch.distributed send/recv and leaves the rest of the reference stock;
    ``reference-table`` injects a partition-table entry for a node count the reference's table lacks
    (4 nodes x 12 layers here, as 8 nodes x 32 layers on the GPU box)."""
    from conftest import free_ports

    env = dict(os.environ, MDI_REF_DEVICE="cpu", MDI_REF_MODEL="pythia-160m", MDI_REF_CKPT_DIR=str(tmp_path))
    (port,) = free_ports(1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "bench.py"), "--impl", impl, "--gpus", str(world), "--steps", "3", "--warmup", "3",
           "--prompt-len", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    out = json.loads(lines[0])
    assert out["impl"] == impl and out["n_gpus"] == world and out.get("value", 0) > 0, out
    if impl == "reference-table":
        assert "INJECTED" in out["details"]["partition"]
    else:
        assert "torch.distributed" in out["details"]["transport"]
    assert set(out["config"]) == {"model", "global_batch", "seq_len", "prompt_len", "parallelism", "tokens_per_step", "sampling",
                                  "l2_policy"}  # the keys the ours arm prints too
