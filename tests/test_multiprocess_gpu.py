"""Multi-process (one rank per GPU, CUDA-IPC peer buffers) device pipeline, launched with torchrun."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, mode, port, *extra):
    if torch.cuda.device_count() < world:
        pytest.skip("not enough GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_mp_pipeline_worker.py"), mode, *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("MP_RESULT ")]
    assert lines, f"worker produced no result\nstdout:\n{p.stdout[-2000:]}\nstderr:\n{p.stderr[-3000:]}"
    res = json.loads(lines[-1][len("MP_RESULT "):])
    assert res["ok"], res
    assert p.returncode == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ipc_ring_device_mode_matches_single_gpu(world):
    _run(world, "device", 29600 + world)


def test_ipc_ring_host_fed_mode_matches_single_gpu():
    _run(2, "host", 29650)


@pytest.mark.parametrize("world", [2, 4])
def test_ipc_ring_half_layer_boundaries(world):
    """Stage boundaries inside layers: o_proj carries the hop out, gate/up acquires it (fused prefill hop
    from the attention-output GEMM included)."""
    _run(world, "device", 29660 + world, "8", "half")


def test_train_ddp_nccl_two_gpus(tmp_path):
    """The trainer's DistributedDataParallel path over NCCL under torchrun (SURVEY #23) on two B200s:
    bf16 autocast, fused AdamW, gradient accumulation with no_sync, checkpoint written by rank 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cli import _train_dir

    data, ck = _train_dir(tmp_path)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", "-m", "mdi_llm_b200.cli.train", "--ckpt", str(ck), "--dataset", str(data), "--init", "scratch",
           "--max-iters", "20", "--batch-size", "8", "--grad-acc-steps", "2", "--ckpt-interval", "10", "--log-interval", "5",
           "--eval-iters", "2", "--learning-rate", "0.01", "--warmup-iters", "2"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "world 2" in p.stdout and "step 20:" in p.stdout and "cuda" in p.stdout
    losses = [float(l.split("val loss ")[1]) for l in p.stdout.splitlines() if "val loss" in l]
    assert losses[-1] < losses[0], losses
    assert (ck / "lit_model.pth").is_file()
