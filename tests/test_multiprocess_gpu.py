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


@pytest.mark.parametrize("world", [2, 4])
def test_ipc_ring_third_layer_boundaries(world):
    """Stage boundaries between attention | gate/up | down units: a boundary after gate/up carries [x | h]
    (decode: row copy by the last CTA of the gate/up kernel; prefill: copy + in-kernel flag release)."""
    _run(world, "device", 29680 + world, "8", "third")


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


def _api_tokens(tmp_path, ck, n_nodes, tag, *starter_flags):
    """One run through the public CLIs: `launch` spawns `secondary` x (N-1) and `starter`; returns the JSON the
    starter wrote with --tokens-out."""
    out = tmp_path / f"tokens_{tag}.json"
    env = dict(os.environ, MDI_LOGS_DIR=str(tmp_path / "logs"), MDI_IMG_DIR=str(tmp_path / "img"))
    cmd = [sys.executable, "-m", "mdi_llm_b200.cli.launch", "--ckpt", str(ck), "--n-nodes", str(n_nodes), "--",
           "--n-samples", str(n_nodes + 1), "--n-tokens", "12", "--prompt", "Hello there", "--greedy", "--sequence-length", "128",
           "--tokens-out", str(out), *starter_flags]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and out.is_file(), p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return json.loads(out.read_text())


@pytest.fixture
def fused_ckpt(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_engine_gpu import _cfg
    from mdi_llm_b200.utils.checkpoint import write_random_checkpoint

    return write_random_checkpoint(tmp_path / "custom" / "TinyFused", _cfg(n_layer=8), dtype=torch.bfloat16, seed=3)


@pytest.mark.parametrize("world,flags", [(2, ()), (2, ("--partition", "third", "--decode-mode", "host")), (4, ("--partition", "half",)),
                                         (8, ("--partition", "third"))])
def test_starter_secondary_clis_device_ring_matches_single_gpu(tmp_path, fused_ckpt, world, flags):
    """The product path THROUGH THE PUBLIC API: `starter` + `secondary` CLI processes (one per GPU), HTTP control
    plane, CUDA-IPC handles exchanged at POST /init, fused NVLink hops — token-identical to one GPU, greedy."""
    if torch.cuda.device_count() < world:
        pytest.skip("not enough GPUs")
    one = _api_tokens(tmp_path, fused_ckpt, 1, "n1")
    many = _api_tokens(tmp_path, fused_ckpt, world, f"n{world}", *flags)
    assert one["transport"] == "p2p" and many["transport"] == "p2p" and many["n_nodes"] == world
    n = min(len(one["tokens"]), len(many["tokens"]))
    assert n >= 2
    for i in range(n):
        assert one["tokens"][str(i)] == many["tokens"][str(i)], (i, one["tokens"][str(i)], many["tokens"][str(i)])
    assert len(many["tok_time"]) == 1 + 12 * (world + 1)  # one timeline point per generated token (device clock or host clock)
    assert all(b[1] >= a[1] for a, b in zip(many["tok_time"], many["tok_time"][1:]))


def test_ring_abort_when_a_node_dies(tmp_path, fused_ckpt):
    """Freeze one secondary mid-generation: the starter's watchdog trips, the poison flag drains the ring and the
    API raises instead of hanging."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_mp_abort_worker.py"), str(fused_ckpt)], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert "ABORT_RESULT ok" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
