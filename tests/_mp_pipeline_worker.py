"""torchrun worker: N-process device pipeline (CUDA IPC hops) vs the single-GPU pipeline on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_engine_gpu import _cfg, _half_stages, _stages, _third_stages  # noqa: E402
from mdi_llm_b200.parallel.pipeline import DevicePipeline  # noqa: E402
from mdi_llm_b200.parallel.scheduler import SamplingParams  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mode = sys.argv[1] if len(sys.argv) > 1 else "device"
    n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg = _cfg(n_layer=max(6, world))
    if len(sys.argv) > 3 and sys.argv[3] == "half":  # boundaries inside layers (attention | MLP units)
        units = [3] + [2] * (world - 2) + [2 * cfg.n_layer - 3 - 2 * (world - 2)]
        stages = _half_stages(cfg, units, [f"cuda:{local}"] * world)
    elif len(sys.argv) > 3 and sys.argv[3] == "third":  # boundaries between attention | gate/up | down units
        n_units = 3 * cfg.n_layer
        units = [4] + [5] * (world - 2) + [n_units - 4 - 5 * (world - 2)]  # cuts after a gate/up unit AND after attention units
        stages = _third_stages(cfg, units, [f"cuda:{local}"] * world)
    else:
        _, stages = _stages(cfg, world, device=[f"cuda:{local}"] * world)  # every rank materialises only its own use
    stage = stages[rank]
    n_samples = world + 1
    prompts = [torch.tensor([1, 10 + i, 20, 30 + i, 7][: 4 + i % 2]) for i in range(n_samples)]
    pipe = DevicePipeline(stage, rank, world, n_samples=n_samples, max_seq_length=64, sampling=SamplingParams.greedy(),
                          wait_max_cycles=8 * 10 ** 9)
    pipe.connect_distributed()

    def sync():
        torch.cuda.synchronize()
        dist.barrier()

    out = pipe.generate(prompts, n_new, sync=sync, mode=mode)
    ok = True
    detail = {}
    if rank == 0:
        _, (st1,) = _stages(cfg, 1, device=f"cuda:{local}")
        single = DevicePipeline(st1, 0, 1, n_samples=n_samples, max_seq_length=64, sampling=SamplingParams.greedy())
        ref = single.generate(prompts, n_new)
        for i in range(n_samples):
            if not torch.equal(out[i], ref[i]):
                ok = False
                detail[i] = {"got": out[i].tolist(), "ref": ref[i].tolist()}
        detail["flags"] = pipe.stage.flags.tolist()
        detail["hidden_in_absmax"] = pipe.stage.hidden_in.float().abs().amax(dim=1).tolist()
        print("MP_RESULT " + json.dumps({"ok": ok, "world": world, "mode": mode, "detail": detail}), flush=True)
    dist.barrier()
    pipe.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
