"""torchrun worker (CPU, gloo): host-driven recurrent pipeline over the torch.distributed transport."""
import json
import os
import sys
import threading

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mdi_llm_b200.models.config import Config  # noqa: E402
from mdi_llm_b200.models.gpt import GPT  # noqa: E402
from mdi_llm_b200.models.partition import split_parameters  # noqa: E402
from mdi_llm_b200.models.stage import build_stage  # noqa: E402
from mdi_llm_b200.parallel.scheduler import EagerStageRunner, SamplingParams, secondary_loop, starter_loop  # noqa: E402
from mdi_llm_b200.parallel.transport.nccl_p2p import TorchDistTransport, make_edge_groups  # noqa: E402
from mdi_llm_b200.utils.checkpoint import random_state_dict  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    cfg = Config.from_name("tiny-llama-1.1b", n_layer=5, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=128,
                           vocab_size=300, padded_vocab_size=320, block_size=64)
    sd = random_state_dict(cfg, dtype=torch.float32)
    full = {k: v.clone() for k, v in sd.items()}
    chunks, info = split_parameters(sd, world)
    role = "starter" if rank == 0 else f"secondary:{rank - 1}"
    stage = build_stage(cfg, role, info["plan"][rank], meta=True)
    stage.load_weights(chunks["starter"] if rank == 0 else chunks["secondary"][rank - 1])
    runner = EagerStageRunner(stage)
    tr = TorchDistTransport(rank, world, torch.device("cpu"), make_edge_groups(world))
    running = threading.Event()
    running.set()
    tr.launch()
    prompts = [torch.tensor([1, 10 + i, 20, 30 + i]) for i in range(world + 1)]
    ok = True
    if rank == 0:
        res = starter_loop(runner, tr, prompts, 6, SamplingParams.greedy(), running, n_nodes=world)
        m = GPT(cfg)
        m.load_state_dict(full)
        m.eval()
        for i, p in enumerate(prompts):
            m.clear_kv_cache()
            ok &= res.samples[i].tolist() == m.generate(p, len(p) + 6, temperature=0.0, top_p=0.0).tolist()
        print("TR_RESULT " + json.dumps({"ok": bool(ok), "world": world, "bytes_sent": tr.stats["bytes_sent"]}), flush=True)
        stop = torch.ones(1)
    else:
        t = threading.Thread(target=secondary_loop, args=(runner, tr, running), kwargs={"recv_timeout": 0.1}, daemon=True)
        t.start()
        stop = torch.zeros(1)
    dist.broadcast(stop, 0)  # generation finished on the starter
    running.clear()
    dist.barrier()
    tr.shutdown()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
