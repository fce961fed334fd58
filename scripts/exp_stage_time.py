#!/usr/bin/env python3
"""Per-step time of ONE pipeline stage on one GPU, free-running (incoming flags pre-satisfied, hop stored locally):
what a stage costs when it never waits — the number the partition planner needs — as a function of its layer
count, its role and the number of steps captured per CUDA graph.  CUDA events, 300 steps after 50 warm-up."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402
from mdi_llm_b200.models.config import Config  # noqa: E402
from mdi_llm_b200.models.stage import build_stage  # noqa: E402
from mdi_llm_b200.parallel.pipeline import DevicePipeline  # noqa: E402
from mdi_llm_b200.parallel.scheduler import SamplingParams  # noqa: E402
from mdi_llm_b200.utils.checkpoint import random_init_stage_  # noqa: E402


def measure(model, role, layers, k, n_samples, ctx, steps=300, warm=50, weights="bf16", env=None, **stage_kw):
    saved = {}
    for key, val in (env or {}).items():
        saved[key] = os.environ.get(key)
        os.environ[key] = val
    try:
        return _measure(model, role, layers, k, n_samples, ctx, steps, warm, weights, **stage_kw)
    finally:
        for key, val in saved.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val


def _measure(model, role, layers, k, n_samples, ctx, steps=300, warm=50, weights="bf16", **stage_kw):
    cfg = Config.from_name(model, n_layer=max(layers, 1), block_size=2048)
    st = build_stage(cfg, role, layers, meta=True, **stage_kw)
    random_init_stage_(st, "cuda", torch.bfloat16, seed=1)
    rank, world = (0, 1) if role == "starter" else (1, 2)
    pipe = DevicePipeline(st, rank, world, n_samples=n_samples, max_seq_length=ctx + steps + warm + 80, sampling=SamplingParams(seed=1),
                          exportable=False, weight_dtype=weights, free_bf16=weights == "fp8")
    pipe.steps_per_graph = k
    prompts = [torch.randint(0, cfg.vocab_size, (ctx,), dtype=torch.int32) for _ in range(n_samples)]
    rounds = (steps + warm) // n_samples + 2
    pipe.prepare(prompts, rounds + 1)
    if role != "starter":
        pipe.stage.flags.fill_(1 << 30)  # every incoming message "already there": the stage never waits
        pipe.next_hop = pipe.stage.hop_self.__class__(pipe.stage.out_local.data_ptr(), torch.zeros(n_samples, dtype=torch.int32, device="cuda").data_ptr())
        pipe._keep = pipe.next_hop
    else:
        pipe.prefill()
    torch.cuda.synchronize()
    pipe.decode_rounds(warm // n_samples)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r = steps // n_samples
    e0.record()
    pipe.decode_rounds(r)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (r * n_samples)
    status = pipe.stage.status[:2].tolist()
    del pipe, st
    torch.cuda.empty_cache()
    return round(us, 2), status


def trace_boundary(model, layers, ctx, use_pdl_adv=True):
    """Timeline of two consecutive steps of a free-running secondary: what happens between the last kernel of
    step t and the first kernel of step t+1."""
    from mdi_llm_b200.parallel.engine import HopTarget

    cfg = Config.from_name(model, n_layer=layers, block_size=2048)
    st = build_stage(cfg, "secondary:0", layers, meta=True)
    random_init_stage_(st, "cuda", torch.bfloat16, seed=1)
    pipe = DevicePipeline(st, 1, 2, n_samples=2, max_seq_length=ctx + 64, sampling=SamplingParams(seed=1), exportable=False)
    pipe.prepare([torch.zeros(ctx, dtype=torch.int32)] * 2, 16)
    fs = pipe.stage
    fs.flags.fill_(1 << 30)
    sink_flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    hop = HopTarget(fs.out_local.data_ptr(), sink_flags.data_ptr())

    def two():
        for j in range(2):
            ops.advance_step(fs.ctx, fs.state, fs.pos_arr, 2, False, use_pdl=use_pdl_adv and j > 0)
            fs.enqueue_blocks(hop, wait_input=True)

    for _ in range(3):
        two()
    torch.cuda.synchronize()
    rows = fs.trace_step(two)
    prev = None
    print(f"{'kernel':14s} {'entry':>8s} {'ready':>8s} {'staged':>8s} {'1st exit':>9s} {'last exit':>9s} {'gap':>6s}")
    for r in rows:
        gap = r["entry"] - prev if prev is not None else 0.0
        print(f"{r['kernel']:14s} {r['entry']:8.2f} {r['ready']:8.2f} {(r['staged'] or 0):8.2f} {r['first_exit']:9.2f} {r['last_exit']:9.2f} {gap:6.2f}")
        prev = r["last_exit"]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Llama-3-8B")
    ap.add_argument("--ctx", type=int, default=460)
    ap.add_argument("--cases", default="secondary:4:1,secondary:4:8,secondary:2:1,secondary:2:8,starter:2:1,starter:2:8,starter:32:1,starter:32:8")
    ap.add_argument("--n-samples", type=int, default=8)
    ap.add_argument("--out", default="")
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    ops.require()
    if a.trace:
        print("== two steps, advance_step with PDL edge ==")
        trace_boundary(a.model, 2, a.ctx, True)
        print("== two steps, advance_step without PDL edge (graph-boundary-like) ==")
        trace_boundary(a.model, 2, a.ctx, False)
    rows = []
    for case in a.cases.split(","):
        f = case.split(":")
        role, layers, k = f[0], int(f[1]), int(f[2])
        env = {}
        if len(f) > 3 and f[3]:
            for flag in f[3].split("+"):
                key, _, val = flag.partition("=")
                env[key] = val
        role_name = "starter" if role == "starter" else "secondary:0"
        us, status = measure(a.model, role_name, layers, k, 1 if (role == "starter" and layers >= 16) else a.n_samples, a.ctx, env=env)
        rows.append({"role": role, "layers": layers, "steps_per_graph": k, "env": env, "us_per_step": us, "status": status})
        print(rows[-1], flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)
