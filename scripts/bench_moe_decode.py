#!/usr/bin/env python3
"""Decode step of a Mixtral-8x7B-shaped pipeline stage (top-2 of 8 SwiGLU experts per token): the fused engine
(device router + expert passes through pointer tables, one CUDA graph per 8 steps) against the eager modules
(torch.topk / torch.where routing, the reference's formulation) on the same GPU, and against the bytes a step has
to stream (attention + the two chosen experts of every layer) at the measured HBM copy bandwidth.
CUDA events, warm; consecutive tokens pick different experts, so the expert matrices (2.8 GB per layer) do not
stay in L2.  Writes gpurun_out/moe_decode_bench_v<variant>.json."""
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


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Mixtral-8x7B-v0.1")
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--ctx", type=int, default=64)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--n-samples", type=int, default=8)
    ap.add_argument("--prefill", type=int, default=1024, help="prompt length of the prefill comparison")
    ap.add_argument("--moe-variant", type=int, default=0, help="expert passes: 0 register-streamed, 2 bulk-copy ring")
    a = ap.parse_args()
    ops.require()
    ops.set_moe_variant(a.moe_variant)
    cfg = Config.from_name(a.model, n_layer=a.layers, block_size=4096)
    st = build_stage(cfg, "secondary:0", a.layers, meta=True)
    random_init_stage_(st, "cuda", torch.bfloat16, seed=1)
    n = a.n_samples
    warm = 5 * n
    pipe = DevicePipeline(st, 1, 2, n_samples=n, max_seq_length=max(a.ctx + a.steps + warm + 80, a.prefill + 16), sampling=SamplingParams(seed=1),
                          exportable=False)
    pipe.prepare([torch.zeros(a.ctx, dtype=torch.int32) for _ in range(n)], (a.steps + warm) // n + 3)
    fs = pipe.stage
    fs.flags.fill_(1 << 30)  # every incoming message "already there": the stage never waits
    fs.hidden_in.copy_((torch.randn(n, cfg.n_embd, device="cuda") * 0.5).bfloat16())  # one distinct row per slot
    sink = torch.zeros(n, dtype=torch.int32, device="cuda")
    pipe.next_hop = fs.hop_self.__class__(fs.out_local.data_ptr(), sink.data_ptr())
    pipe.decode_rounds(warm // n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.decode_rounds(a.steps // n)
    e1.record()
    torch.cuda.synchronize()
    fused_us = e0.elapsed_time(e1) * 1e3 / (a.steps // n * n)
    status = fs.status[:2].tolist()
    launches = pipe.n_kernel_launches

    # eager modules, same stage, same shapes: one token per forward on the slot KV pool
    xs = [(torch.randn(1, 1, cfg.n_embd, device="cuda") * 0.5).bfloat16() for _ in range(n)]
    pos = torch.tensor([a.ctx], device="cuda")
    with torch.inference_mode():
        for i in range(2 * n):
            st(xs[i % n], pos, slot=i % n)
        torch.cuda.synchronize()
        e0.record()
        for i in range(4 * n):
            st(xs[i % n], pos, slot=i % n)
        e1.record()
        torch.cuda.synchronize()
    eager_us = e0.elapsed_time(e1) * 1e3 / (4 * n)

    # the prompt: routed MLPs on the tcgen05 GEMMs (per-expert rows) + tcgen05 attention, against the eager modules
    T = a.prefill
    xp = (torch.randn(1, T, cfg.n_embd, device="cuda") * 0.5).bfloat16()
    ppos = torch.arange(T, device="cuda")
    pf = {}
    with torch.inference_mode():
        for name, fn in (("fused", lambda: fs.prefill(xp, ppos, 0)), ("eager", lambda: st(xp, ppos, slot=0))):
            fs.set_ctx(0, T - 1)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            pf[name] = e0.elapsed_time(e1) / 5
        # accuracy: a routed model is discontinuous where two experts' router logits nearly tie, so rows whose routing
        # differs between the two runs (bf16 noise of the attention / norm in front of the router) are counted apart
        routed = []
        hooks = [blk.mlp.gate.register_forward_hook(lambda _m, _i, o: routed.append(torch.topk(o, cfg.n_expert_per_token, dim=-1).indices.sort(-1).values))
                 for blk in st.transformer.h]
        out_e = st(xp, ppos, slot=0).float()[0]
        out_f = fs.prefill(xp, ppos, 0).float()[0]
        for hk in hooks:
            hk.remove()
        L = len(st.transformer.h)
        same = torch.ones(T, dtype=torch.bool, device="cuda")
        for li in range(L):
            same &= (routed[li].reshape(T, -1) == routed[L + li].reshape(T, -1)).all(-1)
        row_err = (out_e - out_f).abs().max(-1).values
        scale = out_e.abs().max().item()
        acc = {"out_scale": round(scale, 3), "rows_same_routing": int(same.sum()), "rows_routing_flipped": int((~same).sum()),
               "max_err_same_routing": round(row_err[same].max().item(), 4) if bool(same.any()) else None,
               "mean_err_same_routing": round((out_e - out_f).abs()[same].mean().item(), 5) if bool(same.any()) else None,
               "max_err_flipped": round(row_err[~same].max().item(), 4) if bool((~same).any()) else None}

    C, I = cfg.n_embd, cfg.intermediate_size
    step_bytes = a.layers * 2 * (cfg.qkv_size * C + cfg.attn_out_dim * C + cfg.n_expert * C + cfg.n_expert_per_token * 3 * C * I)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass
    bw = float(peaks.get("hbm_gbs", 6577.0)) if isinstance(peaks, dict) else 6577.0
    out = {"model": a.model, "moe_variant": a.moe_variant, "layers": a.layers, "ctx": a.ctx, "n_samples": n, "fused_us_per_step": round(fused_us, 2),
           "eager_us_per_step": round(eager_us, 2), "speedup_vs_eager": round(eager_us / fused_us, 2),
           "weight_bytes_per_step": step_bytes, "hbm_floor_us": round(step_bytes / (bw * 1e3), 2),
           "fraction_of_hbm_floor": round(step_bytes / (bw * 1e3) / fused_us, 3), "hbm_gbps_assumed": bw,
           "kernel_launches": launches, "status": status,
           "prefill": {"tokens": T, "fused_ms": round(pf["fused"], 3), "eager_ms": round(pf["eager"], 3),
                       "speedup_vs_eager": round(pf["eager"] / pf["fused"], 2), "accuracy_vs_eager": acc,
                       "flops": a.layers * 2 * T * (cfg.qkv_size * C + cfg.attn_out_dim * C + cfg.n_expert_per_token * 3 * C * I)}}
    out["prefill"]["tflops_fused"] = round(out["prefill"]["flops"] / (pf["fused"] * 1e-3) / 1e12, 1)
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/moe_decode_bench_v{a.moe_variant}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
