#!/usr/bin/env python3
"""Per-kernel timeline of one decode step of a Llama-3-8B stage, from the device-side tracer
(%globaltimer marks inside the kernels; no profiler attached).  Shows, per launch: when its first
CTA started, when its input was staged, first/last CTA exit — i.e. ramp, steady state, tail and the
gap to the next kernel."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402
from mdi_llm_b200.models.config import Config  # noqa: E402
from mdi_llm_b200.models.stage import build_stage  # noqa: E402
from mdi_llm_b200.parallel.engine import FusedStage  # noqa: E402
from mdi_llm_b200.utils.checkpoint import random_init_stage_  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--variant", type=int, default=2)
ap.add_argument("--no-pdl", action="store_true")
ap.add_argument("--pos", type=int, default=100)
ap.add_argument("--out", default="")
ap.add_argument("--detail", default="", help="comma-separated kernel names that get a per-CTA table, e.g. L1.qkv,L1.gate_up")
ap.add_argument("--ctas-per-sm", type=int, default=4)
a = ap.parse_args()
ops.require()
ops.set_linear_variant(a.variant)
cfg = Config.from_name("Llama-3-8B", n_layer=a.layers, block_size=1024)
st_mod = build_stage(cfg, "starter", a.layers, meta=True)
random_init_stage_(st_mod, "cuda", torch.bfloat16)
st = FusedStage(st_mod, n_slots=1, max_seq_length=512, use_pdl=not a.no_pdl, ctas_per_sm=a.ctas_per_sm)
st.warmup()
st.set_ctx(0, a.pos)


def step():
    st.enqueue_head(wait=False)
    st.enqueue_sample()
    st.enqueue_embed(from_tokens=True)
    st.enqueue_blocks(None, False)


for _ in range(3):
    step()
torch.cuda.synchronize()
rows = st.trace_step(step, detail=[d for d in a.detail.split(",") if d])  # "*" = every kernel
prev_end = None
print(f"{'kernel':14s} {'entry':>8s} {'ready':>8s} {'staged':>8s} {'1st exit':>9s} {'last exit':>9s} {'dur':>7s} {'gap':>6s} ctas")
for r in rows:
    gap = r["entry"] - prev_end if prev_end is not None else 0.0
    print(f"{r['kernel']:14s} {r['entry']:8.2f} {r['ready']:8.2f} {(r['staged'] or 0):8.2f} {r['first_exit']:9.2f} {r['last_exit']:9.2f} "
          f"{r['last_exit'] - r['entry']:7.2f} {gap:6.2f} {r['ctas']}")
    prev_end = r["last_exit"]
for r in rows:
    pc = r.get("per_cta")
    if not pc:
        continue
    import collections
    import statistics as S

    def q(xs, f):
        xs = sorted(xs)
        return xs[min(len(xs) - 1, int(f * len(xs)))]

    rel = lambda k: [c[k] - r["ready"] for c in pc]
    per_sm = collections.Counter(c["sm"] for c in pc)
    print(f"-- {r['kernel']}: {len(pc)} CTAs on {len(per_sm)} SMs (CTAs/SM min {min(per_sm.values())} max {max(per_sm.values())}); "
          f"times relative to first 'ready' (us)")
    for k in ("entry", "ready", "staged", "exit"):
        v = rel(k)
        print(f"   {k:7s} min {min(v):7.2f}  p10 {q(v, .1):7.2f}  p50 {q(v, .5):7.2f}  p90 {q(v, .9):7.2f}  max {max(v):7.2f}")
    st_dur = [c["staged"] - c["ready"] for c in pc]
    run = [c["exit"] - c["staged"] for c in pc]
    print(f"   staging (ready->staged) p50 {S.median(st_dur):.2f} max {max(st_dur):.2f};  stream+epilogue (staged->exit) p50 {S.median(run):.2f} max {max(run):.2f}")
    p6 = [c["p6"] - r["ready"] for c in pc if c.get("p6") is not None]
    if p6:
        print(f"   phase mark (main loop done) p50 {S.median(p6):.2f} max {max(p6):.2f}")
    by_cnt = collections.defaultdict(list)
    for c in pc:
        by_cnt[per_sm[c["sm"]]].append(c["exit"] - r["ready"])
    for n, v in sorted(by_cnt.items()):
        print(f"   SMs holding {n} CTA(s): exit p50 {S.median(v):.2f} max {max(v):.2f}  ({len(v)} CTAs)")
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)
