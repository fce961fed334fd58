#!/usr/bin/env python3
"""Prefill (time to first token) of one Llama-3-8B stage slice: tcgen05 GEMMs + tcgen05 attention vs the same
GEMMs with SDPA attention, and the isolated attention kernels vs SDPA.  Writes gpurun_out/prefill_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402
from mdi_llm_b200.models.config import Config  # noqa: E402
from mdi_llm_b200.models.gpt import build_rope_cache  # noqa: E402
from mdi_llm_b200.models.stage import build_stage  # noqa: E402
from mdi_llm_b200.parallel.engine import FusedStage  # noqa: E402
from mdi_llm_b200.utils.checkpoint import random_init_stage_  # noqa: E402

ops.require()


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"attention": [], "stage": []}
H, G, hs, S = 32, 8, 128, 4096
cos, sin = build_rope_cache(S, hs, device=torch.device("cuda"))
cos, sin = cos.float().contiguous(), sin.float().contiguous()
pool = torch.zeros(1, 2, G, S, hs, device="cuda", dtype=torch.bfloat16)
for T in (128, 512, 2048, 4096):
    qkv = (torch.randn(T, (H + 2 * G) * hs, device="cuda") * 0.5).bfloat16()
    ops.set_prefill_attn_pipe(False)
    ms = timeit(lambda: ops.attn_prefill(qkv, cos, sin, pool, 0, n_head=H, n_groups=G, head_size=hs, rope_n_elem=hs))
    ops.set_prefill_attn_pipe(True)
    ms_pipe = timeit(lambda: ops.attn_prefill(qkv, cos, sin, pool, 0, n_head=H, n_groups=G, head_size=hs, rope_n_elem=hs))
    row_extra = {}
    if True:  # two softmax warpgroups (mode 2, the default)
        ops.set_prefill_attn_pipe(2)
        ms2 = timeit(lambda: ops.attn_prefill(qkv, cos, sin, pool, 0, n_head=H, n_groups=G, head_size=hs, rope_n_elem=hs))
        row_extra = {"tcgen05_pipe2_ms": round(ms2, 4)}
    ops.set_prefill_attn_pipe(int(os.environ.get("MDI_PREFILL_ATTN_PIPE", "1")))
    q = torch.randn(1, H, T, hs, device="cuda").bfloat16()
    k = torch.randn(1, H, T, hs, device="cuda").bfloat16()
    ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, k, is_causal=True))
    flops = 4.0 * H * hs * T * T / 2  # causal
    out["attention"].append({"T": T, "tcgen05_ms": round(ms, 4), "tcgen05_tflops": round(flops / ms / 1e9, 1),
                             "tcgen05_pipelined_ms": round(ms_pipe, 4), "tcgen05_pipelined_tflops": round(flops / ms_pipe / 1e9, 1),
                             "sdpa_ms (attention only, no rope/split/cache)": round(ms_ref, 4), **row_extra})
    print(out["attention"][-1], flush=True)

cfg = Config.from_name("Llama-3-8B", n_layer=4, block_size=4096)
st_mod = build_stage(cfg, "starter", 4, meta=True)
random_init_stage_(st_mod, "cuda", torch.bfloat16)
fs = FusedStage(st_mod, n_slots=1, max_seq_length=4096)
fs.warmup()
for T in (64, 1024, 2048):
    ids = torch.randint(0, cfg.vocab_size, (1, T), device="cuda")
    pos = torch.arange(T, device="cuda")
    row = {"T": T, "layers": 4}
    for mode in ("tcgen05", "sdpa"):
        fs.prefill_attn = mode
        row[f"{mode}_ms"] = round(timeit(lambda: fs.prefill(ids, pos, 0), iters=5), 3)
    out["stage"].append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/prefill_bench.json", "w"), indent=1)
