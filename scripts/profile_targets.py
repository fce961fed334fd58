#!/usr/bin/env python3
"""Small deterministic workloads for `ncu --set full` captures (see profiles/README.md):
   decode: the Llama-3-8B gate/up weight-streaming kernel (+RMSNorm, SiLU*mul epilogue), 235 MB of weights
   gemm  : the tcgen05 CTA-pair prefill GEMM 2048 x 14336 x 4096
   attn  : decode attention at 470 live positions (cluster merge), Llama-3-8B heads
   attnp : tcgen05 prefill attention, 2048-token prompt, Llama-3-8B heads
   gemm_fp8: the block-scaled fp8 GEMM (tcgen05 kind::f8f6f4) 2048 x 6144 x 4096"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
ops.require()
ctx = torch.zeros(ops.CTX_INTS, dtype=torch.int32, device="cuda")
if what == "decode":
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    C, I = 4096, 14336
    Ws = [torch.randn(I, C, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(4)]
    x = torch.randn(C, device="cuda", dtype=torch.bfloat16)
    nw = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(I, device="cuda", dtype=torch.bfloat16)
    for i in range(6):
        ops.linear_decode(Ws[(2 * i) % 4], x, y, ctx, W2=Ws[(2 * i + 1) % 4], norm_w=nw, act="silu_gate", variant=variant, ctas_per_sm=3)
elif what == "gemm":
    M, N, K = 2048, 14336, 4096
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(4):
        ops.gemm(a, w, block_n=512)  # the CTA-pair kernel (cta_group::2)
elif what == "attn":
    H, G, hs, S = 32, 8, 128, 4096
    q = torch.randn(H * hs, device="cuda").bfloat16()
    kv = torch.randn(1, 2, G, S, hs, device="cuda").bfloat16()
    y = torch.zeros(H * hs, device="cuda", dtype=torch.bfloat16)
    part = torch.zeros(H * 40 * (hs + 2), device="cuda")
    tickets = torch.zeros(G, dtype=torch.int32, device="cuda")
    ctx[1] = int(sys.argv[2]) - 1 if len(sys.argv) > 2 else 469  # live context: 470 positions -> 4 spans, cluster (DSMEM) merge
    for _ in range(4):
        ops.attn_decode(q, kv, y, part, tickets, ctx, n_head=H, n_groups=G, head_size=hs, max_seq=S, n_split=40)
elif what == "gemm_fp8":
    from mdi_llm_b200.utils.quantize import quantize_fp8_block

    M, N, K = 2048, 6144, 4096
    a = torch.randn(M, K, device="cuda").bfloat16()
    q, s = quantize_fp8_block((torch.randn(N, K, device="cuda") * 0.02).bfloat16())
    a8, a_st = ops.quantize_rows_fp8(a)
    st = s.t().contiguous()
    for _ in range(4):
        ops.gemm_fp8(a8, a_st, q.view(torch.uint8), st)
torch.cuda.synchronize()
print("done", what)

if what == "attnp":
    from mdi_llm_b200.models.gpt import build_rope_cache

    H, G, hs, S, T = 32, 8, 128, 2048, 2048
    qkv = (torch.randn(T, (H + 2 * G) * hs, device="cuda") * 0.5).bfloat16()
    cos, sin = build_rope_cache(S, hs, device=torch.device("cuda"))
    cos, sin = cos.float().contiguous(), sin.float().contiguous()
    pool = torch.zeros(1, 2, G, S, hs, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        ops.attn_prefill(qkv, cos, sin, pool, 0, n_head=H, n_groups=G, head_size=hs, rope_n_elem=hs)
torch.cuda.synchronize()
