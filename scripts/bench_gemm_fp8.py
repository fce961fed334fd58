#!/usr/bin/env python3
"""Block-scaled fp8 prefill GEMM (tcgen05 kind::f8f6f4) on Llama-3-8B shapes against (a) our bf16 tcgen05 GEMM on the
dequantised weights — what fp8 prefill cost before: expand the checkpoint, then GEMM — and (b) cuBLAS bf16.  CUDA events,
warm, weights cycled through a ring larger than L2.  Writes gpurun_out/gemm_fp8_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402
from mdi_llm_b200.utils.quantize import dequantize_fp8_block, quantize_fp8_block  # noqa: E402

ops.require()
dev = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


rows = []
for (M, N, K, gated) in [(2048, 6144, 4096, False), (2048, 4096, 4096, False), (2048, 14336, 4096, True), (2048, 4096, 14336, False),
                         (8192, 14336, 4096, True), (8192, 4096, 14336, False), (512, 14336, 4096, True), (64, 4096, 4096, False)]:
    ring = max(2, int(200e6 // (N * K)) + 1)
    a = torch.randn(M, K, device=dev).bfloat16()
    qs = [quantize_fp8_block((torch.randn(N, K, device=dev) * 0.02).bfloat16()) for _ in range(ring)]
    w8 = [q.view(torch.uint8) for q, _ in qs]
    st = [s.t().contiguous() for _, s in qs]
    q2 = quantize_fp8_block((torch.randn(N, K, device=dev) * 0.02).bfloat16()) if gated else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * M * N * K * (2 if gated else 1)
    a8, a_st = ops.quantize_rows_fp8(a)
    kw = dict(w2_8=q2[0].view(torch.uint8), w2_scale_t=q2[1].t().contiguous()) if gated else {}
    t_mm = timeit(lambda i: ops.gemm_fp8(a8, a_st, w8[i % ring], st[i % ring], out=out, **kw))
    t_q = timeit(lambda i: ops.quantize_rows_fp8(a))
    wb = dequantize_fp8_block(*qs[0])
    w2b = dequantize_fp8_block(*q2) if gated else None
    t_bf16 = timeit(lambda i: ops.gemm(a, wb, out=out, block_n=256 if (M > 128 and not gated) else 128, w2=w2b))
    t_deq = timeit(lambda i: dequantize_fp8_block(*qs[i % ring]), iters=5)
    rows.append({"M": M, "N": N, "K": K, "gated": gated, "fp8_gemm_us": round(t_mm * 1e6, 1), "act_quant_us": round(t_q * 1e6, 1),
                 "fp8_tflops": round(flops / t_mm / 1e12, 1), "fp8_tflops_incl_quant": round(flops / (t_mm + t_q) / 1e12, 1),
                 "bf16_tcgen05_us": round(t_bf16 * 1e6, 1), "bf16_tflops": round(flops / t_bf16 / 1e12, 1),
                 "dequant_weights_us": round(t_deq * 1e6 * (2 if gated else 1), 1),
                 "speedup_vs_dequant_then_bf16": round((t_bf16 + t_deq * (2 if gated else 1)) / (t_mm + t_q), 2)})
    print(rows[-1], flush=True)
    del qs, w8, st
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_fp8_bench.json", "w"), indent=1)
