#!/usr/bin/env python3
"""First contact of the CTA-pair GEMM with hardware: tiny shapes first, each in this process under the caller's
`timeout` (a protocol bug shows up as a hang, not as a wrong number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402

ops.require()
for (M, N, K) in [(256, 256, 64), (256, 256, 256), (512, 512, 1024), (2048, 6144, 4096), (2048, 4096, 14336), (8192, 14336, 4096)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = ops.gemm(a, w, block_n=512)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    err = (out.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o1 = torch.empty_like(out)
    ops.gemm(a, w, out=o1, block_n=256)
    torch.cuda.synchronize()
    res = {}
    for name, bn in (("pair", 512), ("single256", 256)):
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, out=o1, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(10 * 2 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)
    e0.record()
    for _ in range(10):
        torch.mm(a, w.T, out=o1)
    e1.record()
    torch.cuda.synchronize()
    res["cublas"] = round(10 * 2 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)
    print(f"{M}x{N}x{K}: rel err {err:.4f}  TFLOP/s {res}", flush=True)
