#!/usr/bin/env python3
"""tcgen05 prefill GEMM throughput on Llama-3-8B shapes vs cuBLAS (torch.mm) — CUDA events, warm, the
weights cycled through a ring larger than L2.  Writes gpurun_out/gemm_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402

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
# block_n 512 = the CTA-pair kernel (cta_group::2, 256 x 256 MMA tiles over a 2-CTA cluster)
for (M, N, K, bn, gated) in [(2048, 14336, 4096, 512, False), (2048, 14336, 4096, 256, False), (2048, 4096, 4096, 512, False),
                             (2048, 4096, 4096, 128, False), (2048, 6144, 4096, 512, False), (2048, 6144, 4096, 128, False),
                             (2048, 4096, 14336, 512, False), (2048, 4096, 14336, 256, False), (2048, 14336, 4096, 512, True),
                             (2048, 14336, 4096, 128, True), (8192, 14336, 4096, 512, True), (8192, 4096, 14336, 512, False),
                             (512, 14336, 4096, 512, False), (512, 14336, 4096, 256, False), (64, 14336, 4096, 128, True),
                             (64, 4096, 14336, 64, False)]:
    ring = max(2, int(200e6 // (N * K * 2)) + 1)
    a = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(ring)]
    w2 = (torch.randn(N, K, device=dev) * 0.02).bfloat16() if gated else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * M * N * K * (2 if gated else 1)
    t = timeit(lambda i: ops.gemm(a, ws[i % ring], out=out, block_n=bn, w2=w2))
    if gated:
        def ref(i):
            g = a @ ws[i % ring].T
            u = a @ w2.T
            return torch.nn.functional.silu(g) * u
    else:
        def ref(i):
            return torch.mm(a, ws[i % ring].T, out=out)
    t_ref = timeit(ref)
    rows.append({"M": M, "N": N, "K": K, "block_n": bn, "gated": gated, "ours_us": round(t * 1e6, 1),
                 "ours_tflops": round(flops / t / 1e12, 1), "cublas_us": round(t_ref * 1e6, 1),
                 "cublas_tflops": round(flops / t_ref / 1e12, 1)})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)
