#!/usr/bin/env python3
"""Micro-benchmark of the decode kernels at Llama-3-8B shapes: achieved HBM bandwidth per kernel.

Each launch streams a *different* weight copy (ring of buffers > L2) so the data is cold; timing
is CUDA events over a batch of launches after warm-up.  Prints one JSON line per case and a
summary; `--out` writes them to a file for profiles/.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops  # noqa: E402


def time_launches(fn, n_iter, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_iter):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_iter * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--ctas", default="3")
    ap.add_argument("--peak", type=float, default=6577.0, help="measured copy bandwidth GB/s (MEASURED_PEAKS.json)")
    args = ap.parse_args()
    ops.require()
    dev = "cuda"
    ctx = torch.zeros(ops.CTX_INTS, dtype=torch.int32, device=dev)
    C, I, QKV, V = 4096, 14336, 6144, 128256
    ring = 6
    results = []

    def case(name, N, K, gated=False, norm=False, residual=False, out_fp32=False, pdl=False):
        n_w = 2 if gated else 1
        Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(ring * n_w)]
        x = torch.randn(K, device=dev, dtype=torch.bfloat16)
        nw = torch.ones(K, device=dev, dtype=torch.bfloat16) if norm else None
        res = torch.randn(N, device=dev, dtype=torch.bfloat16) if residual else None
        y = torch.empty(N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
        nbytes = n_w * N * K * 2
        combos = [(0, int(c)) for c in args.ctas.split(",")] + [(1, 1), (2, 2), (2, 3), (3, 2)]
        for variant, cps in combos:
            def fn(i):
                j = i % ring
                ops.linear_decode(Ws[j * n_w], x, y, ctx, W2=Ws[j * n_w + 1] if gated else None, norm_w=nw, residual=res,
                                  act="silu_gate" if gated else "none", ctas_per_sm=cps, use_pdl=pdl, variant=variant)
            us = time_launches(fn, args.iters)
            gbs = nbytes / us / 1e3
            r = {"kernel": name, "N": N, "K": K, "variant": ["ldg", "bulk4", "bulk2", "bulk3"][variant], "ctas_per_sm": cps,
                 "us": round(us, 2), "GBps": round(gbs, 1),
                 "frac_of_measured_copy_bw": round(gbs / args.peak, 3), "MB": round(nbytes / 1e6, 1), "pdl": pdl}
            results.append(r)
            print(json.dumps(r), flush=True)
        del Ws
        torch.cuda.empty_cache()

    case("qkv-shaped linear (+norm)", QKV, C, norm=True)
    case("o_proj (+residual)", C, C, residual=True)
    case("gate_up (+norm, silu*mul)", I, C, gated=True, norm=True)
    case("down (+residual)", C, I, residual=True)
    case("lm_head (+norm, fp32 out)", V, C, norm=True, out_fp32=True)
    case("gate_up PDL back-to-back", I, C, gated=True, norm=True, pdl=True)
    # torch reference points: a bf16 matvec through cuBLAS and a plain copy
    W = [torch.randn(I, C, device=dev, dtype=torch.bfloat16) for _ in range(ring)]
    x = torch.randn(C, 1, device=dev, dtype=torch.bfloat16)
    us = time_launches(lambda i: torch.mm(W[i % ring], x), args.iters)
    r = {"kernel": "cuBLAS torch.mm [14336x4096]x[4096x1]", "us": round(us, 2), "GBps": round(I * C * 2 / us / 1e3, 1)}
    results.append(r); print(json.dumps(r), flush=True)
    dst = torch.empty_like(W[0])
    us = time_launches(lambda i: dst.copy_(W[i % ring]), args.iters)
    r = {"kernel": "torch copy 117MB (read+write bytes)", "us": round(us, 2), "GBps": round(2 * I * C * 2 / us / 1e3, 1)}
    results.append(r); print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
