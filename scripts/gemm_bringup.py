#!/usr/bin/env python3
"""Bring-up helper for the tcgen05 GEMM: sweeps the shared-memory descriptor knobs and reports which
encoding reproduces torch.matmul.  Each trial runs in a subprocess with a timeout so a bad
descriptor (hang / illegal instruction) cannot take the sweep down."""
import itertools
import json
import subprocess
import sys

TRIAL = r'''
import sys, torch
sys.path.insert(0, ".")
from mdi_llm_b200 import ops
sbo, lbo, hi, kstep, bn = map(int, sys.argv[1:6])
torch.manual_seed(0)
M, N, K = 256, 512, 256
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
out = ops.gemm(a, w, block_n=bn, _knobs=(sbo, lbo, hi, kstep))
torch.cuda.synchronize()
ref = a.float() @ w.float().T
print("ERR", (out.float() - ref).abs().max().item(), ref.abs().max().item())
'''


def main():
    results = []
    v1, sw128 = 1, 2 << 15
    for sbo, lbo, hi, kstep, bn in itertools.product([64, 1, 128], [1, 64, 0], [v1 | sw128, sw128, v1 | (1 << 15)], [32, 2], [128]):
        try:
            p = subprocess.run([sys.executable, "-c", TRIAL, str(sbo), str(lbo), str(hi), str(kstep), str(bn)],
                               capture_output=True, text=True, timeout=60)
            line = [l for l in p.stdout.splitlines() if l.startswith("ERR")]
            res = line[0] if line else f"rc={p.returncode} {p.stderr[-200:]}"
        except subprocess.TimeoutExpired:
            res = "TIMEOUT"
        results.append({"sbo": sbo, "lbo": lbo, "hi": hi, "kstep": kstep, "bn": bn, "result": res})
        print(json.dumps(results[-1]), flush=True)


if __name__ == "__main__":
    main()
