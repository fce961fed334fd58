import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdi_llm_b200 import ops
from mdi_llm_b200.utils.quantize import quantize_fp8_block
ops.require()
M, N, K = 2048, 6144, 4096
a = torch.randn(M, K, device="cuda").bfloat16()
q, s = quantize_fp8_block((torch.randn(N, K, device="cuda") * 0.02).bfloat16())
a8, a_st = ops.quantize_rows_fp8(a)
st = s.t().contiguous()
for _ in range(3):
    out = ops.gemm_fp8(a8, a_st, q.view(torch.uint8), st)
torch.cuda.synchronize()
