"""tcgen05/TMEM/TMA GEMM (prefill linears) against an fp32 PyTorch reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mdi_llm_b200 import ops

    ops.require()
    return ops


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 256), (64, 6144, 4096), (200, 4096, 4096), (1000, 14336, 4096),
                                   (37, 4096, 14336), (512, 1000, 520)])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_matches_fp32_reference(M, N, K, block_n):
    ops = _ops()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = ops.gemm(a, w, block_n=block_n)
    ref = a.float() @ w.float().T
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.02 * ref.abs().max().item() + 0.05, f"max err {err}"


def test_gemm_bias_and_residual_epilogue():
    ops = _ops()
    torch.manual_seed(0)
    M, N, K = 300, 2048, 1024
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = ops.gemm(a, w, bias=bias, residual=res)
    ref = (a.float() @ w.float().T + bias.float()).bfloat16().float() + res.float()
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=6e-2)


def test_gemm_throughput_smoke():
    """Not a benchmark (profiles/ has those) — just checks the kernel is in the tensor-core regime."""
    ops = _ops()
    M, N, K = 2048, 14336, 4096
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    out = ops.gemm(a, w, block_n=256)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, w, out=out, block_n=256)
    e1.record()
    torch.cuda.synchronize()
    tflops = 10 * 2 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(f"tcgen05 gemm {M}x{N}x{K}: {tflops:.0f} TFLOP/s")
    assert tflops > 100  # CUDA-core bf16 FMA tops out far below this


@pytest.mark.parametrize("M,N,K", [(64, 14336, 4096), (300, 5632, 2048), (129, 1000, 520)])
@pytest.mark.parametrize("act", ["silu_gate", "gelu_tanh_gate"])
@pytest.mark.parametrize("block_n", [64, 128])
def test_gated_gemm_dual_accumulator(M, N, K, act, block_n):
    """act(a w1^T + b1) * (a w2^T + b2) from two TMEM accumulators in one pass."""
    ops = _ops()
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w1 = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    w2 = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    b1, b2 = torch.randn(N, device="cuda").bfloat16(), torch.randn(N, device="cuda").bfloat16()
    out = ops.gemm(a, w1, bias=b1, w2=w2, bias2=b2, act=act, block_n=block_n)
    g = (a.float() @ w1.float().T + b1.float()).bfloat16().float()
    u = (a.float() @ w2.float().T + b2.float()).bfloat16().float()
    gate = torch.nn.functional.silu(g) if act == "silu_gate" else torch.nn.functional.gelu(g, approximate="tanh")
    ref = gate.bfloat16().float() * u
    torch.testing.assert_close(out.float(), ref, rtol=3e-2, atol=8e-2)


def test_gemm_fused_hop_signal_local():
    """Epilogue stores into a raw destination pointer and the last CTA publishes flag[slot] = signal."""
    ops = _ops()
    torch.manual_seed(1)
    M, N, K = 200, 4096, 1024
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    dst = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    ctx = torch.zeros(ops.CTX_INTS, dtype=torch.int32, device="cuda")
    ctx[0], ctx[3] = 2, 7  # slot 2, signal value 7
    assert ops.gemm(a, w, residual=res, out_ptr=dst.data_ptr(), signal_flag=flags.data_ptr(), done_ctr=ctr, ctx=ctx) is None
    torch.cuda.synchronize()
    assert flags.tolist() == [0, 0, 7, 0] and int(ctr) == 0
    ref = (a.float() @ w.float().T).bfloat16().float() + res.float()
    torch.testing.assert_close(dst.float(), ref, rtol=2e-2, atol=6e-2)
