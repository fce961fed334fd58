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
