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


def _deq_rows(q8, scale_t, M):
    """e4m3 bytes + transposed block scales -> fp32."""
    K = q8.shape[1]
    return (q8.view(torch.float8_e4m3fn).float().view(M, K // 128, 128) * scale_t[:, :M].t().unsqueeze(-1)).view(M, K)


def test_quantize_rows_fp8_matches_torch():
    from mdi_llm_b200 import ops

    ops.require()
    torch.manual_seed(0)
    x = (torch.randn(70, 512, device="cuda") * 3).bfloat16()
    x[3] = 0  # an all-zero row keeps scale 1 and zeros
    q, st = ops.quantize_rows_fp8(x)
    assert q.shape == (70, 512) and st.shape == (4, 128)
    xf = x.float().view(70, 4, 128)
    scale = xf.abs().amax(-1) / 448.0
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    torch.testing.assert_close(st[:, :70].t(), scale, rtol=1e-6, atol=0)
    ref = (xf / scale.unsqueeze(-1)).clamp(-448, 448).to(torch.float8_e4m3fn).view(70, 512)
    assert (q.view(torch.float8_e4m3fn).float() - ref.float()).abs().max().item() <= 32.0 * 0.07  # <= 1 ulp at the top binade
    assert (q.view(torch.float8_e4m3fn).float() == ref.float()).float().mean().item() > 0.99


@pytest.mark.parametrize("M,N,K,bias,res", [(128, 128, 128, False, False), (200, 384, 512, True, True), (64, 4096, 4096, False, True),
                                             (300, 1000, 1792, True, False), (1024, 6144, 4096, False, False)])
def test_gemm_fp8_blockscaled_matches_dequantised_reference(M, N, K, bias, res):
    """tcgen05.mma kind::f8f6f4 with per-128-K block scales on both operands against the same product computed in
    fp32 from the dequantised operands (identical rounding of the inputs, only the accumulation order differs)."""
    from mdi_llm_b200 import ops
    from mdi_llm_b200.utils.quantize import dequantize_fp8_block, quantize_fp8_block

    ops.require()
    torch.manual_seed(1)
    a = (torch.randn(M, K, device="cuda") * 0.7).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    w[:, 128:256] *= 8.0 if K > 128 else 1.0  # block scales that really differ along K
    b = (torch.randn(N, device="cuda") * 0.3).bfloat16() if bias else None
    r = (torch.randn(M, N, device="cuda") * 0.5).bfloat16() if res else None
    q, s = quantize_fp8_block(w)
    a8, a_st = ops.quantize_rows_fp8(a)
    out = ops.gemm_fp8(a8, a_st, q.view(torch.uint8), s.t().contiguous(), bias=b, residual=r)
    ref = _deq_rows(a8, a_st, M) @ dequantize_fp8_block(q, s, torch.float32).t()
    if b is not None:
        ref = ref + b.float()
    ref = ref.bfloat16().float()
    if r is not None:
        ref = ref + r.float()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 0.01 * scale + 0.02, f"max err {err} (scale {scale})"
    # and it is a faithful fp8 GEMM of the ORIGINAL operands (quantisation error only)
    full = a.float() @ w.float().t() + (b.float() if b is not None else 0) + (r.float() if r is not None else 0)
    assert (out.float() - full).norm().item() <= 0.06 * full.norm().item()


@pytest.mark.parametrize("M,I,K", [(128, 64, 128), (130, 448, 512), (512, 3584, 4096)])
def test_gemm_fp8_gated_mlp_one_pass(M, I, K):
    """SwiGLU in one fp8 GEMM: fc_1 and fc_2 rows share a B tile, act(g) * u in the epilogue."""
    from mdi_llm_b200 import ops
    from mdi_llm_b200.utils.quantize import dequantize_fp8_block, quantize_fp8_block

    ops.require()
    torch.manual_seed(2)
    a = (torch.randn(M, K, device="cuda") * 0.6).bfloat16()
    w1 = (torch.randn(I, K, device="cuda") * 0.05).bfloat16()
    w2 = (torch.randn(I, K, device="cuda") * 0.05).bfloat16()
    (q1, s1), (q2, s2) = quantize_fp8_block(w1), quantize_fp8_block(w2)
    a8, a_st = ops.quantize_rows_fp8(a)
    out = ops.gemm_fp8(a8, a_st, q1.view(torch.uint8), s1.t().contiguous(), w2_8=q2.view(torch.uint8), w2_scale_t=s2.t().contiguous(),
                       act="silu_gate")
    ad = _deq_rows(a8, a_st, M)
    g = (ad @ dequantize_fp8_block(q1, s1, torch.float32).t()).bfloat16().float()
    u = (ad @ dequantize_fp8_block(q2, s2, torch.float32).t()).bfloat16().float()
    ref = torch.nn.functional.silu(g).bfloat16().float() * u
    err = (out.float() - ref).abs().max().item()
    assert out.shape == (M, I) and err <= 0.02 * ref.abs().max().item() + 0.02, err


@pytest.mark.parametrize("M,N,K,bias,res", [(256, 256, 64, False, False), (256, 256, 512, False, False), (512, 768, 1024, True, True),
                                             (2048, 6144, 4096, False, False), (1000, 4096, 14336, False, True), (300, 1000, 520, True, False)])
def test_gemm_cta_pair_kernel_matches_reference(M, N, K, bias, res):
    """cta_group::2: two CTAs of a cluster share one 256 x 256 tile (each loads its A rows and half of B, the
    leader issues M = 256 MMAs, commits are multicast to the pair)."""
    import os

    if os.environ.get("MDI_TEST_PAIR", "1") == "0":
        pytest.skip("CTA-pair GEMM disabled")
    ops = _ops()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16() if bias else None
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    out = ops.gemm(a, w, bias=b, residual=r, block_n=512)
    ref = a.float() @ w.float().T + (b.float() if bias else 0)
    ref = ref.bfloat16().float() + (r.float() if res else 0)
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.02 * ref.abs().max().item() + 0.06, f"max err {err}"


@pytest.mark.parametrize("M,I,K", [(256, 128, 64), (300, 1000, 520), (2048, 14336, 4096)])
def test_gemm_cta_pair_gated_matches_reference(M, I, K):
    """Gated MLP on the CTA-pair kernel: rank 0 loads the fc_1 half of the B tile, rank 1 the fc_2 half."""
    ops = _ops()
    torch.manual_seed(M + I + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w1 = (torch.randn(I, K, device="cuda") * 0.05).bfloat16()
    w2 = (torch.randn(I, K, device="cuda") * 0.05).bfloat16()
    out = ops.gemm(a, w1, w2=w2, act="silu_gate", block_n=512)
    g = (a.float() @ w1.float().T).bfloat16().float()
    u = (a.float() @ w2.float().T).bfloat16().float()
    ref = torch.nn.functional.silu(g).bfloat16().float() * u
    err = (out.float() - ref).abs().max().item()
    assert out.shape == (M, I) and err <= 0.02 * ref.abs().max().item() + 0.05, err
