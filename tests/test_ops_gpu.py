"""Numerics of every sm_100a decode kernel against a plain PyTorch fp32 reference of the same op."""
import math
import os

import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mdi_llm_b200 import ops

    ops.require()
    return ops


def _ctx(ops, slot=0, pos=0, wait=0, signal=0, token=0):
    c = torch.zeros(ops.CTX_INTS, dtype=torch.int32)
    c[0], c[1], c[2], c[3], c[4] = slot, pos, wait, signal, token
    return c.cuda()


def _rmsnorm_ref(x, w, eps, unit_offset=False):
    xf = x.float()
    xn = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)
    return xn * ((1 + w) if unit_offset else w)


def test_library_loaded_and_device_is_blackwell():
    ops = _ops()
    import ctypes

    sm, maj, mino, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    assert ops.lib().mdi_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mino), ctypes.byref(mem)) == 0
    assert sm.value > 0 and maj.value >= 10, f"expected an sm_100 device, got sm_{maj.value}{mino.value}"


@pytest.mark.parametrize("N,K", [(4096, 4096), (6144, 4096), (4096, 14336), (2048, 5632), (1000, 256), (37, 64)])
@pytest.mark.parametrize("fused_norm", [False, True])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_linear_decode_plain_and_residual(N, K, fused_norm, variant):
    ops = _ops()
    ops.set_linear_variant(variant)
    torch.manual_seed(N + K)
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(K, device="cuda").bfloat16()
    res = torch.randn(N, device="cuda").bfloat16()
    bias = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    nw = (1 + 0.1 * torch.randn(K, device="cuda")).bfloat16()
    y = torch.empty(N, device="cuda", dtype=torch.bfloat16)
    ops.linear_decode(W, x, y, _ctx(ops), bias=bias, residual=res, norm_w=nw if fused_norm else None, eps=1e-5)
    xin = _rmsnorm_ref(x, nw, 1e-5) if fused_norm else x
    ref = (xin.float() @ W.float().T + bias.float()).bfloat16().float() + res.float()
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)
    # fp32 output path (logits): rounded like a bf16 linear would be, stored as fp32
    y32 = torch.empty(N, device="cuda", dtype=torch.float32)
    ops.linear_decode(W, x, y32, _ctx(ops))
    torch.testing.assert_close(y32, x.float() @ W.float().T, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("act,fn", [("silu_gate", torch.nn.functional.silu),
                                    ("gelu_tanh_gate", lambda t: torch.nn.functional.gelu(t, approximate="tanh")),
                                    ("gelu_erf_gate", torch.nn.functional.gelu)])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_linear_decode_gated_mlp(act, fn, variant):
    ops = _ops()
    ops.set_linear_variant(variant)
    torch.manual_seed(0)
    N, K = 14336, 4096
    W1 = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    W2 = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(K, device="cuda").bfloat16()
    nw = (1 + 0.1 * torch.randn(K, device="cuda")).bfloat16()
    y = torch.empty(N, device="cuda", dtype=torch.bfloat16)
    ops.linear_decode(W1, x, y, _ctx(ops), W2=W2, norm_w=nw, eps=1e-5, act=act)
    xin = _rmsnorm_ref(x, nw, 1e-5).float()
    a, b = (xin @ W1.float().T).bfloat16(), (xin @ W2.float().T).bfloat16()
    ref = (fn(a.float()).bfloat16() * b).float()
    torch.testing.assert_close(y.float(), ref, rtol=3e-2, atol=3e-2)


def test_linear_decode_slotted_pointers():
    """x / residual / y addressed as base + slot * stride with the slot read on device."""
    ops = _ops()
    torch.manual_seed(1)
    N = K = 512
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    xs = torch.randn(4, K, device="cuda").bfloat16()
    ys = torch.zeros(4, N, device="cuda", dtype=torch.bfloat16)
    ops.linear_decode(W, xs, ys, _ctx(ops, slot=2), residual=xs, x_slot_stride=K, res_slot_stride=K, y_slot_stride=N)
    ref = (xs[2].float() @ W.float().T).bfloat16().float() + xs[2].float()
    torch.testing.assert_close(ys[2].float(), ref, rtol=2e-2, atol=2e-2)
    assert ys[[0, 1, 3]].abs().sum() == 0


def _interleaved_qkv_ref(W, bias, x, cfg, cos, sin, pos):
    """litGPT split + RoPE in fp32 on bf16-rounded projections (model.py:686-724)."""
    H, G, hs, ne = cfg
    qpk = H // G
    qkv = (x.float() @ W.float().T + (bias.float() if bias is not None else 0)).bfloat16().float()
    qkv = qkv.view(G, qpk + 2, hs)
    q, k, v = qkv[:, :qpk].reshape(H, hs), qkv[:, qpk], qkv[:, qpk + 1]

    def rope(t):
        r = t[..., :ne]
        half = ne // 2
        rot = torch.cat((-r[..., half:], r[..., :half]), -1)
        return torch.cat((r * cos[pos] + rot * sin[pos], t[..., ne:]), -1)

    return rope(q).bfloat16(), rope(k).bfloat16(), v.bfloat16()


@pytest.mark.parametrize("H,G,hs,ne,C", [(32, 8, 128, 128, 4096), (32, 4, 64, 64, 2048), (8, 8, 64, 32, 512), (4, 1, 128, 128, 512)])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_qkv_decode_rope_and_kv_append(H, G, hs, ne, C, variant):
    ops = _ops()
    ops.set_linear_variant(variant)
    from mdi_llm_b200.models.gpt import build_rope_cache

    torch.manual_seed(H * G)
    S, n_slots, slot, pos = 64, 3, 1, 17
    W = (torch.randn((H + 2 * G) * hs, C, device="cuda") * 0.02).bfloat16()
    bias = (torch.randn((H + 2 * G) * hs, device="cuda") * 0.05).bfloat16()
    x = torch.randn(C, device="cuda").bfloat16()
    nw = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    cos, sin = build_rope_cache(S, ne, device=torch.device("cuda"), base=500000)
    q = torch.zeros(H * hs, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(n_slots, 2, G, S, hs, device="cuda", dtype=torch.bfloat16)
    ops.qkv_decode(W, x, cos, sin, q, kv, _ctx(ops, slot=slot, pos=pos), n_head=H, n_groups=G, head_size=hs,
                   rope_n_elem=ne, max_seq=S, bias=bias, norm_w=nw, eps=1e-5)
    qr, kr, vr = _interleaved_qkv_ref(W, bias, _rmsnorm_ref(x, nw, 1e-5), (H, G, hs, ne), cos, sin, pos)
    torch.testing.assert_close(q.view(H, hs).float(), qr.float(), rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(kv[slot, 0, :, pos].float(), kr.float(), rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(kv[slot, 1, :, pos].float(), vr.float(), rtol=3e-2, atol=3e-2)
    untouched = kv.clone()
    untouched[slot, :, :, pos] = 0
    assert untouched.abs().sum() == 0  # nothing else in the pool was written


@pytest.mark.parametrize("H,G,hs", [(32, 8, 128), (32, 4, 64), (8, 8, 64), (8, 1, 128), (4, 2, 128)])
@pytest.mark.parametrize("L", [1, 15, 16, 17, 31, 32, 33, 129, 500, 1024, 1025, 2048])
def test_attn_decode_matches_sdpa(H, G, hs, L):
    ops = _ops()
    torch.manual_seed(L)
    S, n_slots, slot = 2048, 2, 1
    q = torch.randn(H * hs, device="cuda").bfloat16()
    kv = torch.randn(n_slots, 2, G, S, hs, device="cuda").bfloat16()
    y = torch.zeros(H * hs, device="cuda", dtype=torch.bfloat16)
    tickets = torch.zeros(G, device="cuda", dtype=torch.int32)
    # 8 / 40: multiples of the cluster size -> spans launch as clusters of 8 and contexts of 129..1024 positions merge
    # through distributed shared memory; longer ones (and the other split counts) take the ticket merge
    for n_split in (1, 5, 37, 8, 40):
        part = torch.zeros(H * n_split * (hs + 2), device="cuda", dtype=torch.float32)
        y.zero_()
        ops.attn_decode(q, kv, y, part, tickets, _ctx(ops, slot=slot, pos=L - 1), n_head=H, n_groups=G, head_size=hs,
                        max_seq=S, n_split=n_split)
        k = kv[slot, 0, :, :L].float().repeat_interleave(H // G, 0)
        v = kv[slot, 1, :, :L].float().repeat_interleave(H // G, 0)
        ref = torch.softmax((q.view(H, 1, hs).float() @ k.transpose(1, 2)) / math.sqrt(hs), -1) @ v
        torch.testing.assert_close(y.view(H, hs).float(), ref.view(H, hs), rtol=2e-2, atol=2e-2)
        assert tickets.abs().sum() == 0  # self-resetting


def test_embed_and_rmsnorm_rows():
    ops = _ops()
    torch.manual_seed(3)
    V, C = 1000, 512
    wte = torch.randn(V, C, device="cuda").bfloat16()
    x = torch.zeros(2, C, device="cuda", dtype=torch.bfloat16)
    tokens = torch.zeros(2, 16, dtype=torch.int32, device="cuda")
    tokens[1, 5] = 777
    ops.embed(wte, x, _ctx(ops, slot=1, pos=5), tokens=tokens, tok_slot_stride=16, x_slot_stride=C, scale=math.sqrt(C))
    torch.testing.assert_close(x[1].float(), (wte[777].float() * math.sqrt(C)).bfloat16().float())
    ops.embed(wte, x, _ctx(ops, slot=0, pos=0, token=42), x_slot_stride=C)
    assert torch.equal(x[0], wte[42])
    rows = torch.randn(7, C, device="cuda").bfloat16()
    w = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    for unit in (False, True):
        torch.testing.assert_close(ops.rmsnorm_rows(rows, w, 1e-5, unit).float(), _rmsnorm_ref(rows, w, 1e-5, unit).float(),
                                   rtol=2e-2, atol=2e-2)


def test_sample_greedy_and_topk_distribution():
    ops = _ops()
    torch.manual_seed(4)
    V = 128256
    logits = torch.randn(V, device="cuda")
    logits[[5, 77, 4000, 100000]] = torch.tensor([9.0, 9.5, 9.25, 9.5], device="cuda")
    tokens = torch.zeros(2, 64, dtype=torch.int32, device="cuda")
    last = torch.zeros(2, dtype=torch.int32, device="cuda")
    ops.sample(logits, tokens, _ctx(ops, slot=1, pos=3), vocab=V, top_k=None, temperature=0.0, greedy=True, seed=0,
               tok_slot_stride=64, last_token=last)
    assert tokens[1, 3].item() == 77 and last[1].item() == 77  # first arg-max on ties, like torch.argmax
    # top-k = 3 at T = 0.5: only {77, 100000, 4000} can appear, with softmax probabilities
    counts = {}
    n_draws = 3000
    for i in range(n_draws):
        ops.sample(logits, tokens, _ctx(ops, slot=0, pos=i % 64), vocab=V, top_k=3, temperature=0.5, greedy=False,
                   seed=1234 + i * 7919, tok_slot_stride=64)
        t = tokens[0, i % 64].item()
        counts[t] = counts.get(t, 0) + 1
    assert set(counts) <= {77, 100000, 4000}
    p = torch.softmax(torch.tensor([9.5, 9.5, 9.25]) / 0.5, 0)
    for tok, pi in zip((77, 100000, 4000), p.tolist()):
        assert abs(counts.get(tok, 0) / n_draws - pi) < 0.04
    # same (seed, slot, pos) -> same token
    ops.sample(logits, tokens, _ctx(ops, slot=0, pos=1), vocab=V, top_k=200, temperature=0.8, greedy=False, seed=99,
               tok_slot_stride=64)
    a = tokens[0, 1].item()
    ops.sample(logits, tokens, _ctx(ops, slot=0, pos=1), vocab=V, top_k=200, temperature=0.8, greedy=False, seed=99,
               tok_slot_stride=64)
    assert tokens[0, 1].item() == a
    top200 = set(torch.topk(logits, 200).indices.tolist())
    assert a in top200


@pytest.mark.parametrize("greedy,top_k,temp", [(True, None, 0.0), (False, 200, 0.8), (False, 3, 0.5), (False, 1000, 1.3)])
def test_fast_sampler_equals_reference_sampler(greedy, top_k, temp):
    """lm_head-fused statistics + filter/final kernels pick exactly the token the single-CTA
    full-vocabulary reference sampler picks (same key order, same RNG stream)."""
    ops = _ops()
    ops.set_linear_variant(0)
    torch.manual_seed(5)
    V, K = 128256, 512
    W = torch.randn(V, K, device="cuda").bfloat16()
    nw = torch.ones(K, device="cuda", dtype=torch.bfloat16)
    scratch = ops.sample_scratch("cuda")
    logits = torch.zeros(V, device="cuda", dtype=torch.float32)
    tok_a = torch.zeros(1, 32, dtype=torch.int32, device="cuda")
    tok_b = torch.zeros(1, 32, dtype=torch.int32, device="cuda")
    for step in range(12):
        x = torch.randn(K, device="cuda").bfloat16()
        ctx = _ctx(ops, slot=0, pos=step)
        ops.linear_decode(W, x, logits, ctx, norm_w=nw, stats=scratch)
        ops.sample_fast(logits, scratch, tok_a, ctx, vocab=V, top_k=top_k, temperature=temp, greedy=greedy, seed=77,
                        tok_slot_stride=32)
        ops.sample(logits, tok_b, ctx, vocab=V, top_k=top_k, temperature=temp, greedy=greedy, seed=77, tok_slot_stride=32)
        torch.cuda.synchronize()
        assert tok_a[0, step].item() == tok_b[0, step].item(), f"step {step}"
        if greedy:
            assert tok_a[0, step].item() == int(torch.argmax(logits))
    assert scratch[: 4096 + 4].abs().sum() == 0  # histogram, arg-max and candidate counter were reset


@pytest.mark.parametrize("variant", [0, 1])
def test_hop_flag_wait_and_signal_same_device(variant):
    """Producer kernel publishes flag[slot] = ctx.signal after its stores; consumer waits for it."""
    ops = _ops()
    ops.set_linear_variant(variant)
    N = K = 256
    W = torch.eye(N, device="cuda").bfloat16()
    x = torch.randn(K, device="cuda").bfloat16()
    inbox = torch.zeros(4, N, device="cuda", dtype=torch.bfloat16)
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    ctx = _ctx(ops, slot=3, wait=7, signal=7)
    ops.linear_decode(W, x, None, ctx, y_ptr=inbox.data_ptr(), y_slot_stride=N, signal_flag=flags.data_ptr(),
                      done_ctr=done.data_ptr())
    y = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    ops.linear_decode(W, inbox, y, ctx, x_slot_stride=N, wait_flag=flags.data_ptr(), status=status.data_ptr(),
                      wait_max_cycles=10 ** 9)
    torch.cuda.synchronize()
    assert flags.tolist() == [0, 0, 0, 7] and done.item() == 0 and status[:2].tolist() == [0, 0]
    assert torch.equal(y, x)
    # watchdog: waiting for a value that never comes sets the status word instead of hanging
    ctx2 = _ctx(ops, slot=0, wait=1)
    ops.linear_decode(W, inbox, y, ctx2, x_slot_stride=N, wait_flag=flags.data_ptr(), status=status.data_ptr(),
                      wait_max_cycles=2_000_000)
    torch.cuda.synchronize()
    assert status[:2].tolist() == [1, 1]  # watchdog bit + aborted


def test_poison_flag_aborts_and_propagates():
    """Abort protocol (common.cuh): a poisoned incoming flag satisfies the wait at once and marks the stage
    aborted; an aborted stage never waits again and publishes the poison instead of its round number."""
    ops = _ops()
    N = K = 256
    W = torch.eye(N, device="cuda").bfloat16()
    inbox = torch.randn(4, N, device="cuda").bfloat16()
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    nxt_flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    nxt_in = torch.zeros(4, N, device="cuda", dtype=torch.bfloat16)
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    flags[2] = ops.POISON
    ctx = _ctx(ops, slot=2, wait=5, signal=5)
    t0 = time.time()
    ops.linear_decode(W, inbox, None, ctx, x_slot_stride=N, wait_flag=flags.data_ptr(), status=status.data_ptr(),
                      wait_max_cycles=10 ** 11, y_ptr=nxt_in.data_ptr(), y_slot_stride=N, signal_flag=nxt_flags.data_ptr(),
                      done_ctr=done.data_ptr())
    # now aborted: a wait that could never be satisfied (slot 0, flag 0 < 9) returns immediately, no watchdog bit
    ops.linear_decode(W, inbox, None, _ctx(ops, slot=0, wait=9, signal=9), x_slot_stride=N, wait_flag=flags.data_ptr(),
                      status=status.data_ptr(), wait_max_cycles=10 ** 11, y_ptr=nxt_in.data_ptr(), y_slot_stride=N,
                      signal_flag=nxt_flags.data_ptr(), done_ctr=done.data_ptr())
    torch.cuda.synchronize()
    assert time.time() - t0 < 5.0
    assert status[:2].tolist() == [0, 1]
    assert nxt_flags.tolist() == [ops.POISON, 0, ops.POISON, 0]


def _ref_support_probs(logits, top_k, top_p, temperature):
    from mdi_llm_b200.models.gpt import sample_top_p

    row = logits.float().clone()
    if top_k is not None and top_k < row.numel():
        vals, idx = torch.topk(row, top_k)
        row = torch.full_like(row, float("-inf")).scatter_(-1, idx, vals)
    row = row / temperature
    if top_p < 1.0:
        row = sample_top_p(row, top_p)
    return torch.softmax(row, -1)


@pytest.mark.parametrize("top_k,top_p", [(None, 0.6), (3000, 1.0), (50, 0.9), (None, 1.0)])
def test_device_sampler_top_p_and_large_k(top_k, top_p):
    """Whole-vocabulary sampler (nucleus by probability-mass radix selection, any k): the support and the
    frequencies of many draws follow the eager `sample()` distribution (reference model.py:42-90)."""
    ops = _ops()
    torch.manual_seed(11)
    V = 6000
    logits = (torch.randn(V, device="cuda") * 2.0).float()
    logits[[5, 77, 4000, 5999]] += torch.tensor([9.0, 8.5, 8.0, 7.0], device="cuda")
    temp = 0.9
    probs = _ref_support_probs(logits, top_k, top_p, temp).cpu()
    n = 600
    tokens = torch.zeros(1, n + 1, dtype=torch.int32, device="cuda")
    scratch = ops.sample_scratch("cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    half = n // 2
    for pos in range(half):  # stand-alone kernel
        ops.sample(logits, tokens, _ctx(ops, slot=0, pos=pos), vocab=V, top_k=top_k, temperature=temp, greedy=False, seed=5,
                   tok_slot_stride=n + 1, top_p=top_p)
    W = torch.eye(8, device="cuda").bfloat16()  # statistics as the lm_head epilogue would leave them are not needed for this path
    for pos in range(half, n):  # fast-sampler entry point routes to the same whole-vocabulary code
        ops.sample_fast(logits, scratch, tokens, _ctx(ops, slot=0, pos=pos), vocab=V, top_k=top_k, temperature=temp, greedy=False,
                        seed=5, tok_slot_stride=n + 1, top_p=top_p, status=status)
    torch.cuda.synchronize()
    draws = tokens[0, :n].cpu().long()
    support = set(torch.nonzero(probs > 0).view(-1).tolist())
    assert set(draws.tolist()) <= support, "sampled a token outside the top-k / top-p support"
    freq = torch.bincount(draws, minlength=V).float() / n
    top = torch.topk(probs, 4).indices
    assert (freq[top] - probs[top]).abs().max().item() < 0.07, (freq[top], probs[top])
    # the two entry points draw from the same stream: same (seed, slot, pos) -> same token
    ops.sample(logits, tokens, _ctx(ops, slot=0, pos=half), vocab=V, top_k=top_k, temperature=temp, greedy=False, seed=5,
               tok_slot_stride=n + 1, top_p=top_p)
    torch.cuda.synchronize()
    assert tokens[0, half].item() == int(draws[half])


def test_fast_sampler_candidate_overflow_is_exact_not_silent():
    """A flat logit row puts the whole vocabulary into the threshold bin: the candidate list overflows, status bit 4
    is raised and the draw still comes from the exact distribution (all ties kept -> uniform over the row)."""
    ops = _ops()
    ops.set_linear_variant(0)
    V, K = 20000, 256
    W = torch.zeros(V, K, device="cuda").bfloat16()
    nw = torch.ones(K, device="cuda", dtype=torch.bfloat16)
    scratch = ops.sample_scratch("cuda")
    logits = torch.zeros(V, device="cuda", dtype=torch.float32)
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    tok = torch.zeros(1, 64, dtype=torch.int32, device="cuda")
    x = torch.randn(K, device="cuda").bfloat16()
    for step in range(40):
        ctx = _ctx(ops, slot=0, pos=step)
        ops.linear_decode(W, x, logits, ctx, norm_w=nw, stats=scratch)
        ops.sample_fast(logits, scratch, tok, ctx, vocab=V, top_k=200, temperature=0.8, greedy=False, seed=3, tok_slot_stride=64,
                        status=status)
    torch.cuda.synchronize()
    assert status[0].item() & 4
    got = tok[0, :40].tolist()
    assert all(0 <= t < V for t in got) and len(set(got)) > 30  # spread over the row, not stuck on one id
    assert scratch[: 4096 + 4].abs().sum() == 0


def test_cuda_graph_capture_and_replay():
    ops = _ops()
    ops.set_linear_variant(0)
    N = K = 1024
    W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    x = torch.randn(K, device="cuda").bfloat16()
    y = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    ctx = _ctx(ops)
    ops.linear_decode(W, x, y, ctx, use_pdl=True)
    torch.cuda.synchronize()
    g = ops.CudaGraph()
    with g:
        ops.linear_decode(W, x, y, ctx, use_pdl=True)
        ops.linear_decode(W, y, x, ctx, use_pdl=True)
    assert g.n_nodes == 2
    x0 = x.clone()
    g.launch(1)
    torch.cuda.synchronize()
    y_ref = (x0.float() @ W.float().T).bfloat16()
    torch.testing.assert_close(y.float(), y_ref.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(x.float(), (y_ref.float() @ W.float().T).bfloat16().float(), rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("N,K", [(4096, 4096), (6144, 4096), (4096, 14336), (1000, 256)])
def test_linear_decode_fp8_block_scaled(N, K):
    """W8A16: fp8-e4m3 weights with per-128 block scales vs the dequantised fp32 reference."""
    ops = _ops()
    from mdi_llm_b200.utils.quantize import dequantize_fp8_block, quantize_fp8_block

    torch.manual_seed(N * 3 + K)
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    W2 = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(K, device="cuda").bfloat16()
    nw = (1 + 0.1 * torch.randn(K, device="cuda")).bfloat16()
    res = torch.randn(N, device="cuda").bfloat16()
    (q, s), (q2, s2) = quantize_fp8_block(W), quantize_fp8_block(W2)
    Wd, W2d = dequantize_fp8_block(q, s, torch.float32), dequantize_fp8_block(q2, s2, torch.float32)
    xin = _rmsnorm_ref(x, nw, 1e-5).float()
    y = torch.empty(N, device="cuda", dtype=torch.bfloat16)
    ops.linear_decode(q, x, y, _ctx(ops), wscale=s, norm_w=nw, residual=res, eps=1e-5)
    ref = (xin @ Wd.T).bfloat16().float() + res.float()
    torch.testing.assert_close(y.float(), ref, rtol=3e-2, atol=3e-2)
    ops.linear_decode(q, x, y, _ctx(ops), wscale=s, W2=q2, wscale2=s2, norm_w=nw, act="silu_gate", eps=1e-5)
    a, b = (xin @ Wd.T).bfloat16(), (xin @ W2d.T).bfloat16()
    ref = (torch.nn.functional.silu(a.float()).bfloat16() * b).float()
    torch.testing.assert_close(y.float(), ref, rtol=4e-2, atol=4e-2)


def test_qkv_decode_fp8_block_scaled():
    ops = _ops()
    from mdi_llm_b200.models.gpt import build_rope_cache
    from mdi_llm_b200.utils.quantize import dequantize_fp8_block, quantize_fp8_block

    torch.manual_seed(11)
    H, G, hs, ne, C, S, pos = 32, 8, 128, 128, 4096, 64, 9
    W = (torch.randn((H + 2 * G) * hs, C, device="cuda") * 0.02).bfloat16()
    x = torch.randn(C, device="cuda").bfloat16()
    nw = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    q8, s8 = quantize_fp8_block(W)
    cos, sin = build_rope_cache(S, ne, device=torch.device("cuda"), base=500000)
    q = torch.zeros(H * hs, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(1, 2, G, S, hs, device="cuda", dtype=torch.bfloat16)
    ops.qkv_decode(q8, x, cos, sin, q, kv, _ctx(ops, pos=pos), n_head=H, n_groups=G, head_size=hs, rope_n_elem=ne,
                   max_seq=S, norm_w=nw, eps=1e-5, wscale=s8)
    Wd = dequantize_fp8_block(q8, s8, torch.bfloat16)
    qr, kr, vr = _interleaved_qkv_ref(Wd, None, _rmsnorm_ref(x, nw, 1e-5), (H, G, hs, ne), cos, sin, pos)
    torch.testing.assert_close(q.view(H, hs).float(), qr.float(), rtol=4e-2, atol=4e-2)
    torch.testing.assert_close(kv[0, 0, :, pos].float(), kr.float(), rtol=4e-2, atol=4e-2)
    torch.testing.assert_close(kv[0, 1, :, pos].float(), vr.float(), rtol=4e-2, atol=4e-2)


@pytest.mark.parametrize("H,G,hs,ne", [(32, 8, 128, 128), (8, 2, 64, 64), (4, 4, 128, 64), (8, 1, 64, 32)])
@pytest.mark.parametrize("T", [5, 64, 128, 129, 300, 512])
@pytest.mark.parametrize("pipe", [0, 1, 2])
def test_attn_prefill_tcgen05_matches_eager(H, G, hs, ne, T, pipe):
    """RoPE + KV append + causal flash attention (tcgen05, S / P.V in TMEM) vs the eager attend_qkv."""
    from mdi_llm_b200.models.config import Config
    from mdi_llm_b200.models.gpt import CausalSelfAttention, build_rope_cache

    ops = _ops()
    default_pipe = int(ops.lib().mdi_get_prefill_attn_pipe())
    ops.set_prefill_attn_pipe(pipe)
    torch.manual_seed(T + H)
    cfg = Config.from_name("tiny-llama-1.1b", n_layer=1, n_embd=H * hs, n_head=H, n_query_groups=G,
                           rotary_percentage=ne / hs, block_size=512)
    assert cfg.head_size == hs and cfg.rope_n_elem == ne
    S, n_slots, slot = 512, 3, 1
    qkv = (torch.randn(T, (H + 2 * G) * hs, device="cuda") * 0.7).bfloat16()
    cos, sin = build_rope_cache(S, ne, device=torch.device("cuda"))
    cos, sin = cos.float().contiguous(), sin.float().contiguous()
    pool = (torch.randn(n_slots, 2, G, S, hs, device="cuda") * 0.3).bfloat16()
    pool_ref = pool.clone()
    try:
        y = ops.attn_prefill(qkv, cos, sin, pool, slot, n_head=H, n_groups=G, head_size=hs, rope_n_elem=ne)
        torch.cuda.synchronize()
    finally:
        ops.set_prefill_attn_pipe(default_pipe)
    attn = CausalSelfAttention(cfg).cuda().bfloat16()
    pos = torch.arange(T, device="cuda")
    with torch.no_grad():
        ref = attn.attend_qkv(qkv.unsqueeze(0), cos[:T], sin[:T], pos, (pool_ref[slot, 0], pool_ref[slot, 1]))[0]
    torch.testing.assert_close(y.float(), ref.float(), rtol=3e-2, atol=3e-2)
    # K and V of the prompt landed in the slot; nothing else in the pool moved
    torch.testing.assert_close(pool[slot, :, :, :T].float(), pool_ref[slot, :, :, :T].float(), rtol=2e-2, atol=2e-2)
    other = [s for s in range(n_slots) if s != slot]
    assert torch.equal(pool[other], pool_ref[other]) and torch.equal(pool[slot, :, :, T:], pool_ref[slot, :, :, T:])


@pytest.mark.parametrize("E,top,K", [(8, 2, 4096), (4, 1, 256), (64, 8, 1024)])
def test_moe_router_picks_the_top_experts(E, top, K):
    """Router kernel vs torch: bf16 logits of the normalised row, top-k, softmax over the chosen logits (model.py:835-838)."""
    ops = _ops()
    torch.manual_seed(E * 7 + K)
    Wg = (torch.randn(E, K, device="cuda") * 0.05).bfloat16()
    x = torch.randn(3, K, device="cuda").bfloat16()  # slotted input: slot 2 is the live one
    nw = (1 + 0.1 * torch.randn(K, device="cuda")).bfloat16()
    sel = torch.full((8,), -1, dtype=torch.int32, device="cuda")
    wts = torch.zeros(8, dtype=torch.float32, device="cuda")
    ops.moe_router(Wg, x, sel, wts, _ctx(ops, slot=2), top=top, norm_w=nw, eps=1e-5, x_slot_stride=K)
    torch.cuda.synchronize()
    logits = (_rmsnorm_ref(x[2], nw, 1e-5).float() @ Wg.float().T).bfloat16()
    w_ref, c_ref = torch.topk(logits, top)
    w_ref = w_ref.softmax(dim=-1, dtype=torch.float).bfloat16().float()
    # a different summation order may flip two experts whose bf16 logits are (nearly) tied: compare through the logits
    got = sel[:top].long()
    assert len(set(got.tolist())) == top and int(got.min()) >= 0 and int(got.max()) < E
    torch.testing.assert_close(logits[got].float(), logits[c_ref].float(), rtol=0, atol=2 ** -6 * float(logits.abs().max()))
    if torch.equal(got, c_ref):
        torch.testing.assert_close(wts[:top], w_ref, rtol=2e-2, atol=4e-3)
    assert abs(float(wts[:top].sum()) - 1.0) < 2e-2 and int(sel[top:].max() if top < 8 else -1) == -1


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("N,I", [(4096, 14336), (256, 128), (1000, 384)])
def test_moe_expert_passes_follow_the_router_output(N, I, variant):
    """Two routed experts through the pointer tables: h = silu(W1[e] xn) * W2[e] xn, y = x + sum_k w_k * (W3[e_k] h_k), with
    the eager module's rounding points; the pass of the last expert adds the residual and hops by row copy."""
    ops = _ops()
    ops.set_moe_variant(variant)  # 0: register-streamed, 2: bulk-copy ring
    torch.manual_seed(N + I)
    E, C = 4, N
    W1 = [(torch.randn(I, C, device="cuda") * 0.03).bfloat16() for _ in range(E)]
    W2 = [(torch.randn(I, C, device="cuda") * 0.03).bfloat16() for _ in range(E)]
    W3 = [(torch.randn(C, I, device="cuda") * 0.03).bfloat16() for _ in range(E)]
    tab = lambda ws: torch.tensor([w.data_ptr() for w in ws], dtype=torch.int64, device="cuda")  # noqa: E731
    p1, p2, p3 = tab(W1), tab(W2), tab(W3)
    x = torch.randn(2, C, device="cuda").bfloat16()
    nw = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    sel = torch.tensor([3, 1, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device="cuda")
    wts = torch.tensor([0.625, 0.375, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device="cuda")
    ctx = _ctx(ops, slot=1, signal=7)
    h = torch.empty(I, device="cuda", dtype=torch.bfloat16)
    acc = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    out_local = torch.zeros(2, C, device="cuda", dtype=torch.bfloat16)
    peer = torch.zeros(2, C, device="cuda", dtype=torch.bfloat16)
    flag = torch.zeros(2, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    xn = _rmsnorm_ref(x[1], nw, 1e-5)
    ref_sum = None
    for k, (e, w) in enumerate([(3, 0.625), (1, 0.375)]):
        ops.moe_linear_decode(p1, x, h, ctx, sel, wts, k, N=I, K=C, w2_ptrs=p2, norm_w=nw, eps=1e-5, act="silu_gate", x_slot_stride=C,
                              sel_early=k > 0)
        g = (xn.float() @ W1[e].float().T).bfloat16()
        u = (xn.float() @ W2[e].float().T).bfloat16()
        h_ref = torch.nn.functional.silu(g.float()).bfloat16() * u
        torch.testing.assert_close(h.float(), h_ref.float(), rtol=3e-2, atol=3e-2)
        y_e = ((h.float() @ W3[e].float().T).bfloat16().float() * w).bfloat16()
        if k == 0:
            ops.moe_linear_decode(p3, h, acc, ctx, sel, wts, k, N=C, K=I, sel_early=True)
            torch.cuda.synchronize()
            torch.testing.assert_close(acc.float(), y_e.float(), rtol=2e-2, atol=2e-2)
            ref_sum = y_e
        else:
            ops.moe_linear_decode(p3, h, None, ctx, sel, wts, k, N=C, K=I, prev=acc, residual=x, res_slot_stride=C,
                                  y_ptr=out_local.data_ptr(), y_slot_stride=C, hop_ptr=peer.data_ptr(), hop_slot_stride=C,
                                  signal_flag=flag.data_ptr(), done_ctr=done.data_ptr(), status=status.data_ptr(), sel_early=True)
            ref_sum = (ref_sum.float() + y_e.float()).bfloat16()
    torch.cuda.synchronize()
    ref = x[1].float() + ref_sum.float()
    torch.testing.assert_close(out_local[1].float(), ref, rtol=2e-2, atol=3e-2)
    assert torch.equal(peer[1], out_local[1]) and int(flag[1]) == 7 and int(flag[0]) == 0 and int(peer[0].abs().max()) == 0
    assert int(status[0]) == 0
    ops.set_moe_variant(0)
