"""CPU dry run of the fused engine's launch sequences: every ``ops.*`` kernel entry point is replaced
by a recorder that binds the call against the real Python signature, so wrong/duplicate keyword
arguments, missing buffers and mis-ordered hops are caught without a GPU."""
import contextlib
import inspect
import types
from unittest import mock

import pytest
import torch

from mdi_llm_b200 import ops
from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.stage import build_stage
from mdi_llm_b200.parallel import engine as eng
from mdi_llm_b200.parallel.engine import FusedStage, HopTarget

KERNELS = ["linear_decode", "qkv_decode", "attn_decode", "embed", "sample_fast", "advance_step", "rmsnorm_rows", "gemm",
           "attn_prefill", "gemm_fp8", "quantize_rows_fp8", "moe_router", "moe_linear_decode"]


def _cfg(n_layer=3):
    return Config.from_name("tiny-llama-1.1b", n_layer=n_layer, n_embd=256, n_head=4, n_query_groups=2,
                            intermediate_size=128, vocab_size=120, padded_vocab_size=128, block_size=64)


@contextlib.contextmanager
def dry_ops():
    calls = []
    patches = []
    for name in KERNELS:
        real = getattr(ops, name)
        sig = inspect.signature(real)

        def rec(*a, _n=name, _sig=sig, **kw):
            bound = _sig.bind(*a, **kw)  # TypeError on bad / duplicate / unknown arguments
            calls.append((_n, bound.arguments))
            if _n == "gemm":
                if kw.get("out_ptr") is not None:
                    return None
                return torch.zeros(a[0].shape[0], a[1].shape[0], dtype=torch.bfloat16)
            if _n == "rmsnorm_rows":
                return torch.zeros_like(a[0])
            if _n == "quantize_rows_fp8":
                return torch.zeros(a[0].shape, dtype=torch.uint8), torch.zeros(a[0].shape[1] // 128, 128)
            if _n == "gemm_fp8":
                return None if kw.get("out_ptr") is not None else torch.zeros(a[0].shape[0], a[2].shape[0], dtype=torch.bfloat16)
            if _n == "attn_prefill":
                return torch.zeros(a[0].shape[0], kw["n_head"] * kw["head_size"], dtype=torch.bfloat16)
            return None

        patches.append(mock.patch.object(ops, name, rec))
    props = types.SimpleNamespace(multi_processor_count=148)
    patches += [
        mock.patch.object(ops, "require", lambda: None),
        mock.patch.object(ops, "sample_scratch", lambda dev: torch.zeros(16, dtype=torch.int32)),
        mock.patch.object(torch.cuda, "device", lambda d: contextlib.nullcontext()),
        mock.patch.object(torch.cuda, "get_device_properties", lambda d: props),
        mock.patch.object(torch.Tensor, "pin_memory", lambda self: self),
    ]
    with contextlib.ExitStack() as es:
        for p in patches:
            es.enter_context(p)
        yield calls


def _stage(role, n_blocks, **kw):
    st = build_stage(_cfg(), role, n_blocks, **kw).to(torch.bfloat16)
    st.max_seq_length = 32
    return st


@pytest.mark.parametrize("role,kw,expect_first,expect_last", [
    ("starter", {}, "qkv_decode", "linear_decode"),
    ("secondary:0", {}, "qkv_decode", "linear_decode"),
    ("secondary:0", {"first_mlp_only": True}, "linear_decode", "linear_decode"),
    ("secondary:0", {"last_attn_only": True}, "qkv_decode", "linear_decode"),
    ("secondary:0", {"first_mlp_only": True, "last_attn_only": True}, "linear_decode", "linear_decode"),
])
def test_decode_launch_sequence_binds(role, kw, expect_first, expect_last):
    with dry_ops() as calls:
        fs = FusedStage(_stage(role, 3, **kw), n_slots=2, max_seq_length=32)
        hop = HopTarget(0x1000, 0x2000)
        if fs.is_starter:
            fs.enqueue_head(wait=True)
            fs.enqueue_sample()
            fs.enqueue_embed(from_tokens=True)
        n0 = len(calls)
        fs.enqueue_blocks(hop, wait_input=True)
        seq = calls[n0:]
    names = [c[0] for c in seq]
    assert names[0] == expect_first and names[-1] == expect_last
    units = fs._units()
    assert len(names) == sum(3 if k == "attn" else 2 for _, k in units)
    # hop protocol: only the first kernel may wait (secondaries), only the last one signals
    waits = [c[1].get("wait_flag") for c in seq]
    sigs = [c[1].get("signal_flag") for c in seq]
    assert all(w is None for w in waits[1:]) and all(s is None for s in sigs[:-1]) and sigs[-1] == 0x2000
    assert (waits[0] is None) == fs.is_starter
    # hop by row copy: the row is finished in out_local, the last CTA copies it to the next stage's hidden_in
    assert seq[-1][1]["hop_ptr"] == 0x1000 and seq[-1][1]["y_ptr"] == fs.out_local.data_ptr() and seq[-1][1]["done_ctr"] is not None
    # residual stream ping-pong: a kernel never writes the buffer it reads its residual from
    for n, args in seq:
        if n == "linear_decode" and args.get("residual") is not None and args.get("y") is not None:
            assert args["y"] is not args["residual"]


@pytest.mark.parametrize("role,kw", [("starter", {}), ("secondary:0", {"first_mlp_only": True, "last_attn_only": True}),
                                     ("secondary:0", {})])
def test_moe_decode_launch_sequence_binds(role, kw):
    """Mixture of experts: router, then (gate/up, down) per chosen expert; only the router may carry the incoming
    hop wait, the token's last down pass adds the residual and carries the outgoing hop."""
    cfg = Config.from_name("tiny-llama-1.1b", n_layer=3, n_embd=256, n_head=4, n_query_groups=2, intermediate_size=128, vocab_size=120,
                           padded_vocab_size=128, block_size=64, mlp_class_name="LLaMAMoE", n_expert=4, n_expert_per_token=2)
    assert eng.engine_supports(cfg, torch.bfloat16) and eng.fused_prefill_supports(cfg)
    with dry_ops() as calls:
        st = build_stage(cfg, role, 3, **kw).to(torch.bfloat16)
        st.max_seq_length = 32
        fs = FusedStage(st, n_slots=2, max_seq_length=32)
        fs.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True)
    names = [c[0] for c in calls]
    units = fs._units()
    assert len(names) == sum(3 if k == "attn" else 5 for _, k in units)
    first_mlp = units[0][1] == "mlp"
    assert names[0] == ("moe_router" if first_mlp else "qkv_decode")
    assert (calls[0][1].get("wait_flag") is None) == fs.is_starter
    assert all(c[1].get("wait_flag") is None for c in calls[1:])
    sigs = [c[1].get("signal_flag") for c in calls]
    assert all(s_ is None for s_ in sigs[:-1]) and sigs[-1] == 0x2000
    moe = [c[1] for c in calls if c[0] == "moe_linear_decode"]
    per_layer = [moe[i:i + 4] for i in range(0, len(moe), 4)]
    for gu0, dn0, gu1, dn1 in per_layer:
        assert gu0["w2_ptrs"] is not None and gu1["w2_ptrs"] is not None and dn0.get("w2_ptrs") is None
        assert (gu0["k"], dn0["k"], gu1["k"], dn1["k"]) == (0, 0, 1, 1)
        # only the pass right behind the router has to wait for it before it knows its expert
        assert [bool(c.get("sel_early")) for c in (gu0, dn0, gu1, dn1)] == [False, True, True, True]
        assert dn0.get("residual") is None and dn0.get("prev") is None and dn0["y"] is fs.moe_acc[0]
        assert dn1["prev"] is fs.moe_acc[0] and dn1["residual"] is not None  # running sum + residual on the last expert
        assert dn1.get("y") is not dn1["residual"]
        assert gu0["w_ptrs"].dtype == torch.int64 and gu0["w_ptrs"].numel() == 4 and gu0["act"] == "silu_gate"
    if units[-1][1] == "mlp":
        assert moe[-1]["hop_ptr"] == 0x1000 and moe[-1]["y_ptr"] == fs.out_local.data_ptr()
    with dry_ops():
        with pytest.raises(ValueError, match="fp8"):
            FusedStage(build_stage(cfg, "starter", 3).to(torch.bfloat16), n_slots=1, max_seq_length=32, weight_dtype="fp8")
    # the prompt: every routed MLP = per-expert gated GEMM + down GEMM over that expert's rows only
    if role != "starter":
        with dry_ops() as calls:
            st = build_stage(cfg, role, 3, **kw).to(torch.bfloat16)
            st.set_kv_cache(2, dtype=torch.bfloat16)
            fs = FusedStage(st, n_slots=2, max_seq_length=32)
            with mock.patch.object(ops, "check", lambda *a, **k: None), mock.patch.object(ops, "lib", lambda: mock.MagicMock()), \
                    mock.patch.object(ops, "stream_ptr", lambda: 0):
                out = fs.prefill(torch.zeros(1, 5, 256, dtype=torch.bfloat16), torch.arange(5), 0, hop=(0x3000, 0x4000))
            assert out is None
            gated = [c[1] for c in calls if c[0] == "gemm" and c[1].get("w2") is not None]
            n_moe = sum(1 for _, k in fs._units() if k == "mlp")
            # all-zero router logits: topk picks experts 0 and 1 for every token -> two experts x 5 rows per routed MLP
            assert len(gated) == 2 * n_moe and all(g["a"].shape == (5, 256) and g["act"] == "silu_gate" for g in gated)


def test_local_output_and_prefill_sequences_bind():
    with dry_ops() as calls:
        st = _stage("secondary:0", 3, first_mlp_only=True, last_attn_only=True)
        st.set_kv_cache(2, dtype=torch.bfloat16)
        fs = FusedStage(st, n_slots=2, max_seq_length=32)
        fs.enqueue_blocks(None, wait_input=False)
        assert calls[-1][1]["y_ptr"] == fs.out_local.data_ptr() and calls[-1][1].get("signal_flag") is None
        calls.clear()
        x = torch.zeros(1, 5, 256, dtype=torch.bfloat16)
        assert fs.prefill_attn == "tcgen05"
        assert fs.prefill(x, torch.arange(5), 0, hop=(0x3000, 0x4000)) is None
        assert sum(1 for c in calls if c[0] == "attn_prefill") == 2  # blocks 1 ("both") and 2 ("attn")
    gemms = [c[1] for c in calls if c[0] == "gemm"]
    assert gemms[-1]["out_ptr"] == 0x3000 and gemms[-1]["signal_flag"] == 0x4000 and gemms[-1]["residual"] is not None
    assert all(g.get("out_ptr") is None for g in gemms[:-1])
    assert sum(1 for g in gemms if g.get("w2") is not None) == 2  # two MLP halves -> two gated GEMMs


def test_flag_dependency_wiring_is_a_linear_chain():
    """dep_flags=True: kernel j waits on kernel j-1's flag and publishes its own; the first kernel keeps
    the grid-level wait, the last one has no local consumer (it carries the hop)."""
    with dry_ops() as calls:
        fs = FusedStage(_stage("secondary:0", 3, first_mlp_only=True), n_slots=2, max_seq_length=32)
        fs.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True, dep_flags=True)
    seq = [c[1] for c in calls]
    base = fs.dep_flags.data_ptr()
    assert seq[0].get("dep_wait") is None and seq[-1].get("dep_signal") is None
    for j in range(1, len(seq)):
        assert seq[j]["dep_wait"] == base + 4 * (j - 1) == seq[j - 1]["dep_signal"]
        assert seq[j]["dep_ctr"] == fs.dep_ctr.data_ptr()
    assert seq[-1]["signal_flag"] == 0x2000  # the hop is still signalled by the last kernel
    # host-fed step descriptors carry a monotonically increasing step number
    fs.set_ctx(0, 3); fs.set_ctx(1, 3)
    assert fs._step_seq == 2 and int(fs.ctx_ring[1][ops.CTX_STEP]) == 1
    fs.reset_deps()
    assert fs._step_seq == 0


@pytest.mark.parametrize("kw,n_per_block", [
    (dict(norm_class_name="LayerNorm", parallel_residual=True, shared_attention_norm=False, mlp_class_name="GptNeoxMLP", bias=True,
          rotary_percentage=0.25), 5),
    (dict(norm_class_name="LayerNorm", parallel_residual=True, shared_attention_norm=True, mlp_class_name="GptNeoxMLP"), 5),
    (dict(norm_class_name="LayerNorm", parallel_residual=False, mlp_class_name="GptNeoxMLP", gelu_approximate="tanh", bias=True,
          rotary_percentage=0.0, pos_embedding="learned", tie_embeddings=True), 5),
])
def test_decode_sequences_bind_for_neox_falcon_gpt2_shapes(kw, n_per_block):
    """LayerNorm prologues, parallel-residual blocks (x, x + attn and the block output in three distinct buffers),
    plain GELU MLPs and learned positions go through the same fused kernels."""
    cfg = Config.from_name("tiny-llama-1.1b", n_layer=3, n_embd=256, n_head=4, n_query_groups=4, intermediate_size=128, vocab_size=120,
                           padded_vocab_size=128, block_size=64, **kw)
    assert eng.engine_supports(cfg, torch.bfloat16) and not eng.fused_prefill_supports(cfg)
    with dry_ops() as calls:
        st = build_stage(cfg, "starter", 3).to(torch.bfloat16)
        st.max_seq_length = 32
        fs = FusedStage(st, n_slots=2, max_seq_length=32)
        fs.enqueue_head(wait=True)
        fs.enqueue_sample()
        fs.enqueue_embed(from_tokens=True)
        assert (calls[-1][1]["wpe"] is not None) == (cfg.pos_embedding == "learned")
        n0 = len(calls)
        fs.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True)
        seq = calls[n0:]
    assert len(seq) == 3 * n_per_block
    lin = [c[1] for c in seq if c[0] in ("linear_decode", "qkv_decode")]
    assert all(c.get("layer_norm") for c in lin if c.get("norm_w") is not None)
    assert all((c.get("norm_b") is not None) for c in lin if c.get("norm_w") is not None)
    for n, args in seq:  # a kernel never writes the buffer it reads its input or residual from
        if n == "linear_decode" and args.get("y") is not None:
            assert args["y"] is not args.get("residual") and args["y"] is not args["x"]
    if cfg.parallel_residual:
        o_proj, fc, down = seq[2][1], seq[3][1], seq[4][1]
        assert fc["x"] is o_proj["residual"]          # the MLP reads the block INPUT, not x + attn
        assert down["residual"] is o_proj["y"]        # ... and its output is added to x + attn
        assert (fc["norm_w"] is qkv_norm(seq)) == cfg.shared_attention_norm


def qkv_norm(seq):
    return seq[0][1]["norm_w"]


def test_third_unit_stages_bind_and_route_the_wide_message():
    """Stage boundary between a gated MLP's gate/up and down projections: the upstream stage ends with the gate/up
    kernel (row [x | h]: h written behind x in out_local, x copied by the hop's last CTA), the downstream stage
    starts with the down projection reading h and the residual from the two halves of its hidden_in row."""
    C, I = 256, 128
    with dry_ops() as calls:
        up = FusedStage(_stage("secondary:0", 2, last_parts="attn_gu"), n_slots=2, max_seq_length=32)
        assert (up.W_in, up.W_out) == (C, C + I) and up._units() == [(0, "attn"), (0, "mlp"), (1, "attn"), (1, "gu")]
        up.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True)
        last = calls[-1][1]
        assert calls[-1][0] == "linear_decode" and last["W2"] is not None  # the gate/up kernel is the last one
        assert last["y_ptr"] == up.out_local.data_ptr() + 2 * C and last["y_slot_stride"] == C + I
        assert last["hop_ptr"] == 0x1000 and last["hop_slot_stride"] == C + I and last["signal_flag"] == 0x2000
        assert last["hop_pre"][2] == C  # x goes in front
        calls.clear()
        dn = FusedStage(_stage("secondary:1", 2, first_parts="down"), n_slots=2, max_seq_length=32)
        assert (dn.W_in, dn.W_out) == (C + I, C) and dn._units() == [(0, "down"), (1, "attn"), (1, "mlp")]
        dn.enqueue_blocks(HopTarget(0x3000, 0x4000), wait_input=True)
        first = calls[0][1]
        assert calls[0][0] == "linear_decode" and first["wait_flag"] == dn.flags.data_ptr()
        assert first["x_ptr"] == dn.hidden_in.data_ptr() + 2 * C and first["x_slot_stride"] == C + I
        assert first["residual_ptr"] == dn.hidden_in.data_ptr() and first["res_slot_stride"] == C + I
        assert all(c[1].get("wait_flag") is None for c in calls[1:])
        # prefill: [x | h] in, plain x out through the fused-hop GEMM
        calls.clear()
        dn.model.set_kv_cache(2, dtype=torch.bfloat16)
        assert dn.prefill(torch.zeros(1, 5, C + I, dtype=torch.bfloat16), torch.arange(5), 0, hop=(0x5000, 0x6000)) is None
        gemms = [c[1] for c in calls if c[0] == "gemm"]
        assert gemms[0]["residual"] is not None and gemms[0].get("w2") is None and gemms[-1]["out_ptr"] == 0x5000


@pytest.mark.parametrize("kw", [dict(last_parts="attn"), dict(first_parts="down"), dict(first_parts="mlp", last_parts="attn_gu")])
def test_fp8_quantisation_covers_partial_blocks(kw):
    """Sub-layer stages in fp8: a block cut down to its attention has no MLP to quantise, a gate/up-only or a
    down-only block quantises exactly the projections it owns (the 8-GPU fp8 job of bench.py hit this)."""
    import mdi_llm_b200.utils.quantize as Q

    with dry_ops() as calls:
        st = _stage("secondary:0", 2, **kw)
        fake = lambda w: (torch.zeros(w.shape, dtype=torch.uint8), torch.ones(w.shape[0], max(1, w.shape[1] // 128)))  # noqa: E731
        with mock.patch.object(Q, "quantize_fp8_block", fake):
            fs = FusedStage(st, n_slots=1, max_seq_length=32, weight_dtype="fp8")
        n_lin = sum(1 for n, _ in st.named_parameters() if n.endswith(".weight") and (".attn." in n or ".mlp." in n))
        assert len(fs._q) == n_lin == len(fs._qt)
        fs.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True)
        assert all(c[1].get("wscale") is not None for c in calls if c[0] in ("linear_decode", "qkv_decode"))


# ---- DevicePipeline orchestration (fake graphs, no GPU) ---------------------------------------------------
class _FakeGraph:
    instances = []

    def __init__(self):
        self.n_nodes, self.launched = 0, 0
        _FakeGraph.instances.append(self)

    def __enter__(self):
        self._mark = len(_FakeGraph.calls)
        return self

    def __exit__(self, *exc):
        self.kernels = [c[0] for c in _FakeGraph.calls[self._mark:]]
        self.n_nodes = len(self.kernels)

    def launch(self, times=1):
        self.launched += times


@contextlib.contextmanager
def dry_pipeline():
    with dry_ops() as calls:
        _FakeGraph.calls, _FakeGraph.instances = calls, []
        fake_lib = mock.MagicMock()
        fake_lib.mdi_wait_flag.return_value = 0
        fake_lib.mdi_copy_signal.return_value = 0
        stream = types.SimpleNamespace(synchronize=lambda: None)
        with mock.patch.object(ops, "CudaGraph", _FakeGraph), mock.patch.object(ops, "lib", lambda: fake_lib), \
                mock.patch.object(ops, "stream_ptr", lambda: 0), mock.patch.object(ops, "check", lambda code, what="": None), \
                mock.patch.object(torch.cuda, "current_stream", lambda *a: stream), \
                mock.patch.object(torch.cuda, "synchronize", lambda *a: None):
            yield calls, fake_lib


@pytest.mark.parametrize("mode", ["device", "host"])
def test_device_pipeline_orchestration_single_stage(mode):
    """prepare -> prefill -> decode rounds: graph launches, step numbering and the wrap-around prefill hop."""
    from mdi_llm_b200.parallel.pipeline import DevicePipeline
    from mdi_llm_b200.parallel.scheduler import SamplingParams

    with dry_pipeline() as (calls, lib):
        st = _stage("starter", 3)
        pipe = DevicePipeline(st, 0, 1, n_samples=2, max_seq_length=32, sampling=SamplingParams.greedy())
        prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6, 7])]
        out = pipe.generate(prompts, 5, mode=mode)
    assert set(out) == {0, 1} and out[0].shape == (1, 3 + 5) and out[1].shape == (1, 4 + 5)
    assert out[0][0, :3].tolist() == [1, 2, 3]
    # prefill: one wrap-around copy+signal per sample (world == 1: the last stage is also the starter)
    assert lib.mdi_copy_signal.call_count == 2
    # rounds 1..4 run the full step for both samples, the final round only the head (ln_f + lm_head + sample)
    full = [g for g in _FakeGraph.instances if "qkv_decode" in g.kernels]
    head = [g for g in _FakeGraph.instances if "qkv_decode" not in g.kernels and "sample_fast" in g.kernels]
    steps = lambda g: g.kernels.count("embed")  # a device-mode graph holds a whole round (n_samples steps) -> PDL edges across steps
    assert sum(g.launched * steps(g) for g in full) == 4 * 2 and sum(g.launched for g in head) == 2
    # device mode replays the 8-step graph (and the 1-step one for remainders), both captured in prepare(); host-fed steps
    # are one per graph
    launched = [g for g in full if g.launched]
    assert all((steps(g) > 1) == (mode == "device") for g in launched)
    assert {steps(g) for g in full if "advance_step" in g.kernels} == {1, 8}
    assert pipe.n_graph_launches == 10
    dev_ctx = mode == "device"
    assert all(("advance_step" in g.kernels) == dev_ctx for g in full + head if g.launched)  # (device graphs are pre-captured in both modes)
    # the host's step counter stays ahead of the device's in both modes (2 prefill descriptors + 10 steps): the
    # flag dependencies only need it to be strictly increasing from step to step
    assert pipe.stage._step_seq == 12
    with pytest.raises(ValueError, match="exceed block size"):
        pipe.prepare(prompts, 40)


def test_device_pipeline_orchestration_two_stages():
    """rank 0 of 2: prefill hop fused into the last GEMM (no copy kernel); rank 1 (last): waits for its input
    flag per sample and returns only the final row with the copy+signal kernel; secondaries skip the head."""
    from mdi_llm_b200.parallel.pipeline import DevicePipeline
    from mdi_llm_b200.parallel.scheduler import SamplingParams

    prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6, 7])]
    with dry_pipeline() as (calls, lib):
        p0 = DevicePipeline(_stage("starter", 2), 0, 2, n_samples=2, max_seq_length=32, sampling=SamplingParams.greedy(),
                            exportable=False)
        p0.next_hop, p0.next_prefill_ptr = HopTarget(0x10000, 0x20000), 0x30000
        p0.prepare(prompts, 4)
        calls.clear()
        p0.prefill()
        gemms = [c[1] for c in calls if c[0] == "gemm"]
        hops = [g for g in gemms if g.get("out_ptr") is not None]
        assert len(hops) == 2 and lib.mdi_copy_signal.call_count == 0 and lib.mdi_wait_flag.call_count == 0
        assert hops[0]["out_ptr"] == 0x30000 and hops[1]["out_ptr"] == 0x30000 + 32 * 256 * 2  # slot 1: + max_prompt_len * C * 2
        assert all(h["signal_flag"] == 0x20000 for h in hops)
        p0.decode_rounds(3)
        assert p0.n_graph_launches == 6

    with dry_pipeline() as (calls, lib):
        p1 = DevicePipeline(_stage("secondary:0", 1), 1, 2, n_samples=2, max_seq_length=32, sampling=SamplingParams.greedy(),
                            exportable=False)
        p1.next_hop = HopTarget(0x40000, 0x50000)
        p1.prepare(prompts, 4)
        calls.clear()
        p1.prefill()
        assert lib.mdi_wait_flag.call_count == 2 and lib.mdi_copy_signal.call_count == 2
        assert all(c[1].get("out_ptr") is None for c in calls if c[0] == "gemm")
        p1.decode_rounds(4)  # rounds 1..3 forward; the final round has nothing to do on a secondary
        g = [x for x in _FakeGraph.instances if "qkv_decode" in x.kernels]
        assert sum(x.launched * x.kernels.count("advance_step") for x in g) == 6 \
            and not any("sample_fast" in x.kernels for x in _FakeGraph.instances)
        assert g[0].kernels[0] == "advance_step" and g[0].kernels[1] == "qkv_decode"


def test_moe_prompt_grouping_matches_the_eager_module():
    """Host-side logic of the routed prompt path (route, sort by expert, per-expert GEMMs over the expert's rows, weighted
    scatter-add) with the two kernels replaced by their PyTorch definitions: must equal ``x + mlp(norm_2(x))``."""
    cfg = Config.from_name("tiny-llama-1.1b", n_layer=1, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96, vocab_size=120,
                           padded_vocab_size=128, block_size=64, mlp_class_name="LLaMAMoE", n_expert=6, n_expert_per_token=2)
    torch.manual_seed(3)

    def rms(x, w, eps, unit_offset=False):
        xf = x.float()
        return ((xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * w).to(x.dtype)

    def gemm(a, w, *, bias=None, w2=None, bias2=None, act="silu_gate", out=None, **_kw):
        y = torch.nn.functional.linear(a, w, bias)
        if w2 is not None:
            assert act == "silu_gate"
            y = torch.nn.functional.silu(y) * torch.nn.functional.linear(a, w2, bias2)
        if out is not None:
            out.copy_(y)
            return out
        return y

    with dry_ops():
        st = build_stage(cfg, "secondary:0", 1).float()
        for p_ in st.parameters():
            p_.data.normal_(0, 0.2)
        with mock.patch.object(FusedStage, "_check_weights", lambda self: None), \
                mock.patch.object(eng, "engine_supports", lambda c, d: True):
            fs = FusedStage(st, n_slots=1, max_seq_length=32)
        blk = st.transformer.h[0]
        x = torch.randn(37, cfg.n_embd)
        with mock.patch.object(ops, "rmsnorm_rows", rms), mock.patch.object(ops, "gemm", gemm):
            got = fs._moe_prefill(blk, x, cfg.norm_eps, False)
    with torch.no_grad():
        ref = x + blk.mlp(blk.norm_2(x))
    torch.testing.assert_close(got.float(), ref.float(), rtol=1e-4, atol=1e-4)
    # every expert that received tokens ran exactly once over exactly its rows
    assert got.shape == x.shape


def test_device_pipeline_orchestration_with_routed_mlps():
    """A mixture-of-experts stage through prepare -> prefill -> decode: the prompt takes the per-expert GEMM path, the
    step graphs hold router + expert launches, and a 2-stage ring hands the prompt on with the copy + signal kernel."""
    from mdi_llm_b200.parallel.pipeline import DevicePipeline
    from mdi_llm_b200.parallel.scheduler import SamplingParams

    cfg = Config.from_name("tiny-llama-1.1b", n_layer=2, n_embd=256, n_head=4, n_query_groups=2, intermediate_size=128, vocab_size=120,
                           padded_vocab_size=128, block_size=64, mlp_class_name="LLaMAMoE", n_expert=4, n_expert_per_token=2)
    prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6, 7])]
    with dry_pipeline() as (calls, lib):
        st = build_stage(cfg, "starter", 2).to(torch.bfloat16)
        st.max_seq_length = 32
        pipe = DevicePipeline(st, 0, 1, n_samples=2, max_seq_length=32, sampling=SamplingParams.greedy())
        out = pipe.generate(prompts, 3, mode="device")
        assert out[1].shape == (1, 4 + 3)
        gated = [c[1] for c in calls if c[0] == "gemm" and c[1].get("w2") is not None]
        assert len(gated) == 2 * 2 * 2  # 2 prompts x 2 layers x 2 experts (all-zero logits route to experts 0 and 1)
        full = [g for g in _FakeGraph.instances if "moe_router" in g.kernels]
        assert full and all(g.kernels.count("moe_linear_decode") == 4 * g.kernels.count("moe_router") for g in full)
        assert sum(g.launched * g.kernels.count("embed") for g in full) == 2 * 2  # rounds 1..2 x 2 samples
    with dry_pipeline() as (calls, lib):
        st = build_stage(cfg, "starter", 1).to(torch.bfloat16)
        st.max_seq_length = 32
        p0 = DevicePipeline(st, 0, 2, n_samples=2, max_seq_length=32, sampling=SamplingParams.greedy(), exportable=False)
        p0.next_hop, p0.next_prefill_ptr = HopTarget(0x10000, 0x20000), 0x30000
        p0.prepare(prompts, 3)
        p0.prefill()
        assert lib.mdi_copy_signal.call_count == 2  # the routed MLP's output leaves through the copy + signal kernel
        args = lib.mdi_copy_signal.call_args_list
        assert args[0][0][1] == 0x30000 and args[0][0][3] == 0x20000
