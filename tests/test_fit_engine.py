"""`utils/fit_engine.py`: the re-parametrisations that bring Phi-2 (head size 80) and Falcon-7B (71 query heads on one
KV head) into the fused engine's configuration space are exact — same logits, prefill and cached decoding."""
import pytest
import torch

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.gpt import GPT
from mdi_llm_b200.parallel.engine import engine_supports
from mdi_llm_b200.utils.checkpoint import random_state_dict
from mdi_llm_b200.utils.fit_engine import expand_kv_groups, fit_engine, pad_head_size

PHI = dict(n_layer=2, n_embd=60, n_head=3, n_query_groups=3, head_size=20, rotary_percentage=0.4, parallel_residual=True, shared_attention_norm=True,
           norm_class_name="LayerNorm", mlp_class_name="GptNeoxMLP", gelu_approximate="tanh", bias=True, lm_head_bias=True,
           vocab_size=100, padded_vocab_size=128, block_size=32, intermediate_size=96)
FALCON = dict(n_layer=2, n_embd=56, n_head=7, n_query_groups=1, head_size=8, rotary_percentage=1.0, parallel_residual=True,
              shared_attention_norm=True, norm_class_name="LayerNorm", mlp_class_name="GptNeoxMLP", bias=False, vocab_size=100,
              padded_vocab_size=128, block_size=32, intermediate_size=96)
LLAMA_ODD = dict(n_layer=2, n_embd=72, n_head=6, n_query_groups=2, head_size=12, vocab_size=100, padded_vocab_size=128, block_size=32,
                 intermediate_size=96)  # sequential residual, gated MLP, full rotary, 3 query heads per KV head


def _model(cfg, sd):
    m = GPT(cfg)
    m.load_state_dict(sd)
    return m.eval()


def _run(m, idx, steps=4):
    """Teacher-forced logits of a prompt, then greedy decoding on the KV cache."""
    with torch.no_grad():
        full = m(idx)
        m.max_seq_length = 32
        m.set_kv_cache(batch_size=1)
        pos = torch.arange(idx.size(1))
        logits = m(idx, pos)
        toks, outs = [], [logits[:, -1]]
        for i in range(steps):
            t = logits[:, -1].argmax(-1, keepdim=True)
            toks.append(int(t))
            logits = m(t, torch.tensor([idx.size(1) + i]))
            outs.append(logits[:, -1])
        m.clear_kv_cache()
    return full, torch.stack(outs), toks


WIDE = dict(n_layer=2, n_embd=96, n_head=12, n_query_groups=2, head_size=8, bias=True, vocab_size=100, padded_vocab_size=128, block_size=32,
            intermediate_size=96)  # 6 query heads per KV head -> groups of 2 (Falcon-40B: 16 -> 8)


@pytest.mark.parametrize("base,how", [(PHI, "pad"), (FALCON, "expand"), (LLAMA_ODD, "pad"), (LLAMA_ODD, "both"), (WIDE, "narrow2")])
def test_reparametrised_model_is_the_same_function(base, how):
    cfg = Config.from_name("tiny-llama-1.1b", **base)
    sd = random_state_dict(cfg, dtype=torch.float32, seed=3, std=0.2)
    cfg2, sd2 = cfg, dict(sd)
    if how == "narrow2":
        cfg2, sd2 = expand_kv_groups(cfg2, sd2, 2)
        assert cfg2.q_per_kv == 2 and cfg2.n_query_groups == 6
    if how in ("expand", "both"):
        cfg2, sd2 = expand_kv_groups(cfg2, sd2)
        assert cfg2.q_per_kv == 1 and cfg2.n_query_groups == cfg.n_head
    if how in ("pad", "both"):
        cfg2, sd2 = pad_head_size(cfg2, sd2, 32)
        assert cfg2.head_size == 32 and cfg2.rope_n_elem == cfg.rope_n_elem
    idx = torch.tensor([[5, 17, 3, 88, 42, 7]])
    full_a, dec_a, tok_a = _run(_model(cfg, sd), idx)
    full_b, dec_b, tok_b = _run(_model(cfg2, sd2), idx)
    torch.testing.assert_close(full_b, full_a, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dec_b, dec_a, rtol=1e-4, atol=1e-5)
    assert tok_a == tok_b
    # only the attention tensors changed
    changed = {k for k in sd if sd2[k].shape != sd[k].shape}
    assert changed and all(".attn." in k for k in changed)


def test_every_registry_model_is_inside_the_engine_natively_or_after_fitting():
    """The whole registry (the reference's 112 configurations + ours): 92 run on the fused engine as they are, the
    rest after an exact re-parametrisation — none is left to the eager fallback for its attention geometry."""
    from mdi_llm_b200.models import registry

    native, fitted = 0, {}
    for name in [c["name"] for c in registry.configs]:
        cfg = Config.from_name(name)
        if engine_supports(cfg, torch.bfloat16):
            native += 1
            continue
        new, _, notes = fit_engine(cfg)
        assert engine_supports(new, torch.bfloat16) and notes and new.n_head == cfg.n_head and new.rope_n_elem == cfg.rope_n_elem
        fitted[name] = (new.n_query_groups, new.q_per_kv, new.head_size)
    assert native >= 90 and len(fitted) >= 20
    assert fitted["falcon-40b"] == (16, 8, 64) and fitted["Gemma-2b"] == (4, 2, 256) and fitted["falcon-180B"] == (232, 1, 64)
    assert fitted["open_llama_3b"][2] == 128 and fitted["pythia-14m"][2] == 64


def test_registry_models_land_inside_the_engine():
    for name, expect in (("phi-2", ["head size 80 -> 128"]), ("falcon-7b", ["71 query heads per KV head"])):
        cfg = Config.from_name(name)
        assert not engine_supports(cfg, torch.bfloat16)
        fitted, sd, notes = fit_engine(cfg)
        assert sd is None and engine_supports(fitted, torch.bfloat16)
        assert all(any(e in n for n in notes) for e in expect), notes
        assert fitted.rope_n_elem == cfg.rope_n_elem and fitted.n_head == cfg.n_head and fitted.n_embd == cfg.n_embd
    cfg = Config.from_name("Llama-3-8B")
    assert fit_engine(cfg)[2] == [] and fit_engine(cfg)[0] == cfg  # already inside: untouched


def test_prepare_model_fit_engine_writes_a_loadable_checkpoint(tmp_path, capsys):
    from mdi_llm_b200.cli import prepare_model
    from mdi_llm_b200.utils.checkpoint import load_from_pt, write_random_checkpoint

    cfg = Config.from_name("tiny-llama-1.1b", **{**PHI, "head_size": 80, "n_embd": 240, "n_head": 3})
    ck = write_random_checkpoint(tmp_path / "custom" / "TinyPhi", cfg, dtype=torch.float32, seed=5)
    assert prepare_model.main([str(ck), "--fit-engine", "--n-nodes", "2"]) == 0
    out = capsys.readouterr().out
    assert "head size 80 -> 128" in out
    fitted_dir = ck.parent / (ck.name + "-fused")
    cfg2, sd2 = load_from_pt(fitted_dir)
    assert cfg2.head_size == 128 and engine_supports(cfg2, torch.bfloat16)
    assert (fitted_dir / "chunks" / "2nodes" / "model_starter.pth").is_file()
    _, sd = load_from_pt(ck)
    idx = torch.tensor([[1, 2, 3, 4]])
    with torch.no_grad():
        torch.testing.assert_close(_model(cfg2, sd2)(idx), _model(cfg, sd)(idx), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["phi-2", "falcon-7b"])
def test_fitted_registry_shapes_bind_on_the_fused_engine(name):
    """The launch sequence of a (2-layer, narrow) stage with the fitted attention geometry binds against the kernel
    wrappers: head size 128 with 32 rotated dimensions (Phi-2), 71 KV groups of one query head (Falcon-7B)."""
    import dataclasses

    from test_engine_dryrun import dry_ops

    from mdi_llm_b200.models.stage import build_stage
    from mdi_llm_b200.parallel.engine import FusedStage, HopTarget

    full, _, _ = fit_engine(Config.from_name(name))
    cfg = dataclasses.replace(full, n_layer=2, n_embd=full.n_head * 8, intermediate_size=128, padded_vocab_size=256, vocab_size=250,
                              block_size=64)
    assert engine_supports(cfg, torch.bfloat16)
    with dry_ops() as calls:
        st = build_stage(cfg, "starter", 2).to(torch.bfloat16)
        st.max_seq_length = 32
        fs = FusedStage(st, n_slots=2, max_seq_length=32)
        fs.enqueue_head(wait=True)
        fs.enqueue_sample()
        fs.enqueue_embed(from_tokens=True)
        n0 = len(calls)
        fs.enqueue_blocks(HopTarget(0x1000, 0x2000), wait_input=True)
    seq = calls[n0:]
    assert len(seq) == 2 * 5  # parallel-residual blocks: qkv, attention, o_proj, fc, down
    qkv, attn = seq[0][1], seq[1][1]
    assert (qkv["head_size"], qkv["rope_n_elem"], qkv["n_head"], qkv["n_groups"]) == (full.head_size, full.rope_n_elem, full.n_head,
                                                                                       full.n_query_groups)
    assert attn["head_size"] == full.head_size and attn["n_groups"] == full.n_query_groups and attn["n_split"] >= 1
    head = [c[1] for c in calls[:n0] if c[0] == "linear_decode"][0]
    assert (head.get("bias") is not None) == cfg.lm_head_bias


def test_server_suggests_fitting_when_it_falls_back_to_eager():
    import warnings

    from mdi_llm_b200.parallel.server import GPTServer

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        GPTServer._hint_fit_engine(Config.from_name("phi-2"))
        GPTServer._hint_fit_engine(Config.from_name("Llama-3-8B"))  # inside the engine: nothing to say
    assert len(w) == 1 and "--fit-engine" in str(w[0].message) and "head size 80 -> 128" in str(w[0].message)
