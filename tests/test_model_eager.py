import torch
import pytest

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.gpt import GPT, KVPool, apply_rope, build_rope_cache, sample
from mdi_llm_b200.models.stage import build_stage
from mdi_llm_b200.models.partition import split_parameters
from mdi_llm_b200.utils.checkpoint import random_state_dict
from mdi_llm_b200.utils.functional import scaled_dot_product_attention


def _model(cfg):
    m = GPT(cfg)
    m.load_state_dict(random_state_dict(cfg, dtype=torch.float32))
    return m.eval()


VARIANTS = {
    "llama": dict(name="tiny-llama-1.1b", n_layer=3, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96),
    "neox_parallel": dict(name="pythia-14m", n_layer=2, n_embd=64, n_head=4),
    "falcon_mqa_shared_norm": dict(name="falcon-7b", n_layer=2, n_embd=64, n_head=4),
    "gemma": dict(name="Gemma-2b", n_layer=2, n_embd=64, n_head=4, intermediate_size=96),
    "phi_partial_rope": dict(name="phi-2", n_layer=2, n_embd=64, n_head=4),
    "mixtral_moe": dict(name="Mixtral-8x7B-v0.1", n_layer=2, n_embd=64, n_head=4, n_query_groups=2,
                        intermediate_size=96, n_expert=4, n_expert_per_token=2),
    "gpt2": dict(name="gpt2", n_layer=2, n_embd=64, n_head=4),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@torch.no_grad()
def test_cached_decode_matches_full_forward(variant):
    kw = dict(VARIANTS[variant])
    name = kw.pop("name")
    cfg = Config.from_name(name, block_size=32, vocab_size=100, padded_vocab_size=128, **kw)
    m = _model(cfg)
    prompt = torch.tensor([5, 9, 2, 77, 41])
    out = m.generate(prompt, 12, temperature=0.0, top_p=0.0).clone()
    full = m(out[:, :-1])  # no cache, causal
    assert full[0, 4:].argmax(-1).tolist() == out[0, 5:].tolist()


@torch.no_grad()
def test_kv_slots_are_isolated(tiny_llama_cfg):
    m = _model(tiny_llama_cfg)
    m.set_kv_cache(n_slots=2)
    a, b = torch.tensor([1, 2, 3, 4]), torch.tensor([9, 8, 7])
    oa = m.generate(a, 10, temperature=0.0, top_p=0.0, slot=0).clone()
    ob = m.generate(b, 10, temperature=0.0, top_p=0.0, slot=1).clone()
    m.set_kv_cache(n_slots=1)
    assert torch.equal(oa, m.generate(a, 10, temperature=0.0, top_p=0.0))
    m.kv_pool.reset()
    assert torch.equal(ob, m.generate(b, 10, temperature=0.0, top_p=0.0))


@torch.no_grad()
def test_stage_chain_equals_full_model(tiny_llama_cfg):
    sd = random_state_dict(tiny_llama_cfg, dtype=torch.float32)
    m = GPT(tiny_llama_cfg)
    m.load_state_dict(sd)
    m.eval()
    chunks, info = split_parameters(dict(sd), 3)
    stages = [build_stage(tiny_llama_cfg, "starter", info["plan"][0], meta=True)]
    stages[0].load_weights(chunks["starter"])
    for i, c in enumerate(chunks["secondary"]):
        s = build_stage(tiny_llama_cfg, f"secondary:{i}", info["plan"][i + 1], meta=True)
        s.load_weights(c)
        stages.append(s)
    idx = torch.tensor([[3, 1, 4, 1, 5, 9]])
    pos = torch.arange(6)
    for s in stages:
        s.set_kv_cache(1)
    x = stages[0](idx, pos)
    for s in stages[1:]:
        x = s(x, pos)
    logits = stages[0](x, first_pass=False)
    m.set_kv_cache(1)
    torch.testing.assert_close(logits, m(idx, pos), rtol=1e-5, atol=1e-5)


def test_rope_is_a_rotation():
    cos, sin = build_rope_cache(16, 8, base=10000)
    x = torch.randn(2, 4, 16, 8)
    y = apply_rope(x, cos, sin)
    torch.testing.assert_close(y.norm(dim=-1), x.norm(dim=-1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(y[..., 0, :], x[..., 0, :])  # position 0 is the identity


def test_sampling_branches():
    logits = torch.tensor([[[0.1, 3.0, 0.2, 2.9]]])
    assert sample(logits, temperature=0.0, top_p=0.0).item() == 1
    g = torch.Generator().manual_seed(0)
    draws = {sample(logits, temperature=1.0, top_k=2, generator=g).item() for _ in range(50)}
    assert draws <= {1, 3} and len(draws) == 2
    with pytest.raises(ValueError):
        sample(logits, top_p=1.5)
    # nucleus with tiny p keeps only the arg-max
    assert all(sample(logits, temperature=1.0, top_p=0.01, generator=g).item() == 1 for _ in range(10))


def test_python_sdpa_matches_torch():
    q, k, v = (torch.randn(1, 2, 7, 16) for _ in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    torch.testing.assert_close(scaled_dot_product_attention(q, k, v, is_causal=True), ref, rtol=1e-5, atol=1e-5)


def test_generate_chat_stops_and_streams(tiny_llama_cfg):
    m = _model(tiny_llama_cfg)
    prompt = torch.tensor([1, 2, 3])
    full = m.generate(prompt, 13, temperature=0.0, top_p=0.0)[0, 3:].tolist()
    m.kv_pool.reset()
    streamed = [int(t) for t in m.generate_chat(prompt, 13, temperature=0.0, top_p=0.0)]
    assert streamed == full
    stop = full[4:6]
    m.kv_pool.reset()
    cut = [int(t) for t in m.generate_chat(prompt, 13, temperature=0.0, top_p=0.0, stop_tokens=(stop,))]
    first = next(i for i in range(len(full) - 1) if full[i:i + 2] == stop)
    assert cut == full[:first]
