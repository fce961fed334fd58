"""HF <-> litGPT weight conversion (SURVEY §4 item d: the QKV interleave round trip)."""
import pytest
import torch

from mdi_llm_b200.models.config import Config
from mdi_llm_b200.models.gpt import GPT
from mdi_llm_b200.utils.checkpoint import random_state_dict
from mdi_llm_b200.utils.convert_hf_checkpoint import convert_hf_checkpoint, convert_state_dict, interleave_qkv
from mdi_llm_b200.utils.convert_lit_checkpoint import convert_lit_checkpoint, convert_state_dict_to_hf, qkv_split

FAMILIES = {
    "llama_gqa": dict(name="tiny-llama-1.1b", n_layer=2, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96),
    "neox": dict(name="pythia-14m", n_layer=2, n_embd=64, n_head=4),
    "falcon7b": dict(name="falcon-7b", n_layer=2, n_embd=64, n_head=4),
    "falcon40b": dict(name="falcon-40b", n_layer=2, n_embd=64, n_head=4, n_query_groups=2),
    "phi": dict(name="phi-2", n_layer=2, n_embd=64, n_head=4),
    "mixtral": dict(name="Mixtral-8x7B-v0.1", n_layer=2, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96, n_expert=4),
    "gpt2": dict(name="gpt2", n_layer=2, n_embd=64, n_head=4),
}


def test_interleave_layout_matches_litgpt_definition():
    cfg = Config.from_name("tiny-llama-1.1b", n_embd=32, n_head=4, n_query_groups=2, intermediate_size=64)
    hs = cfg.head_size
    q = torch.arange(4 * hs).float().view(-1, 1).expand(-1, 3).contiguous()  # row id as value
    k = 1000 + torch.arange(2 * hs).float().view(-1, 1).expand(-1, 3).contiguous()
    v = 2000 + torch.arange(2 * hs).float().view(-1, 1).expand(-1, 3).contiguous()
    w = interleave_qkv(q, k, v, cfg)
    # group 0: q heads 0,1 ; k head 0 ; v head 0 ; group 1: q heads 2,3 ; k head 1 ; v head 1
    expect = torch.cat((q[:2 * hs], k[:hs], v[:hs], q[2 * hs:], k[hs:], v[hs:]))
    assert torch.equal(w, expect)
    q2, k2, v2 = qkv_split(w, cfg)
    assert torch.equal(q2, q) and torch.equal(k2, k) and torch.equal(v2, v)


@pytest.mark.parametrize("family", list(FAMILIES))
def test_lit_to_hf_to_lit_roundtrip(family):
    kw = dict(FAMILIES[family])
    cfg = Config.from_name(kw.pop("name"), block_size=32, vocab_size=100, padded_vocab_size=128, **kw)
    lit = random_state_dict(cfg, dtype=torch.float32)
    hf = convert_state_dict_to_hf(dict(lit), cfg)
    assert not any(k.startswith("transformer.h.0.attn.attn") for k in hf) or family in ("neox", "falcon7b", "falcon40b") or False
    back = convert_state_dict(hf, cfg)
    assert back.keys() == lit.keys()
    for k in lit:
        assert torch.equal(back[k], lit[k]), k


def test_converted_hf_llama_weights_run_identically(tmp_path):
    """HF-layout checkpoint dir -> convert_hf_checkpoint -> GPT gives the same logits as the
    litGPT weights it came from; then convert_lit_checkpoint writes model.pth back."""
    cfg = Config.from_name("tiny-llama-1.1b")  # registered name so the CLI path can resolve it
    small = Config.from_name("tiny-llama-1.1b", n_layer=2, n_embd=64, n_head=4, n_query_groups=2, intermediate_size=96,
                             vocab_size=100, padded_vocab_size=128, block_size=32)
    lit = random_state_dict(small, dtype=torch.float32)
    hf = convert_state_dict_to_hf(dict(lit), small)
    d = tmp_path / "tiny-llama-1.1b"
    d.mkdir()
    torch.save({k: v for k, v in list(hf.items())[: len(hf) // 2]}, d / "pytorch_model-00001-of-00002.bin")
    torch.save({k: v for k, v in list(hf.items())[len(hf) // 2:]}, d / "pytorch_model-00002-of-00002.bin")
    import mdi_llm_b200.utils.convert_hf_checkpoint as C

    orig = C.Config.from_name
    C.Config.from_name = classmethod(lambda cls, name, **kw: small)  # the tiny stand-in for the registry entry
    try:
        convert_hf_checkpoint(d, model_name="tiny-llama-1.1b", dtype="float32")
    finally:
        C.Config.from_name = orig
    out = torch.load(d / "lit_model.pth", weights_only=True)
    assert all(torch.equal(out[k], lit[k]) for k in lit)
    assert (d / "model_config.yaml").is_file()
    m = GPT(small)
    m.load_state_dict(out)
    idx = torch.tensor([[1, 2, 3, 4]])
    m2 = GPT(small)
    m2.load_state_dict(lit)
    torch.testing.assert_close(m(idx), m2(idx))
    convert_lit_checkpoint(d, tmp_path / "hf_out")
    hf_back = torch.load(tmp_path / "hf_out" / "model.pth", weights_only=True)
    assert all(torch.equal(hf_back[k], hf[k]) for k in hf)
