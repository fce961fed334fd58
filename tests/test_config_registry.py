import importlib.util
import os

import pytest

from mdi_llm_b200.models.config import Config, find_multiple
from mdi_llm_b200.models.registry import configs, lookup, name_to_config

REF_CFG = "/root/reference/src/sub/config.py"


def test_find_multiple():
    assert find_multiple(50254, 512) == 50688
    assert find_multiple(512, 512) == 512


def test_llama3_8b_shape():
    c = Config.from_name("Llama-3-8B")
    assert (c.n_layer, c.n_embd, c.n_head, c.n_query_groups, c.head_size) == (32, 4096, 32, 8, 128)
    assert c.qkv_size == 6144 and c.intermediate_size == 14336 and c.padded_vocab_size == 128256
    assert c.rope_n_elem == 128 and c.rope_base == 500000
    # lm_head costs ~2.4 blocks of bytes
    assert 2.3 < c.head_param_count() / c.block_param_count() < 2.5


def test_lookup_by_hf_name_and_yaml_roundtrip(tmp_path):
    c = Config.from_name("Meta-Llama-3-8B-Instruct")
    assert c.name == "Llama-3-8B-Instruct"
    c.save(tmp_path)
    c2 = Config.from_checkpoint(tmp_path)
    assert c2.asdict() == c.asdict()
    assert "pos_embedding" not in c.asdict()  # schema stays the reference's for litGPT models


def test_gpt2_family_extension(tmp_path):
    c = Config.from_name("gpt2")
    assert c.pos_embedding == "learned" and c.tie_embeddings and c.rope_n_elem == 0
    c.save(tmp_path)
    assert Config.from_file(tmp_path / "model_config.yaml").pos_embedding == "learned"


def test_unknown_name():
    with pytest.raises(ValueError):
        lookup("no-such-model")


@pytest.mark.skipif(not os.path.isfile(REF_CFG), reason="reference tree not mounted")
def test_registry_matches_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # the reference has a `typing.py` that shadows stdlib when cwd=sub/
    spec = importlib.util.spec_from_file_location("_refcfg", REF_CFG)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    assert len(ref.configs) <= len(configs)
    for rc in ref.configs:
        mine = Config(**name_to_config[rc["name"]])
        theirs = Config(**rc)
        assert mine.asdict() == theirs.asdict(), rc["name"]


@pytest.mark.skipif(not os.path.isfile(REF_CFG), reason="reference not mounted")
def test_family_lists_match_reference(tmp_path, monkeypatch):
    """The per-family module-level lists of the reference (`from sub.config import llama_3` ...) exist here as
    views of the table-driven registry, with the same members."""
    import mdi_llm_b200.config as C

    monkeypatch.chdir(tmp_path)
    spec = importlib.util.spec_from_file_location("_refcfg2", REF_CFG)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    families = ["stablecode", "pythia", "dolly", "redpajama_incite", "falcon", "falcon180b", "open_LLaMA", "vicuna", "long_chat",
                "nous_research", "llama_2", "llama_3", "gemma", "codegemma", "danube2", "freewilly_2", "code_llama", "platypus",
                "together_llama2_32k", "phi", "mistral", "tiny_llama", "llama_2_function_calling"]

    all_names = {c["name"] for c in configs}

    def members(obj):  # the reference keeps lists of dicts, single dicts and "{}" name templates
        items = [obj] if isinstance(obj, dict) else [c for c in obj if isinstance(c, dict)]
        out = set()
        for c in items:
            n = c["name"]
            if "{}" in n:
                head, tail = n.split("{}")
                out |= {x for x in all_names if x.startswith(head) and x.endswith(tail)}
            else:
                out.add(n)
        return out

    for fam in families:
        theirs, mine = members(getattr(ref, fam)), {c["name"] for c in getattr(C, fam)}
        assert theirs and theirs <= mine, (fam, sorted(theirs - mine))
    assert C.PLOTS is False


def test_compat_names_exist():
    from mdi_llm_b200.models.gpt import KVCache, multinomial_num_samples_1
    from mdi_llm_b200.models.stage import NodePrototype, StageModule
    from mdi_llm_b200.parallel.transport.socket_transport import InputNodeConnection, NodeConnection, OutputNodeConnection
    from mdi_llm_b200.utils.misc import format_output
    import torch

    assert NodePrototype is StageModule and issubclass(InputNodeConnection, NodeConnection) and issubclass(OutputNodeConnection, NodeConnection)
    kv = KVCache((1, 2, 8, 4), (1, 2, 8, 4), dtype=torch.float32)
    k, v = kv(torch.tensor([3]), torch.ones(1, 2, 1, 4), torch.full((1, 2, 1, 4), 2.0))
    assert k[0, 0, 3].tolist() == [1.0] * 4 and v[0, 1, 3].tolist() == [2.0] * 4 and float(k.sum()) == 8.0
    kv.reset_parameters()
    assert float(kv.k.abs().sum()) == 0.0
    assert multinomial_num_samples_1(torch.tensor([[0.0, 1.0, 0.0]])).tolist() == [[1]]
    txt = format_output("<|user|>hi there<|assistant|>hello!<|user|>bye")
    assert txt == "User: hi there\n\nAssistant: hello!\n\nUser: bye"


def test_legacy_gptconfig_factory():
    import torch
    from mdi_llm_b200.models.config import GPTConfig
    from mdi_llm_b200.models.gpt import GPT

    cfg = GPTConfig(block_size=32, vocab_size=64, n_layer=2, n_head=2, n_embd=16, dropout=0.1)
    assert cfg.pos_embedding == "learned" and cfg.tie_embeddings and cfg.norm_class_name == "LayerNorm"
    m = GPT(cfg).eval()
    assert m.lm_head.weight is m.transformer.wte.weight
    assert m(torch.tensor([[1, 2, 3]])).shape == (1, 3, 64)
    with pytest.raises(ValueError):
        GPTConfig(activation_function="ReLU")
