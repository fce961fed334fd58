"""Named model configurations.

Same coverage as the reference registry (``src/sub/config.py:170-1669`` — StableLM, Pythia,
Dolly, RedPajama, Falcon, OpenLLaMA, Vicuna, LongChat, Nous-Hermes, Llama-2, Llama-3, Gemma,
CodeGemma, Danube2, FreeWilly2, CodeLlama, Platypus, LLaMA-2-7B-32K, Phi, Mistral/Mixtral,
TinyLlama, Trelis) but table-driven: every family is a base dict plus per-variant overrides.
Added here (not expressible by the reference's litGPT ``Config``): the GPT-2 family that the
reference keeps in ``old/GPT2`` and the NanoLlama 304M model of ``README.md:394-399``.

``lookup(name)`` accepts either the registry name or the HF repo name (``hf_config.name``).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Tuple

Variant = Tuple[str, str, Dict[str, Any]]  # (registry name, hf name or "", overrides)

configs: List[Dict[str, Any]] = []


def _add(org: str, base: Dict[str, Any], variants: Iterable[Variant]) -> None:
    for name, hf_name, over in variants:
        c = dict(base)
        c.update(over)
        c["name"] = name
        c["hf_config"] = {"org": org, "name": hf_name or name}
        configs.append(c)


# ---- shared bases ----------------------------------------------------------------------------
_NEOX: Dict[str, Any] = {}  # dataclass defaults == GPT-NeoX style (parallel residual, LayerNorm, bias)
_LLAMA = dict(
    rotary_percentage=1.0, parallel_residual=False, bias=False,
    norm_class_name="RMSNorm", mlp_class_name="LLaMAMLP",
)
# (n_layer, n_head, n_embd, intermediate_size) of the LLaMA-1/2 size ladder
_L7 = dict(n_layer=32, intermediate_size=11008)
_L13 = dict(n_layer=40, n_head=40, n_embd=5120, intermediate_size=13824)
_L30 = dict(n_layer=60, n_head=52, n_embd=6656, intermediate_size=17920)
_L34 = dict(n_layer=48, n_head=64, n_embd=8192, n_query_groups=8, intermediate_size=22016)
_L70 = dict(n_layer=80, n_head=64, n_embd=8192, n_query_groups=8, intermediate_size=28672)


def _m(*dicts: Dict[str, Any], **kw: Any) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for d in dicts:
        out.update(d)
    out.update(kw)
    return out


# ---- Stability AI ----------------------------------------------------------------------------
_add("stabilityai", _NEOX, [
    ("stablelm-base-alpha-3b", "", {}),
    ("stablelm-base-alpha-7b", "", dict(n_head=48, n_embd=6144, padding_multiple=256)),
    ("stablelm-tuned-alpha-3b", "", {}),
    ("stablelm-tuned-alpha-7b", "", dict(n_head=48, n_embd=6144, padding_multiple=256)),
])
_STABLE3B = dict(padded_vocab_size=50304, n_layer=32, n_embd=2560, parallel_residual=False,
                 bias=False, mlp_class_name="LLaMAMLP", intermediate_size=6912)
_add("stabilityai", _STABLE3B, [
    ("stablelm-3b-4e1t", "", {}),
    ("stablelm-zephyr-3b", "", {}),
    ("stable-code-3b", "", dict(block_size=16384)),
])
_add("stabilityai", dict(vocab_size=49152, n_layer=32, n_embd=2560), [
    ("stablecode-completion-alpha-3b", "", dict(block_size=16384)),
    ("stablecode-completion-alpha-3b-4k", "", {}),
    ("stablecode-instruct-alpha-3b", "", {}),
])

# ---- EleutherAI Pythia / Databricks Dolly -----------------------------------------------------
_PYTHIA = {
    "14m": dict(block_size=512, n_layer=6, n_embd=128, n_head=4, padding_multiple=128),
    "31m": dict(block_size=1024, n_layer=6, n_embd=256, n_head=8, padding_multiple=128),
    "70m": dict(block_size=2048, n_layer=6, n_embd=512, n_head=8, padding_multiple=128),
    "160m": dict(block_size=2048, n_layer=12, n_embd=768, n_head=12, padding_multiple=128),
    "410m": dict(block_size=2048, n_layer=24, n_embd=1024, n_head=16, padding_multiple=128),
    "1b": dict(block_size=2048, n_embd=2048, n_head=8, padding_multiple=128),
    "1.4b": dict(block_size=2048, n_layer=24, n_embd=2048, n_head=16, padding_multiple=128),
    "2.8b": dict(block_size=2048, n_layer=32, n_embd=2560, padding_multiple=128),
    "6.9b": dict(block_size=2048, n_layer=32, padding_multiple=256),
    "12b": dict(block_size=2048, n_layer=36, n_embd=5120, n_head=40),
}
_add("EleutherAI", _NEOX, [(f"pythia-{k}", "", v) for k, v in _PYTHIA.items()])
_add("EleutherAI", _NEOX,
     [(f"pythia-{k}-deduped", "", v) for k, v in _PYTHIA.items() if k not in ("14m", "31m")])
_add("databricks", dict(block_size=2048, padded_vocab_size=50280), [
    ("dolly-v2-3b", "", dict(n_layer=32, n_embd=2560)),
    ("dolly-v2-7b", "", dict(n_layer=32)),
    ("dolly-v2-12b", "", dict(n_layer=36, n_embd=5120, n_head=40)),
])

# ---- Together RedPajama-INCITE ----------------------------------------------------------------
_RP = dict(block_size=2048, n_layer=32, padding_multiple=256, rotary_percentage=1.0,
           parallel_residual=False)
_add("togethercomputer", _RP,
     [(f"RedPajama-INCITE-{k}-3B-v1", "", dict(n_embd=2560)) for k in ("Base", "Chat", "Instruct")]
     + [(f"RedPajama-INCITE-7B-{k}", "", {}) for k in ("Base", "Chat", "Instruct")]
     + [(f"RedPajama-INCITE-{k}-7B-v0.1", "", {}) for k in ("Base", "Chat", "Instruct")])

# ---- TII Falcon -------------------------------------------------------------------------------
_FALCON = dict(block_size=2048, vocab_size=65024, padded_vocab_size=65024, rotary_percentage=1.0,
               bias=False)
_F7 = dict(n_layer=32, n_head=71, n_embd=4544, n_query_groups=1, shared_attention_norm=True)
_F40 = dict(n_layer=60, n_head=128, n_embd=8192, n_query_groups=8)
_F180 = dict(n_layer=80, n_head=232, n_embd=14848, n_query_groups=8)
_add("tiiuae", _FALCON, [
    ("falcon-7b", "", _F7), ("falcon-7b-instruct", "", _F7),
    ("falcon-40b", "", _F40), ("falcon-40b-instruct", "", _F40),
    ("falcon-180B", "", _F180), ("falcon-180B-chat", "", _F180),
])

# ---- LLaMA-1 derived (norm_eps 1e-6, 2k context) ----------------------------------------------
_LLAMA1 = _m(_LLAMA, block_size=2048, vocab_size=32000, padding_multiple=64, norm_eps=1e-6)
_add("openlm-research", _LLAMA1, [
    ("open_llama_3b", "", dict(n_layer=26, n_embd=3200, intermediate_size=8640)),
    ("open_llama_7b", "", _L7),
    ("open_llama_13b", "", _L13),
])
_add("lmsys", _LLAMA1, [
    ("vicuna-7b-v1.3", "", _L7), ("vicuna-13b-v1.3", "", _L13), ("vicuna-33b-v1.3", "", _L30),
    ("longchat-7b-16k", "", _m(_L7, block_size=16384, rope_condense_ratio=8)),
    ("longchat-13b-16k", "", _m(_L13, block_size=16384, rope_condense_ratio=8)),
])

# ---- Llama-2 and derivatives ------------------------------------------------------------------
_LLAMA2 = _m(_LLAMA, vocab_size=32000, padding_multiple=64)
_add("lmsys", _LLAMA2, [
    ("vicuna-7b-v1.5", "", _L7),
    ("vicuna-7b-v1.5-16k", "", _m(_L7, block_size=16384, rope_condense_ratio=4)),
    ("vicuna-13b-v1.5", "", _L13),
    ("vicuna-13b-v1.5-16k", "", _m(_L13, block_size=16384, rope_condense_ratio=4)),
])
_add("NousResearch", _LLAMA, [
    ("Nous-Hermes-llama-2-7b", "", _m(_L7, padded_vocab_size=32000)),
    ("Nous-Hermes-13b", "", _m(_L13, block_size=2048, vocab_size=32000, padded_vocab_size=32001,
                               norm_eps=1e-6)),
    ("Nous-Hermes-Llama2-13b", "", _m(_L13, vocab_size=32000, padded_vocab_size=32032)),
])
_add("meta-llama", _LLAMA2, [
    ("Llama-2-7b-hf", "", _L7), ("Llama-2-7b-chat-hf", "", _L7),
    ("Llama-2-13b-hf", "", _L13), ("Llama-2-13b-chat-hf", "", _L13),
    ("Llama-2-70b-hf", "", _L70), ("Llama-2-70b-chat-hf", "", _L70),
])

# ---- Llama-3 (BASELINE flagship) --------------------------------------------------------------
_LLAMA3 = _m(_LLAMA, block_size=8192, vocab_size=128000, padded_vocab_size=128256,
             rope_base=500000)
_L3_8B = dict(n_layer=32, n_query_groups=8, intermediate_size=14336)
_add("meta-llama", _LLAMA3, [
    ("Llama-3-8B", "Meta-Llama-3-8B", _L3_8B),
    ("Llama-3-8B-Instruct", "Meta-Llama-3-8B-Instruct", _L3_8B),
    ("Llama-3-70B", "Meta-Llama-3-70B", _L70),
    ("Llama-3-70B-Instruct", "Meta-Llama-3-70B-Instruct", _L70),
])

# ---- Google Gemma -----------------------------------------------------------------------------
_GEMMA = _m(_LLAMA, scale_embeddings=True, vocab_size=256000, padding_multiple=64,
            mlp_class_name="GemmaMLP", gelu_approximate="tanh")
_G2B = dict(n_embd=2048, n_layer=18, n_head=8, n_query_groups=1, intermediate_size=16384)
_G7B = dict(n_embd=3072, n_layer=28, n_head=16, head_size=256, intermediate_size=24576)
_add("google", _GEMMA, [
    ("Gemma-2b", "gemma-2b", _G2B), ("Gemma-7b", "gemma-7b", _G7B),
    ("Gemma-2b-it", "gemma-2b-it", _G2B), ("Gemma-7b-it", "gemma-7b-it", _G7B),
    ("CodeGemma-7b-it", "codegemma-7b-it", _G7B),
])

# ---- misc Llama-2 shaped ----------------------------------------------------------------------
_add("h2oai", _LLAMA2, [
    ("Danube2-1.8b-chat", "h2o-danube2-1.8b-chat",
     dict(n_layer=24, n_embd=2560, intermediate_size=6912, n_query_groups=8)),
])
_add("stabilityai", _LLAMA2, [("FreeWilly2", "", _L70)])

# ---- CodeLlama --------------------------------------------------------------------------------
_CL = _m(_LLAMA, block_size=16384, rope_base=1000000)
_CL16 = dict(vocab_size=32016, padding_multiple=16)
_CL00 = dict(vocab_size=32000, padded_vocab_size=32000)
_add("codellama", _CL, [
    ("CodeLlama-7b-hf", "", _m(_L7, _CL16)), ("CodeLlama-13b-hf", "", _m(_L13, _CL16)),
    ("CodeLlama-34b-hf", "", _m(_L34, _CL00)), ("CodeLlama-70b-hf", "", _m(_L70, _CL16)),
    ("CodeLlama-7b-Python-hf", "", _m(_L7, _CL00)), ("CodeLlama-13b-Python-hf", "", _m(_L13, _CL00)),
    ("CodeLlama-34b-Python-hf", "", _m(_L34, _CL00)), ("CodeLlama-70b-Python-hf", "", _m(_L70, _CL16)),
    ("CodeLlama-7b-Instruct-hf", "", _m(_L7, _CL16)),
    ("CodeLlama-13b-Instruct-hf", "", _m(_L13, _CL16, block_size=2048)),
    ("CodeLlama-34b-Instruct-hf", "", _m(_L34, _CL00)), ("CodeLlama-70b-Instruct-hf", "", _m(_L70, _CL16)),
])

# ---- Platypus ---------------------------------------------------------------------------------
_PLAT = _m(_LLAMA, padded_vocab_size=32000)
_L70_MHA = {k: v for k, v in _L70.items() if k != "n_query_groups"}
_add("garage-bAInd", _PLAT, [
    ("Platypus-30B", "", _m(_L30, block_size=2048, norm_eps=1e-6)),
    ("Platypus2-7B", "", _L7), ("Platypus2-13B", "", _L13),
    ("Platypus2-70B", "", _L70_MHA),  # the reference registers this one without GQA
    ("Camel-Platypus2-13B", "", _L13), ("Camel-Platypus2-70B", "", _L70),
    ("Stable-Platypus2-13B", "", _L13), ("Platypus2-70B-instruct", "", _L70),
])
_add("togethercomputer", _LLAMA2, [("LLaMA-2-7B-32K", "", _m(_L7, rope_condense_ratio=8))])

# ---- Microsoft Phi ----------------------------------------------------------------------------
_PHI = dict(vocab_size=50257, padded_vocab_size=51200, block_size=2048, shared_attention_norm=True,
            lm_head_bias=True, gelu_approximate="tanh")
_add("microsoft", _PHI, [
    ("phi-1_5", "", dict(n_embd=2048, n_layer=24, rotary_percentage=0.5)),
    ("phi-2", "", dict(n_embd=2560, n_layer=32, rotary_percentage=0.4)),
])

# ---- Mistral / Mixtral (sliding window not implemented -> capped contexts as in the reference) --
_MISTRAL = _m(_LLAMA, padded_vocab_size=32000, n_layer=32, n_query_groups=8, intermediate_size=14336)
_MOE = dict(block_size=32768, mlp_class_name="LLaMAMoE", rope_base=1000000, n_expert=8,
            n_expert_per_token=2)
_add("mistralai", _MISTRAL, [
    ("Mistral-7B-v0.1", "", {}), ("Mistral-7B-Instruct-v0.1", "", {}),
    ("Mixtral-8x7B-v0.1", "", _MOE), ("Mixtral-8x7B-Instruct-v0.1", "", _MOE),
    ("Mistral-7B-Instruct-v0.2", "", dict(block_size=32768)),
    ("Mistral-7B-v0.3", "", dict(block_size=32768, padded_vocab_size=32768)),
    ("Mistral-7B-Instruct-v0.3", "", dict(block_size=32768, padded_vocab_size=32768)),
])
_add("unsloth", _MISTRAL, [("Mistral-7B-v0.2", "", dict(block_size=32768))])

# ---- TinyLlama (BASELINE config #4) / Trelis ----------------------------------------------------
_TINY = _m(_LLAMA, block_size=2048, vocab_size=32000, padding_multiple=64, n_layer=22, n_embd=2048,
           intermediate_size=5632, n_query_groups=4)
_add("TinyLlama", _TINY, [
    ("tiny-llama-1.1b", "TinyLlama-1.1B-intermediate-step-1431k-3T", {}),
    ("tiny-llama-1.1b-chat", "TinyLlama-1.1B-Chat-v1.0", {}),
])
_add("Trelis", _LLAMA2, [
    ("Llama-2-7b-chat-hf-function-calling-v2", "", _m(_L7, norm_eps=1e-6)),
])

# ---- additions of this framework ----------------------------------------------------------------
# NanoLlama 304M (README.md:394-399: 12 layers, n_embd 1024) trained with the repo's trainer.
_add("custom", _m(_LLAMA, block_size=2048, vocab_size=32000, padding_multiple=64), [
    ("NanoLlama", "", dict(n_layer=12, n_embd=1024, n_head=16, intermediate_size=5632)),
])
# GPT-2 family (reference: old/GPT2/sub/model.py:41-62) — learned positions, tied head.
_GPT2 = dict(block_size=1024, vocab_size=50257, padded_vocab_size=50257, rotary_percentage=0.0,
             parallel_residual=False, bias=True, norm_class_name="LayerNorm",
             mlp_class_name="GptNeoxMLP", gelu_approximate="tanh", pos_embedding="learned",
             tie_embeddings=True)
_add("openai-community", _GPT2, [
    ("gpt2", "", dict(n_layer=12, n_head=12, n_embd=768)),
    ("gpt2-medium", "", dict(n_layer=24, n_head=16, n_embd=1024)),
    ("gpt2-large", "", dict(n_layer=36, n_head=20, n_embd=1280)),
    ("gpt2-xl", "", dict(n_layer=48, n_head=25, n_embd=1600)),
])

name_to_config: Dict[str, Dict[str, Any]] = {c["name"]: c for c in configs}


def lookup(name: str) -> Dict[str, Any]:
    """Resolve a registry name or an HF repo name to its config dict (model.py:184-201)."""
    if name in name_to_config:
        return name_to_config[name]
    for c in configs:
        if c["hf_config"]["name"] == name:
            return c
    raise ValueError(f"{name!r} is not a supported config name")
