"""litGPT → Hugging Face checkpoint conversion (inverse of :mod:`.convert_hf_checkpoint`).

Parity: reference ``src/sub/utils/convert_lit_checkpoint.py`` — ``convert_lit_checkpoint(
checkpoint_dir, output_dir)`` (:232-266), ``qkv_split`` undoing the per-group interleave
(:215-230), name maps for Llama/Mistral/Gemma, GPT-NeoX, Falcon and Phi (:16-213); writes
``model.pth`` in ``output_dir``.  Built as the same rule table run backwards.
"""
from __future__ import annotations

import gc
import re
from pathlib import Path
from typing import Dict, Tuple, Union

import torch

from ..models.config import Config
from .checkpoint import incremental_save, lazy_load
from .convert_hf_checkpoint import hf_family

__all__ = ["convert_lit_checkpoint", "qkv_split", "convert_state_dict_to_hf"]


def qkv_split(param: torch.Tensor, config: Config) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Undo the group interleave: returns contiguous ``q [H*hs,..]``, ``k [G*hs,..]``, ``v [G*hs,..]``."""
    hs, qpk, g = config.head_size, config.q_per_kv, config.n_query_groups
    blocks = param.split((qpk + 2) * hs)
    if len(blocks) != g:
        raise ValueError("interleaved qkv tensor does not match the config")
    qs, ks, vs = [], [], []
    for b in blocks:
        q, k, v = b.split((qpk * hs, hs, hs))
        qs.append(q); ks.append(k); vs.append(v)
    return torch.cat(qs), torch.cat(ks), torch.cat(vs)


# litGPT regex -> HF template(s)
_LLAMA = [
    (r"transformer\.wte\.weight", "model.embed_tokens.weight"),
    (r"transformer\.h\.(\d+)\.norm_1\.(weight|bias)", "model.layers.{0}.input_layernorm.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.attn\.(weight|bias)", "QKV:model.layers.{0}.self_attn.{p}_proj.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.proj\.(weight|bias)", "model.layers.{0}.self_attn.o_proj.{1}"),
    (r"transformer\.h\.(\d+)\.norm_2\.(weight|bias)", "model.layers.{0}.post_attention_layernorm.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.fc_1\.(weight|bias)", "model.layers.{0}.mlp.gate_proj.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.fc_2\.(weight|bias)", "model.layers.{0}.mlp.up_proj.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.proj\.(weight|bias)", "model.layers.{0}.mlp.down_proj.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.gate\.weight", "model.layers.{0}.block_sparse_moe.gate.weight"),
    (r"transformer\.h\.(\d+)\.mlp\.experts\.(\d+)\.fc_1\.weight", "model.layers.{0}.block_sparse_moe.experts.{1}.w1.weight"),
    (r"transformer\.h\.(\d+)\.mlp\.experts\.(\d+)\.fc_2\.weight", "model.layers.{0}.block_sparse_moe.experts.{1}.w3.weight"),
    (r"transformer\.h\.(\d+)\.mlp\.experts\.(\d+)\.proj\.weight", "model.layers.{0}.block_sparse_moe.experts.{1}.w2.weight"),
    (r"transformer\.ln_f\.(weight|bias)", "model.norm.{0}"),
    (r"lm_head\.(weight|bias)", "lm_head.{0}"),
]
_NEOX = [
    (r"transformer\.wte\.weight", "gpt_neox.embed_in.weight"),
    (r"transformer\.h\.(\d+)\.norm_1\.(weight|bias)", "gpt_neox.layers.{0}.input_layernorm.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.attn\.(weight|bias)", "gpt_neox.layers.{0}.attention.query_key_value.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.proj\.(weight|bias)", "gpt_neox.layers.{0}.attention.dense.{1}"),
    (r"transformer\.h\.(\d+)\.norm_2\.(weight|bias)", "gpt_neox.layers.{0}.post_attention_layernorm.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.fc\.(weight|bias)", "gpt_neox.layers.{0}.mlp.dense_h_to_4h.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.proj\.(weight|bias)", "gpt_neox.layers.{0}.mlp.dense_4h_to_h.{1}"),
    (r"transformer\.ln_f\.(weight|bias)", "gpt_neox.final_layer_norm.{0}"),
    (r"lm_head\.weight", "embed_out.weight"),
]
_PHI = [
    (r"transformer\.wte\.weight", "model.embed_tokens.weight"),
    (r"transformer\.h\.(\d+)\.norm_1\.(weight|bias)", "model.layers.{0}.input_layernorm.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.attn\.(weight|bias)", "QKV:model.layers.{0}.self_attn.{p}_proj.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.proj\.(weight|bias)", "model.layers.{0}.self_attn.dense.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.fc\.(weight|bias)", "model.layers.{0}.mlp.fc1.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.proj\.(weight|bias)", "model.layers.{0}.mlp.fc2.{1}"),
    (r"transformer\.ln_f\.(weight|bias)", "model.final_layernorm.{0}"),
    (r"lm_head\.(weight|bias)", "lm_head.{0}"),
]


def _falcon_rules(config: Config):
    norm_1 = "input_layernorm" if config.shared_attention_norm else "ln_attn"
    return [
        (r"transformer\.wte\.weight", "transformer.word_embeddings.weight"),
        (r"transformer\.h\.(\d+)\.attn\.attn\.(weight|bias)", "transformer.h.{0}.self_attention.query_key_value.{1}"),
        (r"transformer\.h\.(\d+)\.attn\.proj\.(weight|bias)", "transformer.h.{0}.self_attention.dense.{1}"),
        (r"transformer\.h\.(\d+)\.mlp\.fc\.(weight|bias)", "transformer.h.{0}.mlp.dense_h_to_4h.{1}"),
        (r"transformer\.h\.(\d+)\.mlp\.proj\.(weight|bias)", "transformer.h.{0}.mlp.dense_4h_to_h.{1}"),
        (r"transformer\.h\.(\d+)\.norm_1\.(weight|bias)", "transformer.h.{0}." + norm_1 + ".{1}"),
        (r"transformer\.h\.(\d+)\.norm_2\.(weight|bias)", "transformer.h.{0}.ln_mlp.{1}"),
        (r"transformer\.ln_f\.(weight|bias)", "transformer.ln_f.{0}"),
        (r"lm_head\.weight", "lm_head.weight"),
    ]


_GPT2 = [
    (r"transformer\.wte\.weight", "transformer.wte.weight"),
    (r"transformer\.wpe\.weight", "transformer.wpe.weight"),
    (r"transformer\.h\.(\d+)\.norm_1\.(weight|bias)", "transformer.h.{0}.ln_1.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.attn\.(weight|bias)", "QKVCAT:transformer.h.{0}.attn.c_attn.{1}"),
    (r"transformer\.h\.(\d+)\.attn\.proj\.(weight|bias)", "T:transformer.h.{0}.attn.c_proj.{1}"),
    (r"transformer\.h\.(\d+)\.norm_2\.(weight|bias)", "transformer.h.{0}.ln_2.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.fc\.(weight|bias)", "T:transformer.h.{0}.mlp.c_fc.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.proj\.(weight|bias)", "T:transformer.h.{0}.mlp.c_proj.{1}"),
    (r"transformer\.ln_f\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"lm_head\.weight", "lm_head.weight"),
]


def _rules(config: Config):
    fam = hf_family(config)
    table = {"llama": _LLAMA, "gpt_neox": _NEOX, "phi": _PHI, "gpt2": _GPT2}.get(fam)
    if fam == "falcon":
        table = _falcon_rules(config)
    return [(re.compile(p + r"$"), t) for p, t in table]


def convert_state_dict_to_hf(lit_sd: Dict[str, torch.Tensor], config: Config, store=lambda t: t) -> Dict[str, torch.Tensor]:
    rules = _rules(config)
    out: Dict[str, torch.Tensor] = {}
    tied = config.tie_embeddings or ("lm_head.weight" in lit_sd and "transformer.wte.weight" in lit_sd and
                                     lit_sd["lm_head.weight"].data_ptr() == lit_sd["transformer.wte.weight"].data_ptr())
    for name in list(lit_sd.keys()):
        t = lit_sd[name]
        if name == "lm_head.weight" and tied and hf_family(config) in ("llama", "gpt2") and "Gemma" in config.name:
            continue  # Gemma ties the head in HF: no separate tensor
        for rx, target in rules:
            m = rx.match(name)
            if not m:
                continue
            g = m.groups()
            if target.startswith("QKV:"):
                q, k, v = qkv_split(t, config)
                for p, part in zip("qkv", (q, k, v)):
                    out[target[4:].replace("{p}", p).format(*g)] = store(part)
            elif target.startswith("QKVCAT:"):
                q, k, v = qkv_split(t, config)
                cat = torch.cat((q, k, v))
                out[target[7:].format(*g)] = store(cat.t().contiguous() if g[-1] == "weight" else cat)
            elif target.startswith("T:"):
                out[target[2:].format(*g)] = store(t.t().contiguous() if g[-1] == "weight" else t)
            else:
                out[target.format(*g)] = store(t)
            break
        else:
            raise KeyError(f"no conversion rule for litGPT tensor {name!r}")
    return out


@torch.inference_mode()
def convert_lit_checkpoint(checkpoint_dir: Union[str, Path], output_dir: Union[str, Path]) -> None:
    checkpoint_dir, output_dir = Path(checkpoint_dir), Path(output_dir)
    config = Config.from_file(checkpoint_dir / "model_config.yaml")
    output_dir.mkdir(parents=True, exist_ok=True)
    with incremental_save(output_dir / "model.pth") as saver:
        lit = lazy_load(checkpoint_dir / "lit_model.pth")
        lit = lit.get("model", lit)
        saver.save(convert_state_dict_to_hf(lit, config, store=saver.store_early))
        gc.collect()


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("checkpoint_dir", type=Path)
    ap.add_argument("output_dir", type=Path)
    a = ap.parse_args()
    convert_lit_checkpoint(a.checkpoint_dir, a.output_dir)
