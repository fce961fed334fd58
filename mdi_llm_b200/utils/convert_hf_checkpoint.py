"""Hugging Face → litGPT checkpoint conversion.

Parity: reference ``src/sub/utils/convert_hf_checkpoint.py`` — ``convert_hf_checkpoint(
checkpoint_dir, model_name, dtype)`` (:305-388) with the GPT-NeoX, Falcon, Llama/Mistral/
Mixtral/Gemma and Phi weight maps (:18-303), the per-group **interleaving of q/k/v into one
``attn.attn.weight``** (:183-198), lazy shard loading + incremental save, ``model_config.yaml``
written next to ``lit_model.pth`` and ``lm_head`` tied to ``wte`` when absent (:180-181).
Added: the HF GPT-2 family (Conv1D transposes; reference ``old/GPT2/sub/utils.py:406-470``).

The implementation is a rule table (regex → target name or a q/k/v staging slot) instead of one
copy function per family; q/k/v rows may arrive from different shards, so they are staged per
layer until complete.
"""
from __future__ import annotations

import gc
import json
import re
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch

from ..models.config import Config
from .checkpoint import incremental_save, lazy_load

__all__ = ["convert_hf_checkpoint", "interleave_qkv", "convert_state_dict", "hf_family"]

Rule = Tuple[str, Optional[str]]  # (regex on the HF name, litGPT template | None = drop | "QKV:<part>")


def interleave_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, config: Config) -> torch.Tensor:
    """Stack per query group: ``q_per_kv`` query heads, 1 key head, 1 value head (model.py:686-699).
    Works for weights ``[rows, C]`` and biases ``[rows]``."""
    hs, qpk, g = config.head_size, config.q_per_kv, config.n_query_groups
    qs = q.split(hs * qpk)
    ks, vs = k.split(hs), v.split(hs)
    if not (len(qs) == len(ks) == len(vs) == g):
        raise ValueError(f"q/k/v shapes {tuple(q.shape)}/{tuple(k.shape)}/{tuple(v.shape)} do not match the config")
    return torch.cat([t for trio in zip(qs, ks, vs) for t in trio])


_LLAMA: List[Rule] = [
    (r"model\.embed_tokens\.weight", "transformer.wte.weight"),
    (r"model\.layers\.(\d+)\.input_layernorm\.(weight|bias)", "transformer.h.{0}.norm_1.{1}"),
    (r"model\.layers\.(\d+)\.self_attn\.([qkv])_proj\.(weight|bias)", "QKV"),
    (r"model\.layers\.(\d+)\.self_attn\.o_proj\.(weight|bias)", "transformer.h.{0}.attn.proj.{1}"),
    (r"model\.layers\.(\d+)\.self_attn\.rotary_emb\.inv_freq", None),
    (r"model\.layers\.(\d+)\.post_attention_layernorm\.(weight|bias)", "transformer.h.{0}.norm_2.{1}"),
    (r"model\.layers\.(\d+)\.mlp\.gate_proj\.(weight|bias)", "transformer.h.{0}.mlp.fc_1.{1}"),
    (r"model\.layers\.(\d+)\.mlp\.up_proj\.(weight|bias)", "transformer.h.{0}.mlp.fc_2.{1}"),
    (r"model\.layers\.(\d+)\.mlp\.down_proj\.(weight|bias)", "transformer.h.{0}.mlp.proj.{1}"),
    (r"model\.layers\.(\d+)\.block_sparse_moe\.gate\.weight", "transformer.h.{0}.mlp.gate.weight"),
    (r"model\.layers\.(\d+)\.block_sparse_moe\.experts\.(\d+)\.w1\.weight", "transformer.h.{0}.mlp.experts.{1}.fc_1.weight"),
    (r"model\.layers\.(\d+)\.block_sparse_moe\.experts\.(\d+)\.w3\.weight", "transformer.h.{0}.mlp.experts.{1}.fc_2.weight"),
    (r"model\.layers\.(\d+)\.block_sparse_moe\.experts\.(\d+)\.w2\.weight", "transformer.h.{0}.mlp.experts.{1}.proj.weight"),
    (r"model\.norm\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"lm_head\.(weight|bias)", "lm_head.{0}"),
]
_NEOX: List[Rule] = [
    (r"gpt_neox\.embed_in\.weight", "transformer.wte.weight"),
    (r"gpt_neox\.layers\.(\d+)\.input_layernorm\.(weight|bias)", "transformer.h.{0}.norm_1.{1}"),
    (r"gpt_neox\.layers\.(\d+)\.attention\.query_key_value\.(weight|bias)", "transformer.h.{0}.attn.attn.{1}"),
    (r"gpt_neox\.layers\.(\d+)\.attention\.dense\.(weight|bias)", "transformer.h.{0}.attn.proj.{1}"),
    (r"gpt_neox\.layers\.(\d+)\.attention\.(rotary_emb\.inv_freq|bias|masked_bias)", None),
    (r"gpt_neox\.layers\.(\d+)\.post_attention_layernorm\.(weight|bias)", "transformer.h.{0}.norm_2.{1}"),
    (r"gpt_neox\.layers\.(\d+)\.mlp\.dense_h_to_4h\.(weight|bias)", "transformer.h.{0}.mlp.fc.{1}"),
    (r"gpt_neox\.layers\.(\d+)\.mlp\.dense_4h_to_h\.(weight|bias)", "transformer.h.{0}.mlp.proj.{1}"),
    (r"gpt_neox\.final_layer_norm\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"embed_out\.weight", "lm_head.weight"),
]
_FALCON: List[Rule] = [
    (r"transformer\.word_embeddings\.weight", "transformer.wte.weight"),
    (r"transformer\.h\.(\d+)\.self_attention\.query_key_value\.(weight|bias)", "transformer.h.{0}.attn.attn.{1}"),
    (r"transformer\.h\.(\d+)\.self_attention\.dense\.(weight|bias)", "transformer.h.{0}.attn.proj.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.dense_h_to_4h\.(weight|bias)", "transformer.h.{0}.mlp.fc.{1}"),
    (r"transformer\.h\.(\d+)\.mlp\.dense_4h_to_h\.(weight|bias)", "transformer.h.{0}.mlp.proj.{1}"),
    (r"transformer\.h\.(\d+)\.(?:input_layernorm|ln_attn)\.(weight|bias)", "transformer.h.{0}.norm_1.{1}"),
    (r"transformer\.h\.(\d+)\.ln_mlp\.(weight|bias)", "transformer.h.{0}.norm_2.{1}"),
    (r"transformer\.ln_f\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"lm_head\.weight", "lm_head.weight"),
]
_PHI: List[Rule] = [
    (r"model\.embed_tokens\.weight", "transformer.wte.weight"),
    (r"model\.layers\.(\d+)\.input_layernorm\.(weight|bias)", "transformer.h.{0}.norm_1.{1}"),
    (r"model\.layers\.(\d+)\.self_attn\.([qkv])_proj\.(weight|bias)", "QKV"),
    (r"model\.layers\.(\d+)\.self_attn\.dense\.(weight|bias)", "transformer.h.{0}.attn.proj.{1}"),
    (r"model\.layers\.(\d+)\.mlp\.fc1\.(weight|bias)", "transformer.h.{0}.mlp.fc.{1}"),
    (r"model\.layers\.(\d+)\.mlp\.fc2\.(weight|bias)", "transformer.h.{0}.mlp.proj.{1}"),
    (r"model\.final_layernorm\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"lm_head\.(weight|bias)", "lm_head.{0}"),
]
# HF GPT-2 stores linears as Conv1D ([in, out]): "T:" = transpose, "QKVCAT" = transposed [q;k;v] block
_GPT2: List[Rule] = [
    (r"(?:transformer\.)?wte\.weight", "transformer.wte.weight"),
    (r"(?:transformer\.)?wpe\.weight", "transformer.wpe.weight"),
    (r"(?:transformer\.)?h\.(\d+)\.ln_1\.(weight|bias)", "transformer.h.{0}.norm_1.{1}"),
    (r"(?:transformer\.)?h\.(\d+)\.attn\.c_attn\.(weight|bias)", "QKVCAT"),
    (r"(?:transformer\.)?h\.(\d+)\.attn\.c_proj\.(weight|bias)", "T:transformer.h.{0}.attn.proj.{1}"),
    (r"(?:transformer\.)?h\.(\d+)\.attn\.(?:bias|masked_bias)", None),
    (r"(?:transformer\.)?h\.(\d+)\.ln_2\.(weight|bias)", "transformer.h.{0}.norm_2.{1}"),
    (r"(?:transformer\.)?h\.(\d+)\.mlp\.c_fc\.(weight|bias)", "T:transformer.h.{0}.mlp.fc.{1}"),
    (r"(?:transformer\.)?h\.(\d+)\.mlp\.c_proj\.(weight|bias)", "T:transformer.h.{0}.mlp.proj.{1}"),
    (r"(?:transformer\.)?ln_f\.(weight|bias)", "transformer.ln_f.{0}"),
    (r"lm_head\.weight", "lm_head.weight"),
]
_FAMILIES = {"llama": _LLAMA, "gpt_neox": _NEOX, "falcon": _FALCON, "phi": _PHI, "gpt2": _GPT2}


def hf_family(config: Config) -> str:
    """Which HF naming scheme a config comes from (reference picks the copy function the same way,
    convert_hf_checkpoint.py:330-351)."""
    name = config.name.lower()
    if config.pos_embedding == "learned":
        return "gpt2"
    if "falcon" in name:
        return "falcon"
    if config.mlp_class_name in ("LLaMAMLP", "GemmaMLP", "LLaMAMoE"):
        return "llama"
    if "phi" in name:
        return "phi"
    return "gpt_neox"


class _Converter:
    def __init__(self, config: Config, family: str, store: Callable[[torch.Tensor], torch.Tensor],
                 dtype: Optional[torch.dtype]) -> None:
        self.cfg, self.rules = config, [(re.compile(p + r"$"), t) for p, t in _FAMILIES[family]]
        self.store, self.dtype = store, dtype
        self.out: Dict[str, torch.Tensor] = {}
        self.pending: Dict[Tuple[str, str], Dict[str, torch.Tensor]] = {}  # (layer, weight|bias) -> {q,k,v}

    def _put(self, name: str, t: torch.Tensor) -> None:
        if self.dtype is not None and t.is_floating_point():
            t = t.to(self.dtype)
        self.out[name] = self.store(t)

    def feed(self, hf_name: str, tensor: torch.Tensor) -> None:
        for rx, target in self.rules:
            m = rx.match(hf_name)
            if not m:
                continue
            if target is None:
                return
            g = m.groups()
            if target == "QKV":
                layer, part, kind = g
                slot = self.pending.setdefault((layer, kind), {})
                slot[part] = tensor
                if len(slot) == 3:
                    self._put(f"transformer.h.{layer}.attn.attn.{kind}",
                              interleave_qkv(slot["q"], slot["k"], slot["v"], self.cfg))
                    del self.pending[(layer, kind)]
                return
            if target == "QKVCAT":
                layer, kind = g
                t = tensor.t() if kind == "weight" else tensor
                q, k, v = t.split(self.cfg.n_embd, dim=0)
                self._put(f"transformer.h.{layer}.attn.attn.{kind}", interleave_qkv(q, k, v, self.cfg))
                return
            if target.startswith("T:"):
                t = tensor.t().contiguous() if g[-1] == "weight" else tensor
                self._put(target[2:].format(*g), t)
                return
            self._put(target.format(*g), tensor)
            return
        raise KeyError(f"no conversion rule for HF tensor {hf_name!r}")

    def finish(self) -> Dict[str, torch.Tensor]:
        if self.pending:
            raise RuntimeError(f"incomplete q/k/v sets for layers {sorted(k[0] for k in self.pending)}")
        if "lm_head.weight" not in self.out:  # tied embeddings (Gemma, GPT-2, ...)
            self.out["lm_head.weight"] = self.out["transformer.wte.weight"]
        return self.out


def convert_state_dict(hf_sd: Dict[str, torch.Tensor], config: Config, dtype: Optional[torch.dtype] = None,
                       family: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """In-memory conversion (tests, small models)."""
    conv = _Converter(config, family or hf_family(config), lambda t: t, dtype)
    for k, v in hf_sd.items():
        conv.feed(k, v)
    return conv.finish()


def _shards(checkpoint_dir: Path) -> List[Path]:
    for index in ("pytorch_model.bin.index.json", "model.safetensors.index.json"):
        f = checkpoint_dir / index
        if f.is_file():
            names = sorted(set(json.loads(f.read_text())["weight_map"].values()))
            return [checkpoint_dir / n for n in names]
    files = sorted(checkpoint_dir.glob("*.bin")) or sorted(checkpoint_dir.glob("*.safetensors"))
    files = [f for f in files if f.name not in ("training_args.bin",)]
    if not files:
        raise ValueError(f"Expected {str(checkpoint_dir)!r} to contain .bin or .safetensors files")
    return files


def _load_shard(path: Path) -> Dict[str, torch.Tensor]:
    if path.suffix == ".safetensors":
        from safetensors.torch import load_file

        return load_file(str(path))
    return lazy_load(path)


@torch.inference_mode()
def convert_hf_checkpoint(checkpoint_dir: Union[str, Path] = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"),
                          model_name: Optional[str] = None, dtype: Optional[Union[str, torch.dtype]] = None) -> None:
    """Convert the HF weights in ``checkpoint_dir`` into ``lit_model.pth`` + ``model_config.yaml``."""
    checkpoint_dir = Path(checkpoint_dir)
    model_name = model_name or checkpoint_dir.name
    if isinstance(dtype, str):
        dtype = getattr(torch, dtype)
    config = Config.from_name(model_name)
    config.save(checkpoint_dir)
    with incremental_save(checkpoint_dir / "lit_model.pth") as saver:
        conv = _Converter(config, hf_family(config), saver.store_early, dtype)
        for shard in _shards(checkpoint_dir):
            sd = _load_shard(shard)
            for name in list(sd.keys()):
                conv.feed(name, sd[name])
            del sd
            gc.collect()
        saver.save(conv.finish())


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("checkpoint_dir", type=Path)
    ap.add_argument("--model-name", default=None)
    ap.add_argument("--dtype", default=None)
    a = ap.parse_args()
    convert_hf_checkpoint(a.checkpoint_dir, a.model_name, a.dtype)
