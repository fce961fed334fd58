"""Model configuration (litGPT-compatible) for the B200 MDI engine.

Parity target: reference ``src/sub/model.py:94-273`` (``Config`` dataclass, ``from_name``,
``from_file``, ``from_checkpoint``, ``asdict``) and the registry in
``src/sub/config.py:170-1669``.  The field set and the ``model_config.yaml`` schema are kept
so checkpoints written by the reference load unchanged.  Two things are new:

* a GPT-2 family (learned positional embedding, LayerNorm with bias, tied ``lm_head``) can be
  expressed (``pos_embedding="learned"``, ``tie_embeddings=True``) — the reference needs a
  separate legacy tree for that (``old/GPT2/sub/model.py``);
* derived, kernel-facing quantities (``q_per_kv``, ``qkv_size``, ``kv_dim`` ...) are
  properties so the CUDA ops and the eager model agree on one definition.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from pathlib import Path
from typing import Any, Dict, Optional, Union

import yaml

__all__ = ["Config", "GPTConfig", "find_multiple"]


def find_multiple(n: int, k: int) -> int:
    """Smallest multiple of ``k`` that is >= ``n``."""
    if k <= 0:
        raise ValueError("k must be positive")
    return n if n % k == 0 else n + k - (n % k)


# Fields that exist in the reference's ``Config.asdict`` (model.py:245-273) — this is the
# on-disk ``model_config.yaml`` schema.  Extension fields are only emitted when they differ
# from their defaults so that reference-written/reference-read files stay valid.
_REFERENCE_FIELDS = (
    "name", "hf_config", "scale_embeddings", "block_size", "vocab_size", "padding_multiple",
    "padded_vocab_size", "n_layer", "n_head", "head_size", "n_embd", "rotary_percentage",
    "parallel_residual", "bias", "lm_head_bias", "n_query_groups", "shared_attention_norm",
    "norm_class_name", "norm_eps", "mlp_class_name", "gelu_approximate", "intermediate_size",
    "rope_condense_ratio", "rope_base", "n_expert", "n_expert_per_token",
)
_EXTENSION_DEFAULTS = {"pos_embedding": "rope", "tie_embeddings": False}


@dataclass
class Config:
    name: str = ""
    hf_config: dict = field(default_factory=dict)
    scale_embeddings: bool = False
    block_size: int = 4096
    vocab_size: int = 50254
    padding_multiple: int = 512
    padded_vocab_size: Optional[int] = None
    n_layer: int = 16
    n_head: int = 32
    head_size: Optional[int] = None
    n_embd: int = 4096
    rotary_percentage: float = 0.25
    parallel_residual: bool = True
    bias: bool = True
    lm_head_bias: bool = False
    n_query_groups: Optional[int] = None
    shared_attention_norm: bool = False
    norm_class_name: str = "LayerNorm"  # "LayerNorm" | "RMSNorm"
    norm_eps: float = 1e-5
    mlp_class_name: str = "GptNeoxMLP"  # GptNeoxMLP | LLaMAMLP | GemmaMLP | LLaMAMoE
    gelu_approximate: str = "none"
    intermediate_size: Optional[int] = None
    rope_condense_ratio: int = 1
    rope_base: int = 10000
    n_expert: int = 0
    n_expert_per_token: int = 0
    # ---- extensions (not in the reference schema) -------------------------------------
    pos_embedding: str = "rope"  # "rope" | "learned" (GPT-2 family: adds transformer.wpe)
    tie_embeddings: bool = False  # lm_head.weight is wte.weight (GPT-2 family)

    def __post_init__(self) -> None:
        if not self.name:
            self.name = self.hf_config.get("name", self.name)
        if self.head_size is None:
            if self.n_embd % self.n_head:
                raise ValueError("n_embd must be divisible by n_head when head_size is unset")
            self.head_size = self.n_embd // self.n_head
        if self.padded_vocab_size is None:
            self.padded_vocab_size = find_multiple(self.vocab_size, self.padding_multiple)
        else:
            self.vocab_size = min(self.vocab_size, self.padded_vocab_size)
        if self.n_query_groups is None:
            self.n_query_groups = self.n_head
        elif self.n_head % self.n_query_groups:
            raise ValueError("n_head must be a multiple of n_query_groups")
        if self.intermediate_size is None:
            if self.mlp_class_name == "LLaMAMLP":
                raise ValueError(f"config {self.name!r} must set `intermediate_size`")
            self.intermediate_size = 4 * self.n_embd
        if self.pos_embedding not in ("rope", "learned"):
            raise ValueError(f"unknown pos_embedding {self.pos_embedding!r}")
        self.rope_n_elem = int(self.rotary_percentage * self.head_size)
        if self.pos_embedding == "learned":
            self.rope_n_elem = 0

    # ---- derived quantities shared by eager model and CUDA ops -------------------------
    @property
    def q_per_kv(self) -> int:
        return self.n_head // self.n_query_groups

    @property
    def qkv_size(self) -> int:
        """Rows of the fused, group-interleaved QKV projection (model.py:644)."""
        return (self.n_head + 2 * self.n_query_groups) * self.head_size

    @property
    def kv_dim(self) -> int:
        return self.n_query_groups * self.head_size

    @property
    def attn_out_dim(self) -> int:
        return self.n_head * self.head_size

    @property
    def unit_offset_norm(self) -> bool:
        """Gemma adds 1 to the RMSNorm weight (model.py:242)."""
        return self.norm_class_name == "RMSNorm" and "Gemma" in self.name

    def block_param_count(self) -> int:
        """Parameters of one transformer block (used by the balanced stage planner)."""
        c, i = self.n_embd, self.intermediate_size
        attn = self.qkv_size * c + c * self.attn_out_dim
        if self.mlp_class_name in ("LLaMAMLP", "GemmaMLP"):
            mlp = 3 * c * i
        elif self.mlp_class_name == "LLaMAMoE":
            mlp = self.n_expert * 3 * c * i + c * self.n_expert
        else:
            mlp = 2 * c * i
        return attn + mlp + 2 * c

    def head_param_count(self) -> int:
        """Parameters the starter reads per token on top of its blocks (ln_f + lm_head)."""
        return self.padded_vocab_size * self.n_embd + self.n_embd

    # ---- constructors ------------------------------------------------------------------
    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "Config":
        from .registry import lookup

        conf = dict(lookup(name))
        conf.update(kwargs)
        return cls(**conf)

    @classmethod
    def from_file(cls, path: Union[str, Path], **kwargs: Any) -> "Config":
        with open(path, encoding="utf-8") as fp:
            loaded = yaml.safe_load(fp)
        if loaded is None:
            raise ValueError(f"{path} is empty")
        known = {f.name for f in fields(cls)}
        loaded = {k: v for k, v in loaded.items() if k in known}
        loaded.update(kwargs)
        return cls(**loaded)

    @classmethod
    def from_checkpoint(cls, path: Union[str, Path], **kwargs: Any) -> "Config":
        path = Path(path)
        cfg_file = path / "model_config.yaml"
        if cfg_file.is_file():
            return cls.from_file(cfg_file, **kwargs)
        from .registry import name_to_config

        if path.name in name_to_config:
            return cls.from_name(path.name, **kwargs)
        raise FileNotFoundError(
            f"{str(path)!r}: neither 'model_config.yaml' nor a registry entry named {path.name!r}"
        )

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "Config":
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in known})

    def asdict(self) -> Dict[str, Any]:
        out = {k: getattr(self, k) for k in _REFERENCE_FIELDS}
        for k, default in _EXTENSION_DEFAULTS.items():
            if getattr(self, k) != default:
                out[k] = getattr(self, k)
        return out

    def save(self, checkpoint_dir: Union[str, Path]) -> Path:
        """Write ``model_config.yaml`` (reference ``utils.py:608-611``)."""
        p = Path(checkpoint_dir) / "model_config.yaml"
        with open(p, "w", encoding="utf-8") as fp:
            yaml.safe_dump(self.asdict(), fp)
        return p


def GPTConfig(block_size: int = 1024, vocab_size: int = 50304, n_layer: int = 12, n_head: int = 12, n_embd: int = 768,
              dropout: float = 0.0, bias: bool = True, activation_function: str = "GELU", **extra: Any) -> Config:
    """The second-generation (GPT-2 / nanoGPT) configuration object as a factory for :class:`Config`
    (reference ``old/GPT2/sub/model.py:38-62``): learned positions, LayerNorm, GELU MLP, tied head.
    ``dropout`` is accepted and ignored (inference / the trainer here run without dropout)."""
    if activation_function.upper() != "GELU":
        raise ValueError("only the GELU MLP of GPT-2 is supported")
    return Config(name=extra.pop("name", f"gpt2-custom-{n_layer}l"), block_size=block_size, vocab_size=vocab_size,
                  padded_vocab_size=vocab_size, n_layer=n_layer, n_head=n_head, n_embd=n_embd, rotary_percentage=0.0,
                  parallel_residual=False, bias=bias, norm_class_name="LayerNorm", mlp_class_name="GptNeoxMLP",
                  gelu_approximate="tanh", pos_embedding="learned", tie_embeddings=True, **extra)
