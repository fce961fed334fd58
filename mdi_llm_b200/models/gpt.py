"""Eager (plain PyTorch) decoder-only transformer — the numerical oracle and the CPU path.

Behavioural parity with the reference ``src/sub/model.py`` (``GPT`` :276, ``Block`` :576,
``CausalSelfAttention`` :632, MLPs :782-853, ``build_rope_cache`` :856, ``apply_rope`` :881,
``KVCache`` :894, ``RMSNorm`` :950, ``sample`` :67) with the same state-dict key names
(``transformer.wte.weight``, ``transformer.h.{l}.attn.attn.weight`` ...), so litGPT
checkpoints load unchanged.  Design differences (B200-first):

* the KV cache stores the ``n_query_groups`` heads only (the reference expands K/V to
  ``n_head`` before caching, model.py:704-714 — 4x more memory for Llama-3) and lives in a
  slot pool (:class:`KVPool`) indexed by *sample slot*, which is what the CUDA kernels
  index on-device; nothing is swapped by attribute assignment (gptserver.py:975-978);
* attention under a cache scans only the live prefix ``[0, max(input_pos)]`` instead of the
  whole ``S`` slots through a ``[1,1,S,S]`` boolean mask (model.py:940-947);
* a GPT-2 family (learned ``wpe``) is expressible in the same class.
"""
from __future__ import annotations

import time
from typing import Any, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import Config

__all__ = [
    "GPT", "Block", "CausalSelfAttention", "RMSNorm", "KVPool", "build_rope_cache", "build_mask_cache",
    "apply_rope", "sample", "sample_top_p", "multinomial_num_samples_1", "KVCache", "build_norm", "build_mlp", "LLaMAMLP",
    "GptNeoxMLP", "GemmaMLP", "LLaMAMoE",
]


# =============================================================================================
# sampling
# =============================================================================================
def multinomial_num_samples_1(probs: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One draw per row (reference model.py:34-39)."""
    return torch.multinomial(probs, num_samples=1, generator=generator)


def sample_top_p(logits: torch.Tensor, top_p: float) -> torch.Tensor:
    """Nucleus filtering on a 1-D logits vector (reference model.py:42-64)."""
    order = torch.argsort(logits, descending=False)
    sorted_logits = logits[order]
    cdf = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    drop_sorted = cdf <= (1.0 - top_p)
    drop_sorted[-1] = False  # always keep the most likely token
    drop = torch.zeros_like(drop_sorted).scatter(0, order, drop_sorted)
    return logits.masked_fill(drop, float("-inf"))


def sample(
    logits: torch.Tensor,
    temperature: float = 1.0,
    top_k: Optional[int] = None,
    top_p: float = 1.0,
    generator: Optional[torch.Generator] = None,
) -> torch.Tensor:
    """Pick the next token from ``logits[0, -1]``; returns a tensor of shape ``(1,)``.

    Semantics of reference ``model.py:67-90``: top-k crop, then (if ``temperature > 0`` or
    ``top_p > 0``) temperature scaling, optional nucleus crop, softmax and one multinomial
    draw; otherwise ``argmax`` (the greedy branch used by the equivalence tests).
    """
    if not 0.0 <= top_p <= 1.0:
        raise ValueError(f"top_p must be in [0, 1], got {top_p}")
    row = logits[0, -1] if logits.dim() == 3 else logits.reshape(-1)
    if top_k is not None:
        k = min(int(top_k), row.size(-1))
        vals, idx = torch.topk(row, k)
        row = torch.full_like(row, float("-inf")).scatter_(-1, idx, vals)
    if temperature > 0.0 or top_p > 0.0:
        if temperature > 0.0:
            row = row / temperature
        if top_p < 1.0:
            row = sample_top_p(row, top_p)
        probs = F.softmax(row.float(), dim=-1)
        return torch.multinomial(probs, 1, generator=generator)
    return torch.argmax(row, dim=-1, keepdim=True)


# =============================================================================================
# building blocks
# =============================================================================================
class RMSNorm(nn.Module):
    """fp32 RMS normalisation; ``add_unit_offset`` is the Gemma variant (model.py:950-980)."""

    def __init__(self, size: int, eps: float = 1e-6, add_unit_offset: bool = False) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(size))
        self.eps = eps
        self.add_unit_offset = add_unit_offset

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + self.eps)
        y = xf.to(x.dtype)
        w = (1 + self.weight) if self.add_unit_offset else self.weight
        return y * w

    def reset_parameters(self) -> None:
        nn.init.ones_(self.weight)


def build_norm(config: Config) -> nn.Module:
    if config.norm_class_name == "RMSNorm":
        return RMSNorm(config.n_embd, eps=config.norm_eps, add_unit_offset=config.unit_offset_norm)
    if config.norm_class_name == "LayerNorm":
        return nn.LayerNorm(config.n_embd, eps=config.norm_eps)
    raise ValueError(f"unknown norm class {config.norm_class_name!r}")


def build_rope_cache(
    seq_len: int,
    n_elem: int,
    device: Optional[torch.device] = None,
    base: int = 10000,
    condense_ratio: int = 1,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin tables of shape ``[seq_len, n_elem]`` (reference model.py:856-878)."""
    if n_elem <= 0:
        z = torch.zeros(seq_len, 0, device=device)
        return z, z.clone()
    inv_freq = 1.0 / (base ** (torch.arange(0, n_elem, 2, device=device).float() / n_elem))
    pos = torch.arange(seq_len, device=device).float() / condense_ratio
    ang = torch.outer(pos, inv_freq)
    ang = torch.cat((ang, ang), dim=-1)
    return torch.cos(ang), torch.sin(ang)


def build_mask_cache(max_seq_length: int, device: Optional[torch.device] = None) -> torch.Tensor:
    """Lower-triangular boolean mask ``[1,1,S,S]`` (reference model.py:940-947).  The reference
    indexes rows of it by ``input_pos`` on every decode step and so always attends over all ``S``
    cache slots; this framework masks by the live length instead and keeps the helper for parity."""
    ones = torch.ones((max_seq_length, max_seq_length), device=device, dtype=torch.bool)
    return torch.tril(ones).unsqueeze(0).unsqueeze(0)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """NeoX-style half rotation of the last dim of ``x [..., T, n_elem]`` (model.py:881-891)."""
    half = x.size(-1) // 2
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    return (x * cos + rot * sin).to(x.dtype)


class KVCache(nn.Module):
    """Stand-alone per-layer KV cache with the reference's interface (model.py:894-937): ``forward(input_pos,
    k, v)`` writes the rows at ``input_pos`` and returns the whole cache.  The models here use slots of a
    :class:`KVPool` instead (G group heads, indexed by sample on the device); this class is kept for code written
    against the reference's ``KVCache``."""

    def __init__(self, k_shape: Tuple[int, int, int, int], v_shape: Tuple[int, int, int, int],
                 device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None) -> None:
        super().__init__()
        self.register_buffer("k", torch.zeros(k_shape, device=device, dtype=dtype), persistent=False)
        self.register_buffer("v", torch.zeros(v_shape, device=device, dtype=dtype), persistent=False)

    def forward(self, input_pos: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        self.k = self.k.to(k.dtype)
        self.v = self.v.to(v.dtype)
        return self.k.index_copy_(2, input_pos, k), self.v.index_copy_(2, input_pos, v)

    def reset_parameters(self) -> None:
        torch.nn.init.zeros_(self.k)
        torch.nn.init.zeros_(self.v)


class KVPool:
    """Per-sample KV-cache slots for a contiguous range of layers.

    One tensor ``[n_layers, n_slots, 2, G, S, hs]`` (k at index 0, v at 1).  A *slot* is one
    sample's cache; the recurrent-pipeline scheduler maps ``sample_id -> slot`` and the CUDA
    decode kernels receive the slot as a device scalar.  Replaces the reference's per-sample
    lists of ``KVCache`` modules (gptserver.py:751-784) and stores G (not H) heads.
    """

    def __init__(
        self,
        config: Config,
        n_layers: int,
        n_slots: int,
        max_seq_length: int,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
    ) -> None:
        self.n_layers, self.n_slots, self.max_seq_length = n_layers, n_slots, max_seq_length
        self.data = torch.zeros(
            n_layers, n_slots, 2, config.n_query_groups, max_seq_length, config.head_size,
            device=device, dtype=dtype,
        )

    def layer(self, layer: int, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        blk = self.data[layer, slot]
        return blk[0], blk[1]

    def reset(self, slot: Optional[int] = None) -> None:
        with torch.inference_mode():  # the pool may have been created inside inference mode
            if slot is None:
                self.data.zero_()
            else:
                self.data[:, slot].zero_()

    @property
    def nbytes(self) -> int:
        return self.data.numel() * self.data.element_size()


class CausalSelfAttention(nn.Module):
    """Fused-QKV attention with MHA/GQA/MQA, partial RoPE and an optional KV slot.

    ``attn.weight`` rows are group-interleaved exactly like litGPT (model.py:686-699): for
    each of the ``G`` groups, ``q_per_kv`` query heads, then one key head, then one value head.
    """

    def __init__(self, config: Config) -> None:
        super().__init__()
        self.config = config
        self.attn = nn.Linear(config.n_embd, config.qkv_size, bias=config.bias)
        self.proj = nn.Linear(config.attn_out_dim, config.n_embd, bias=config.bias)

    def split_qkv(self, qkv: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``[B,T,(H+2G)hs]`` -> q ``[B,H,T,hs]``, k/v ``[B,G,T,hs]``."""
        cfg = self.config
        B, T, _ = qkv.shape
        g, qpk, hs = cfg.n_query_groups, cfg.q_per_kv, cfg.head_size
        qkv = qkv.view(B, T, g, qpk + 2, hs)
        q = qkv[:, :, :, :qpk].reshape(B, T, g * qpk, hs).transpose(1, 2)
        k = qkv[:, :, :, qpk].transpose(1, 2)
        v = qkv[:, :, :, qpk + 1].transpose(1, 2)
        return q, k, v

    def attend_qkv(
        self,
        qkv: torch.Tensor,
        cos: torch.Tensor,
        sin: torch.Tensor,
        input_pos: Optional[torch.Tensor] = None,
        kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    ) -> torch.Tensor:
        """Everything between the two projections: split, RoPE, KV-slot update, attention.
        ``qkv [B,T,(H+2G)hs]`` -> ``y [B,T,H*hs]`` (also used by the tcgen05 prefill path, which
        computes the projections itself)."""
        cfg = self.config
        B, T, _ = qkv.shape
        q, k, v = self.split_qkv(qkv)
        n = cfg.rope_n_elem
        if n > 0:
            q = torch.cat((apply_rope(q[..., :n], cos, sin), q[..., n:]), dim=-1)
            k = torch.cat((apply_rope(k[..., :n], cos, sin), k[..., n:]), dim=-1)

        if kv is not None:
            if input_pos is None:
                raise ValueError("a KV slot needs `input_pos`")
            if B != 1:
                raise NotImplementedError("cached attention runs one sample per slot (B=1)")
            k_cache, v_cache = kv  # [G, S, hs]
            k_cache.index_copy_(1, input_pos, k[0].to(k_cache.dtype))
            v_cache.index_copy_(1, input_pos, v[0].to(v_cache.dtype))
            live = int(input_pos.max().item()) + 1
            k = k_cache[:, :live].unsqueeze(0).to(q.dtype)
            v = v_cache[:, :live].unsqueeze(0).to(q.dtype)
            # rows attend to cache positions <= their own position
            mask = torch.arange(live, device=qkv.device)[None, :] <= input_pos[:, None]
            mask = mask[None, None]
            causal = False
        else:
            mask, causal = None, True

        if cfg.q_per_kv > 1:
            k = k.repeat_interleave(cfg.q_per_kv, dim=1)
            v = v.repeat_interleave(cfg.q_per_kv, dim=1)
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=causal)
        return y.transpose(1, 2).reshape(B, T, cfg.attn_out_dim)

    def forward(
        self,
        x: torch.Tensor,
        cos: torch.Tensor,
        sin: torch.Tensor,
        input_pos: Optional[torch.Tensor] = None,
        kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    ) -> torch.Tensor:
        return self.proj(self.attend_qkv(self.attn(x), cos, sin, input_pos, kv))


class GptNeoxMLP(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        self.fc = nn.Linear(config.n_embd, config.intermediate_size, bias=config.bias)
        self.proj = nn.Linear(config.intermediate_size, config.n_embd, bias=config.bias)
        self.approx = config.gelu_approximate

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(F.gelu(self.fc(x), approximate=self.approx))


class LLaMAMLP(nn.Module):
    """SwiGLU: ``proj(silu(fc_1 x) * fc_2 x)`` (model.py:796-809)."""

    def __init__(self, config: Config) -> None:
        super().__init__()
        self.fc_1 = nn.Linear(config.n_embd, config.intermediate_size, bias=config.bias)
        self.fc_2 = nn.Linear(config.n_embd, config.intermediate_size, bias=config.bias)
        self.proj = nn.Linear(config.intermediate_size, config.n_embd, bias=config.bias)

    def gate(self, a: torch.Tensor) -> torch.Tensor:
        return F.silu(a)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(self.gate(self.fc_1(x)) * self.fc_2(x))


class GemmaMLP(LLaMAMLP):
    """GeGLU variant (model.py:812-820)."""

    def __init__(self, config: Config) -> None:
        super().__init__(config)
        self.approx = config.gelu_approximate

    def gate(self, a: torch.Tensor) -> torch.Tensor:
        return F.gelu(a, approximate=self.approx)


class LLaMAMoE(nn.Module):
    """Top-k routed mixture of SwiGLU experts, all experts local (model.py:823-853)."""

    def __init__(self, config: Config) -> None:
        super().__init__()
        self.gate = nn.Linear(config.n_embd, config.n_expert, bias=False)
        self.experts = nn.ModuleList(LLaMAMLP(config) for _ in range(config.n_expert))
        self.top = config.n_expert_per_token

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        flat = x.reshape(-1, shape[-1])
        weight, chosen = torch.topk(self.gate(flat), self.top, dim=-1)
        weight = weight.softmax(dim=-1, dtype=torch.float).to(flat.dtype)
        out = torch.zeros_like(flat)
        for e, expert in enumerate(self.experts):
            tok, which = torch.where(chosen == e)
            if tok.numel():
                out.index_add_(0, tok, weight[tok, which, None] * expert(flat[tok]))
        return out.view(shape)


_MLPS = {"GptNeoxMLP": GptNeoxMLP, "LLaMAMLP": LLaMAMLP, "GemmaMLP": GemmaMLP, "LLaMAMoE": LLaMAMoE}


def build_mlp(config: Config) -> nn.Module:
    try:
        return _MLPS[config.mlp_class_name](config)
    except KeyError:
        raise ValueError(f"unknown mlp class {config.mlp_class_name!r}") from None


SUB_UNITS = ("attn", "gu", "down")  # the three residual sub-units of a sequential-residual block with a gated MLP
_PART_ALIASES = {"both": ("attn", "gu", "down"), "attn": ("attn",), "mlp": ("gu", "down"), "attn_gu": ("attn", "gu"),
                 "gu": ("gu",), "down": ("down",)}


def part_units(parts: str) -> Tuple[str, ...]:
    try:
        return _PART_ALIASES[parts]
    except KeyError:
        raise ValueError(f"parts must be one of {sorted(_PART_ALIASES)}, got {parts!r}") from None


class _PartialGatedMLP(nn.Module):
    """The gate/up half (``fc_1``, ``fc_2``) or the down half (``proj``) of a gated MLP — a pipeline boundary
    between them carries ``[x | act(fc_1 n) * fc_2 n]`` (width C + I)."""

    def __init__(self, config: Config, units: Tuple[str, ...]) -> None:
        super().__init__()
        if "gu" in units:
            self.fc_1 = nn.Linear(config.n_embd, config.intermediate_size, bias=config.bias)
            self.fc_2 = nn.Linear(config.n_embd, config.intermediate_size, bias=config.bias)
        if "down" in units:
            self.proj = nn.Linear(config.intermediate_size, config.n_embd, bias=config.bias)
        self.silu = config.mlp_class_name == "LLaMAMLP"
        self.approx = config.gelu_approximate

    def gate(self, a: torch.Tensor) -> torch.Tensor:
        return F.silu(a) if self.silu else F.gelu(a, approximate=self.approx)


class Block(nn.Module):
    """Transformer block, sequential or parallel residual (model.py:576-629)."""

    def __init__(self, config: Config, parts: str = "both") -> None:
        """``parts``: "both" (a whole block) or a contiguous run of the sub-units of a sequential-residual block:
        "attn" (``norm_1`` + attention + residual), "gu" (``norm_2`` + the gate/up projections of a gated MLP),
        "down" (the MLP's output projection + residual); "mlp" = gu+down, "attn_gu" = attn+gu.  Partial blocks
        let a pipeline boundary fall *inside* a layer (finer stage balancing than the reference's whole-layer
        chunks); the owner of the attention sub-unit owns the layer's KV cache.  A boundary after "gu" carries
        ``[x | h]`` (width ``n_embd + intermediate_size``): the residual stream and the gated activations."""
        super().__init__()
        if not config.parallel_residual and config.shared_attention_norm:
            raise NotImplementedError("sequential residual with a shared attention norm")
        units = part_units(parts)
        if parts != "both" and config.parallel_residual:
            raise ValueError("a parallel-residual block cannot be split between stages")
        self.config = config
        self.parts = parts
        self.units = units
        self.has_attn, self.has_gu, self.has_down = "attn" in units, "gu" in units, "down" in units
        self.has_mlp = self.has_gu and self.has_down
        if self.has_attn:
            self.norm_1 = build_norm(config)
            self.attn = CausalSelfAttention(config)
        if self.has_gu:
            self.norm_2 = None if config.shared_attention_norm else build_norm(config)
        if self.has_mlp:
            self.mlp = build_mlp(config)
        elif self.has_gu or self.has_down:
            if config.mlp_class_name not in ("LLaMAMLP", "GemmaMLP"):
                raise ValueError("only gated MLPs can be cut between their up and down projections")
            self.mlp = _PartialGatedMLP(config, units)

    def forward(
        self,
        x: torch.Tensor,
        cos: torch.Tensor,
        sin: torch.Tensor,
        input_pos: Optional[torch.Tensor] = None,
        kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    ) -> torch.Tensor:
        if self.parts == "both":
            h = self.norm_1(x)
            a = self.attn(h, cos, sin, input_pos, kv)
            if self.config.parallel_residual:
                h2 = h if self.norm_2 is None else self.norm_2(x)
                return self.mlp(h2) + a + x
            x = x + a
            return x + self.mlp(self.norm_2(x))
        C = self.config.n_embd
        g = None
        if self.units[0] == "down":  # the message is [x | h]
            x, g = x[..., :C], x[..., C:]
        if self.has_attn:
            x = x + self.attn(self.norm_1(x), cos, sin, input_pos, kv)
        if self.has_mlp:
            return x + self.mlp(self.norm_2(x))
        if self.has_gu:
            n2 = self.norm_2(x)
            return torch.cat((x, self.mlp.gate(self.mlp.fc_1(n2)) * self.mlp.fc_2(n2)), dim=-1)
        if self.has_down:
            return x + self.mlp.proj(g)
        return x


# =============================================================================================
# shared trunk logic (also used by the pipeline-stage modules in models/stage.py)
# =============================================================================================
class RopeMixin:
    """max_seq_length handling + RoPE tables (reference model.py:299-327, submodels.py:45-84)."""

    config: Config

    @property
    def max_seq_length(self) -> int:
        return self._max_seq_length

    @max_seq_length.setter
    def max_seq_length(self, value: int) -> None:
        if value > self.config.block_size:
            raise ValueError(f"Cannot attend to {value}, block size is only {self.config.block_size}")
        self._max_seq_length = int(value)
        dev = self.cos.device if hasattr(self, "cos") else torch.device("cpu")
        if dev.type == "meta":  # built under `torch.device("meta")`: tables are real, on cpu
            dev = torch.device("cpu")
        cos, sin = build_rope_cache(
            self._max_seq_length, self.config.rope_n_elem, device=dev,
            base=self.config.rope_base, condense_ratio=self.config.rope_condense_ratio,
        )
        if hasattr(self, "cos"):
            self.cos, self.sin = cos, sin
        else:
            self.register_buffer("cos", cos, persistent=False)
            self.register_buffer("sin", sin, persistent=False)

    def rope_for(self, T: int, input_pos: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        if input_pos is None:
            return self.cos[:T], self.sin[:T]
        return self.cos.index_select(0, input_pos), self.sin.index_select(0, input_pos)


def run_blocks(
    blocks: Sequence[Block],
    x: torch.Tensor,
    cos: torch.Tensor,
    sin: torch.Tensor,
    input_pos: Optional[torch.Tensor],
    kv_pool: Optional[KVPool],
    slot: int,
) -> torch.Tensor:
    for li, blk in enumerate(blocks):
        kv = kv_pool.layer(li, slot) if kv_pool is not None else None
        x = blk(x, cos, sin, input_pos, kv)
    return x


class GPT(nn.Module, RopeMixin):
    """Full single-device model (reference model.py:276-573)."""

    def __init__(self, config: Config) -> None:
        super().__init__()
        assert config.padded_vocab_size is not None
        self.config = config
        self.lm_head = nn.Linear(config.n_embd, config.padded_vocab_size, bias=config.lm_head_bias)
        parts = dict(
            wte=nn.Embedding(config.padded_vocab_size, config.n_embd),
            h=nn.ModuleList(Block(config) for _ in range(config.n_layer)),
            ln_f=build_norm(config),
        )
        if config.pos_embedding == "learned":
            parts["wpe"] = nn.Embedding(config.block_size, config.n_embd)
        self.transformer = nn.ModuleDict(parts)
        if config.tie_embeddings:
            self.lm_head.weight = self.transformer.wte.weight
        self.max_seq_length = config.block_size
        self.kv_pool: Optional[KVPool] = None

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "GPT":
        return cls(Config.from_name(name, **kwargs))

    # ---- init / bookkeeping ------------------------------------------------------------------
    def _init_weights(self, module: nn.Module) -> None:
        """``gpt.apply(gpt._init_weights)`` — N(0, 0.02) like model.py:333-340."""
        if isinstance(module, nn.Linear):
            nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.Embedding):
            nn.init.normal_(module.weight, mean=0.0, std=0.02)

    def get_num_params(self, non_embedding: bool = True) -> int:
        n = sum(p.numel() for p in self.parameters())
        if non_embedding and "wpe" in self.transformer:
            n -= self.transformer.wpe.weight.numel()
        return n

    def estimate_mfu(self, fwdbwd_per_iter: float, dt: float, peak_flops: float = 1.45e15) -> float:
        """Model-FLOPs utilisation; default peak = sustained bf16 of MEASURED_PEAKS.json
        (the reference divides by the A100's 312 TF, model.py:348-368)."""
        cfg = self.config
        n = self.get_num_params()
        T = cfg.block_size
        per_token = 6 * n + 12 * cfg.n_layer * cfg.n_head * cfg.head_size * T
        return per_token * T * fwdbwd_per_iter / dt / peak_flops

    def set_kv_cache(
        self,
        batch_size: int = 1,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
        n_slots: Optional[int] = None,
    ) -> None:
        """Allocate the slot pool.  ``batch_size`` is kept for API compatibility
        (model.py:423-447); slots play the role of independent B=1 caches."""
        p = next(self.parameters())
        self.kv_pool = KVPool(
            self.config, self.config.n_layer, n_slots or batch_size, self.max_seq_length,
            device=device or p.device, dtype=dtype or p.dtype,
        )

    def clear_kv_cache(self) -> None:
        self.kv_pool = None

    # ---- forward -----------------------------------------------------------------------------
    def embed(self, idx: torch.Tensor, input_pos: Optional[torch.Tensor]) -> torch.Tensor:
        x = self.transformer.wte(idx)
        if self.config.scale_embeddings:
            x = x * (self.config.n_embd ** 0.5)
        if "wpe" in self.transformer:
            pos = input_pos if input_pos is not None else torch.arange(idx.size(1), device=idx.device)
            x = x + self.transformer.wpe(pos)
        return x

    def forward(self, idx: torch.Tensor, input_pos: Optional[torch.Tensor] = None, slot: int = 0) -> torch.Tensor:
        T = idx.size(1)
        if self.max_seq_length < T:
            raise ValueError(f"Cannot forward sequence of length {T}, max seq length is only {self.max_seq_length}.")
        if input_pos is not None and self.kv_pool is None:
            raise TypeError("You need to call `gpt.set_kv_cache()`")
        cos, sin = self.rope_for(T, input_pos)
        x = self.embed(idx, input_pos)
        x = run_blocks(self.transformer.h, x, cos, sin, input_pos,
                       self.kv_pool if input_pos is not None else None, slot)
        return self.lm_head(self.transformer.ln_f(x))

    # ---- generation --------------------------------------------------------------------------
    def next_token(self, x: torch.Tensor, input_pos: torch.Tensor, slot: int = 0, **kw: Any) -> torch.Tensor:
        return sample(self(x, input_pos, slot=slot), **kw).to(dtype=x.dtype).view(1, -1)

    @torch.inference_mode()
    def generate(
        self,
        prompt: torch.Tensor,
        max_returned_tokens: int,
        *,
        temperature: float = 1.0,
        top_k: Optional[int] = None,
        top_p: float = 1.0,
        tok_time: Optional[List[Tuple[int, float]]] = None,
        slot: int = 0,
    ) -> torch.Tensor:
        """Prefill on the prompt ``(T,)`` then decode one token at a time up to a total of
        ``max_returned_tokens``; returns ``(1, max_returned_tokens)`` (model.py:461-524)."""
        T = prompt.size(0)
        if max_returned_tokens <= T:
            raise ValueError("max_returned_tokens must exceed the prompt length")
        if self.max_seq_length < max_returned_tokens - 1:
            raise NotImplementedError(f"max_seq_length {self.max_seq_length} needs to be >= {max_returned_tokens - 1}")
        if self.kv_pool is None:
            self.set_kv_cache(1)
        input_pos = torch.arange(0, T, device=prompt.device)
        tokens = token = prompt.view(1, -1)
        t0 = time.time()
        for t in range(1, max_returned_tokens - T + 1):
            if tok_time is not None:
                tok_time.append((t - 1, time.time() - t0))
            token = self.next_token(token.view(1, -1), input_pos, slot=slot,
                                    temperature=temperature, top_k=top_k, top_p=top_p)
            tokens = torch.cat((tokens, token), dim=1)
            input_pos = input_pos[-1:] + 1
        return tokens

    @torch.inference_mode()
    def generate_chat(
        self,
        prompt: torch.Tensor,
        max_returned_tokens: int,
        *,
        temperature: float = 1.0,
        top_k: Optional[int] = None,
        top_p: float = 1.0,
        stop_tokens: Tuple[List[int], ...] = (),
        slot: int = 0,
    ) -> Iterator[torch.Tensor]:
        """Streaming generation with the reference's buffering rule (model.py:526-573): tokens are released in chunks
        of the longest stop sequence's length, generation ends — dropping the pending chunk — when a stop sequence
        completes.  Yields token tensors, prompt excluded."""
        T = prompt.size(0)
        if self.max_seq_length < max_returned_tokens - 1:
            raise NotImplementedError(f"max_seq_length {self.max_seq_length} needs to be >= {max_returned_tokens - 1}")
        if self.kv_pool is None:
            self.set_kv_cache(1)
        input_pos = torch.arange(0, T, device=prompt.device)
        token = prompt.view(1, -1)
        produced: List[torch.Tensor] = []
        emitted = 0
        hold = max((len(s) for s in stop_tokens), default=1)
        for t in range(1, max_returned_tokens - T + 1):
            token = self.next_token(token.view(1, -1), input_pos, slot=slot,
                                    temperature=temperature, top_k=top_k, top_p=top_p)
            produced.append(token)
            ids = [int(x) for x in produced[-hold:]]
            if any(len(s) <= len(produced) and ids[-len(s):] == list(s) for s in stop_tokens):
                return
            if t - emitted >= hold:
                yield from produced[emitted:t]
                emitted = t
            input_pos = input_pos[-1:] + 1
        yield from produced[emitted:]
