"""Exact re-parametrisations that bring a checkpoint into the fused engine's configuration space.

The decode kernels are instantiated for head sizes 64 / 128 / 256 and 1 / 2 / 4 / 8 query heads per KV head
(``parallel/engine.py::engine_supports``).  Two model families of the reference's registry fall outside:

* **odd head sizes** (Phi-2: 80).  ``pad_head_size`` appends zero rows to every q / k / v head and zero columns to the
  output projection.  Zero k dimensions add nothing to q·k, zero v dimensions produce zeros that meet zero columns of
  ``proj``; the softmax scale ``1/sqrt(head_size)`` changes with the padded size, so the q rows (and bias) are scaled by
  ``sqrt(new / old)`` — RoPE is linear, the rotated dimensions stay the first ``rope_n_elem`` of the head, and
  ``rotary_percentage`` is re-expressed against the new size.
* **wide GQA / MQA groups** (Falcon-7B: 71 query heads on one KV head, Falcon-40B: 16, Gemma-2B: 8 at head size 256).
  ``expand_kv_groups`` splits a group into narrower ones (the widest supported divisor: 16 -> 8, 71 -> 1), each with
  its own copy of the group's k / v rows: the same function with a larger KV cache.

Both are *exact* (same logits up to floating-point summation order): ``tests/test_fit_engine.py`` checks the
transformed eager model against the original, prefill and cached decoding.  Costs: Phi-2's attention matrices and KV
slots grow by 128/80 (+17 % of a layer's bytes); Falcon-7B's QKV matrix grows 2.9x (+20 % of a layer's bytes) and its
KV slots 71x (37 MB per layer per 2048-token sample) — the price of running on the streaming kernels instead of eager
PyTorch.  ``prepare_model --fit-engine`` writes the converted checkpoint next to the original.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, Dict, List, Optional, Tuple

import torch

from ..models.config import Config

SUPPORTED_HEAD_SIZES = (64, 128, 256)
SUPPORTED_Q_PER_KV = (1, 2, 4, 8)


def _blocks(sd: Dict[str, Any]) -> List[str]:
    return sorted({k.rsplit(".attn.attn.weight", 1)[0] for k in sd if k.endswith(".attn.attn.weight")})


def pad_head_size(config: Config, sd: Optional[Dict[str, torch.Tensor]], new_head_size: int) -> Tuple[Config, Optional[Dict[str, torch.Tensor]]]:
    hs, G, qpk, H, C = config.head_size, config.n_query_groups, config.q_per_kv, config.n_head, config.n_embd
    if new_head_size < hs:
        raise ValueError(f"cannot shrink heads ({hs} -> {new_head_size})")
    # keep exactly the same rotated dimensions: int((n + 0.5) / new * new) == n whatever the rounding of the division
    new_cfg = dataclasses.replace(config, head_size=new_head_size, rotary_percentage=(config.rope_n_elem + 0.5) / new_head_size
                                  if config.rope_n_elem else 0.0)
    assert new_cfg.rope_n_elem == config.rope_n_elem
    if sd is None or new_head_size == hs:
        return new_cfg, sd
    s = math.sqrt(new_head_size / hs)  # softmax scale 1/sqrt(head_size) is taken from the padded size
    out = dict(sd)
    for b in _blocks(sd):
        w = sd[f"{b}.attn.attn.weight"]
        wp = w.new_zeros(G, qpk + 2, new_head_size, C)
        wp[:, :, :hs] = w.view(G, qpk + 2, hs, C)
        wp[:, :qpk] = (wp[:, :qpk].float() * s).to(w.dtype)
        out[f"{b}.attn.attn.weight"] = wp.reshape((qpk + 2) * G * new_head_size, C)
        if f"{b}.attn.attn.bias" in sd:
            bias = sd[f"{b}.attn.attn.bias"]
            bp = bias.new_zeros(G, qpk + 2, new_head_size)
            bp[:, :, :hs] = bias.view(G, qpk + 2, hs)
            bp[:, :qpk] = (bp[:, :qpk].float() * s).to(bias.dtype)
            out[f"{b}.attn.attn.bias"] = bp.reshape(-1)
        p = sd[f"{b}.attn.proj.weight"]
        pp = p.new_zeros(C, H, new_head_size)
        pp[:, :, :hs] = p.view(C, H, hs)
        out[f"{b}.attn.proj.weight"] = pp.reshape(C, H * new_head_size)
    return new_cfg, out


def expand_kv_groups(config: Config, sd: Optional[Dict[str, torch.Tensor]], q_per_kv: int = 1) -> Tuple[Config, Optional[Dict[str, torch.Tensor]]]:
    """Split every KV group into groups of ``q_per_kv`` query heads (a divisor of the current width), each with its own
    copy of the group's k / v rows.  ``q_per_kv = 1``: one KV head per query head."""
    hs, G, qpk, H, C = config.head_size, config.n_query_groups, config.q_per_kv, config.n_head, config.n_embd
    if qpk % q_per_kv:
        raise ValueError(f"{q_per_kv} does not divide the {qpk} query heads of a KV group")
    r = qpk // q_per_kv  # new groups per old group
    new_cfg = dataclasses.replace(config, n_query_groups=G * r)
    if sd is None or r == 1:
        return new_cfg, sd
    out = dict(sd)

    def regroup(t: torch.Tensor, tail: Tuple[int, ...]) -> torch.Tensor:
        t = t.view(G, qpk + 2, hs, *tail)
        q = t[:, :qpk].reshape(G, r, q_per_kv, hs, *tail)
        kv = t[:, qpk:].unsqueeze(1).expand(G, r, 2, hs, *tail)
        return torch.cat((q, kv), dim=2).reshape(G * r * (q_per_kv + 2) * hs, *tail).contiguous()

    for b in _blocks(sd):
        out[f"{b}.attn.attn.weight"] = regroup(sd[f"{b}.attn.attn.weight"], (C,))
        if f"{b}.attn.attn.bias" in sd:
            out[f"{b}.attn.attn.bias"] = regroup(sd[f"{b}.attn.attn.bias"], ())
    return new_cfg, out


def fit_engine(config: Config, sd: Optional[Dict[str, torch.Tensor]] = None) -> Tuple[Config, Optional[Dict[str, torch.Tensor]], List[str]]:
    """Apply whichever of the two re-parametrisations the fused engine needs.  Returns ``(config, state dict, notes)``;
    raises ``ValueError`` when the architecture is outside the engine for another reason (nothing is converted then)."""
    from ..parallel.engine import engine_supports

    notes: List[str] = []
    cfg = config

    def narrow(limit: int) -> None:
        nonlocal cfg, sd
        target = next(q for q in (8, 4, 2, 1) if q <= limit and cfg.q_per_kv % q == 0)
        notes.append(f"{cfg.q_per_kv} query heads per KV head -> {target} ({cfg.n_query_groups} -> "
                     f"{cfg.n_query_groups * cfg.q_per_kv // target} KV groups, k / v rows repeated)")
        cfg, sd = expand_kv_groups(cfg, sd, target)

    new_hs = cfg.head_size if cfg.head_size in SUPPORTED_HEAD_SIZES else next((h for h in SUPPORTED_HEAD_SIZES if h >= cfg.head_size), None)
    if new_hs is None:
        raise ValueError(f"head size {cfg.head_size} is larger than any kernel instantiation")
    limit = 2 if new_hs > 128 else 8  # the head-size-256 attention kernel serves at most 2 query heads per KV head
    if cfg.q_per_kv not in SUPPORTED_Q_PER_KV or cfg.q_per_kv > limit:
        narrow(limit)
    if new_hs != cfg.head_size:
        notes.append(f"head size {cfg.head_size} -> {new_hs} (zero-padded, q rows scaled by sqrt({new_hs}/{cfg.head_size}))")
        cfg, sd = pad_head_size(cfg, sd, new_hs)
    if not engine_supports(cfg, torch.bfloat16):
        raise ValueError(f"{config.name}: outside the fused engine for a reason no re-parametrisation removes "
                         f"(norm {cfg.norm_class_name}, mlp {cfg.mlp_class_name}, positions {cfg.pos_embedding})")
    return cfg, sd, notes
