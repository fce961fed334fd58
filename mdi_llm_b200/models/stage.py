"""Pipeline-stage modules: the local slice of the model that one node (GPU) owns.

Parity: reference ``src/sub/submodels.py`` — ``StarterNode`` (:132-220: ``wte`` + first
blocks **and** ``ln_f`` + ``lm_head``; ``forward(first_pass=True)`` runs embedding+blocks,
``first_pass=False`` runs the output head) and ``SecondaryNode`` (:223-282: blocks only).
State-dict keys match the chunk files ``model_starter.pth`` / ``model_secondary{i}.pth``
(locally re-indexed ``transformer.h.{0..n-1}``).

These eager modules are what runs on CPU, what the GPU tests use as oracle, and what holds
the weights that the CUDA stage executor (``parallel/engine.py``) reads in place.
Each stage owns a :class:`KVPool` with one slot per in-flight sample; the per-sample cache
"rotation" of the reference (gptserver.py:975-978) becomes passing ``slot=``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .config import Config
from .gpt import Block, KVPool, RopeMixin, build_norm, run_blocks

__all__ = ["StageModule", "NodePrototype", "StarterNode", "SecondaryNode", "FinisherNode", "build_stage"]


class StageModule(nn.Module, RopeMixin):
    """Common behaviour of a pipeline stage (reference ``NodePrototype``, submodels.py:34-127)."""

    role = "stage"

    def __init__(self, config: Config, n_transf_layers: int, **kwargs: Any) -> None:
        super().__init__()
        self.config = config
        self.n_local_layers = int(n_transf_layers)
        # sub-layer boundaries (models/partition.py: plan_half_units / plan_third_units): the first local block
        # may start at its MLP ("mlp") or at the MLP's down projection ("down"), the last one may end after its
        # attention ("attn") or after the gate/up projections ("attn_gu"); "both" = whole block
        self.first_parts = str(kwargs.get("first_parts") or ("mlp" if kwargs.get("first_mlp_only") else "both"))
        self.last_parts = str(kwargs.get("last_parts") or ("attn" if kwargs.get("last_attn_only") else "both"))
        self.first_mlp_only = self.first_parts == "mlp"
        self.last_attn_only = self.last_parts == "attn"
        self.verb = bool(kwargs.get("verb", False))
        self.params_init = False
        self.kv_pool: Optional[KVPool] = None
        self.transformer = nn.ModuleDict()

    # -- cache management --------------------------------------------------------------------
    def set_kv_cache(
        self,
        n_slots: int,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
    ) -> KVPool:
        p = next(self.parameters())
        self.kv_pool = KVPool(
            self.config, self.n_local_layers, n_slots, self.max_seq_length,
            device=device or p.device, dtype=dtype or p.dtype,
        )
        return self.kv_pool

    def ensure_slots(self, n_slots: int) -> KVPool:
        """Grow the pool lazily — secondaries learn about samples as they arrive
        (gptserver.py:1083-1088)."""
        if self.kv_pool is None:
            return self.set_kv_cache(n_slots)
        if self.kv_pool.n_slots < n_slots:
            old = self.kv_pool
            new = self.set_kv_cache(n_slots, device=old.data.device, dtype=old.data.dtype)
            new.data[:, : old.n_slots].copy_(old.data)
        return self.kv_pool

    def clear_kv_cache(self) -> None:
        self.kv_pool = None

    def load_weights(self, params: Dict[str, Any], **kwargs: Any) -> int:
        """Fill this (possibly meta) module from a chunk state dict (submodels.py:161-168)."""
        from ..utils.checkpoint import init_from_state_dict

        init_from_state_dict(self, params)
        self.params_init = True
        return 1

    def _make_blocks(self) -> nn.ModuleList:
        from .gpt import SUB_UNITS, part_units

        n = self.n_local_layers
        names = {("attn", "gu", "down"): "both", ("attn",): "attn", ("gu", "down"): "mlp", ("attn", "gu"): "attn_gu",
                 ("gu",): "gu", ("down",): "down"}

        def parts(i: int) -> str:
            units = set(SUB_UNITS)
            if i == 0:
                units &= set(part_units(self.first_parts))
            if i == n - 1:
                units &= set(part_units(self.last_parts))
            key = tuple(u for u in SUB_UNITS if u in units)
            if key not in names:
                raise ValueError(f"block {i}: first_parts={self.first_parts!r} and last_parts={self.last_parts!r} leave no "
                                 "contiguous sub-units")
            return names[key]

        return nn.ModuleList(Block(self.config, parts(i)) for i in range(n))

    @property
    def in_width(self) -> int:
        """Width of the incoming hidden message: ``[x | h]`` when the stage starts at a down projection."""
        c = self.config
        blocks = self.transformer.h
        return c.n_embd + c.intermediate_size if len(blocks) and blocks[0].units[0] == "down" else c.n_embd

    @property
    def out_width(self) -> int:
        c = self.config
        blocks = self.transformer.h
        return c.n_embd + c.intermediate_size if len(blocks) and blocks[-1].units[-1] == "gu" else c.n_embd

    def _blocks(self, x: torch.Tensor, input_pos: Optional[torch.Tensor], slot: int) -> torch.Tensor:
        T = x.size(1)
        if self.max_seq_length < T:
            raise ValueError(f"Cannot forward sequence of length {T}, max seq length is only {self.max_seq_length}.")
        if input_pos is not None and self.kv_pool is None:
            raise TypeError("You need to call `set_kv_cache()`")
        cos, sin = self.rope_for(T, input_pos)
        return run_blocks(self.transformer.h, x, cos, sin, input_pos,
                          self.kv_pool if input_pos is not None else None, slot)


NodePrototype = StageModule  # the reference's name for the common base (submodels.py:34)


class StarterNode(StageModule):
    role = "starter"

    def __init__(self, config: Config, n_transf_layers: int, *, with_head: bool = True, **kwargs: Any) -> None:
        """``with_head=False`` builds the first-generation starter (old/nanoGPT/sub/model_dist.py:90-144):
        embeddings + blocks only, the output head lives on a :class:`FinisherNode`."""
        super().__init__(config, n_transf_layers, **kwargs)
        self.with_head = bool(with_head)
        parts = dict(
            wte=nn.Embedding(config.padded_vocab_size, config.n_embd),
            h=self._make_blocks(),
        )
        if self.with_head:
            parts["ln_f"] = build_norm(config)
        if config.pos_embedding == "learned":
            parts["wpe"] = nn.Embedding(config.block_size, config.n_embd)
        self.transformer = nn.ModuleDict(parts)
        if self.with_head:
            self.lm_head = nn.Linear(config.n_embd, config.padded_vocab_size, bias=config.lm_head_bias)
            if config.tie_embeddings:
                self.lm_head.weight = self.transformer.wte.weight
        self.max_seq_length = config.block_size

    def embed(self, idx: torch.Tensor, input_pos: Optional[torch.Tensor]) -> torch.Tensor:
        x = self.transformer.wte(idx)
        if self.config.scale_embeddings:
            x = x * (self.config.n_embd ** 0.5)
        if "wpe" in self.transformer:
            pos = input_pos if input_pos is not None else torch.arange(idx.size(1), device=idx.device)
            x = x + self.transformer.wpe(pos)
        return x

    def head(self, x: torch.Tensor) -> torch.Tensor:
        """``ln_f`` + ``lm_head``; only the last row is needed for sampling, so callers may pass
        ``x[:, -1:]`` (the reference computes all T rows, submodels.py:219-220)."""
        if not self.with_head:
            raise RuntimeError("this starter has no output head (finisher topology)")
        return self.lm_head(self.transformer.ln_f(x))

    def forward(
        self,
        idx: torch.Tensor,
        input_pos: Optional[torch.Tensor] = None,
        *,
        first_pass: bool = True,
        slot: int = 0,
    ) -> torch.Tensor:
        if not first_pass:
            return self.head(idx)
        B, T = idx.shape
        if T > self.config.block_size:
            raise ValueError(f"Cannot forward sequence of length {T}, block size is {self.config.block_size}")
        if B > 1 and input_pos is not None:
            raise NotImplementedError("one sample per message (B=1) on the cached path")
        return self._blocks(self.embed(idx, input_pos), input_pos, slot)


class SecondaryNode(StageModule):
    role = "secondary"

    def __init__(self, config: Config, n_transf_layers: int, **kwargs: Any) -> None:
        super().__init__(config, n_transf_layers, **kwargs)
        self.transformer = nn.ModuleDict(dict(h=self._make_blocks()))
        self.max_seq_length = config.block_size

    def forward(self, x: torch.Tensor, input_pos: Optional[torch.Tensor] = None, *, slot: int = 0) -> torch.Tensor:
        return self._blocks(x, input_pos, slot)


class FinisherNode(StageModule):
    """Last node of the first-generation open chain: its blocks, then ``ln_f`` + ``lm_head``; returns
    the logits of the last position, which travel back to the starter for sampling
    (old/nanoGPT/sub/model_dist.py:181-221; the ring generation moved the head to the starter)."""

    role = "finisher"

    def __init__(self, config: Config, n_transf_layers: int, **kwargs: Any) -> None:
        super().__init__(config, n_transf_layers, **kwargs)
        self.transformer = nn.ModuleDict(dict(
            h=self._make_blocks(), ln_f=build_norm(config)))
        self.lm_head = nn.Linear(config.n_embd, config.padded_vocab_size, bias=config.lm_head_bias)
        self.max_seq_length = config.block_size

    def forward(self, x: torch.Tensor, input_pos: Optional[torch.Tensor] = None, *, slot: int = 0) -> torch.Tensor:
        x = self._blocks(x, input_pos, slot)
        return self.lm_head(self.transformer.ln_f(x[:, -1:]))


def build_stage(config: Config, role: str, n_layers: int, *, meta: bool = False, **kw: Any) -> StageModule:
    """Instantiate the stage for ``role`` ("starter" | "secondary[:i]" | "intermediate[:i]" | "finisher"),
    optionally on the meta device so that loading a chunk does not double the memory
    (gptserver.py:657-664)."""
    cls = StarterNode if role.startswith("starter") else FinisherNode if role.startswith("finisher") else SecondaryNode
    if meta:
        with torch.device("meta"):
            return cls(config, n_layers, **kw)
    return cls(config, n_layers, **kw)
