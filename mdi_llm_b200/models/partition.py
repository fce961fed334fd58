"""Layer partitioning of a litGPT model over pipeline stages and chunk files.

Parity: reference partition table ``N_LAYERS_NODES`` (``src/sub/config.py:56-98``),
``split_parameters`` (``src/sub/utils/utils.py:241-385``) and ``split_and_store``
(``utils.py:388-438``): the starter owns ``wte`` + the first blocks + ``ln_f`` + ``lm_head``;
secondary ``i`` owns the next contiguous blocks, re-indexed from 0; files are written to
``<ckpt>/chunks/<N>nodes/model_starter.pth`` and ``model_secondary{i}.pth``.

New: the reference raises ``KeyError`` for any topology missing from its table (e.g. 8 nodes).
Here :func:`plan_layers` falls back to a *balanced planner* that accounts for the bytes the
starter additionally streams per token for ``ln_f`` + ``lm_head`` (≈2.4 Llama-3 blocks) and
may give secondaries different layer counts — per-stage layer lists that the reference
cannot express.  For topologies the reference supports the table wins, so chunk files are
interchangeable.
"""
from __future__ import annotations

import gc
import os
import warnings
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from .config import Config

__all__ = [
    "N_LAYERS_NODES", "plan_layers", "balanced_plan", "layer_ranges", "split_parameters",
    "split_and_store", "merge_chunks", "count_transformer_blocks", "chunk_dir", "chunk_file",
    "HalfStage", "plan_half_units", "half_stages", "split_parameters_half", "decode_unit_costs",
    "stage_shape_from_state_dict", "stage_specs", "plan_third_units", "third_stages", "split_parameters_units",
    "decode_third_costs",
]

# n_nodes -> n_layer -> (starter layers, layers per secondary).  Data of config.py:56-98.
_TABLE = {
    1: {n: (n, 0) for n in (5, 7, 9, 12, 22, 24, 32, 36, 48)},
    2: {5: (2, 3), 7: (3, 4), 9: (4, 5), 12: (5, 7), 22: (10, 12), 24: (10, 14), 32: (14, 18),
        36: (16, 20), 48: (22, 26)},
    3: {5: (1, 2), 7: (1, 3), 9: (1, 4), 12: (2, 5), 22: (6, 8), 24: (4, 10), 32: (8, 12),
        36: (10, 13), 48: (14, 17)},
    4: {22: (4, 6), 32: (5, 9)},
    5: {22: (2, 5), 32: (4, 7)},
}
N_LAYERS_NODES: Dict[int, Dict[int, Dict[str, int]]] = {
    n: {
        L: ({"N_LAYERS_START": s} if n == 1 else {"N_LAYERS_START": s, "N_LAYERS_SECONDARY": r})
        for L, (s, r) in per.items()
    }
    for n, per in _TABLE.items()
}


def balanced_plan(n_nodes: int, n_layer: int, head_cost_blocks: float = 0.0) -> List[int]:
    """Layers per stage ``[starter, sec0, ...]`` minimising the slowest stage when the starter
    carries an extra ``head_cost_blocks`` blocks-worth of work (``ln_f``+``lm_head``+sampling).

    Every stage gets at least one block.  Ties are broken towards the *later* stages being
    lighter, so the wrap-around hop finds the starter ready.
    """
    if n_nodes < 1:
        raise ValueError("n_nodes must be >= 1")
    if n_nodes == 1:
        return [n_layer]
    if n_layer < n_nodes:
        raise ValueError(f"cannot split {n_layer} layers over {n_nodes} nodes")
    best: Optional[List[int]] = None
    best_cost = float("inf")
    for s in range(1, n_layer - (n_nodes - 1) + 1):
        rest = n_layer - s
        q, r = divmod(rest, n_nodes - 1)
        if q == 0:
            continue
        secs = [q + 1] * r + [q] * (n_nodes - 1 - r)
        cost = max(s + head_cost_blocks, max(secs))
        # prefer lower bottleneck, then a heavier starter (fewer hops of imbalance downstream)
        if cost < best_cost - 1e-9 or (abs(cost - best_cost) <= 1e-9 and best is not None and s > best[0]):
            best, best_cost = [s] + secs, cost
    assert best is not None
    return best


def plan_layers(
    n_nodes: int,
    n_layer: int,
    config: Optional[Config] = None,
    policy: str = "auto",
) -> List[int]:
    """Layers per stage.  ``policy``: ``"table"`` (reference table only, ``KeyError`` when
    missing), ``"balanced"`` (planner only) or ``"auto"`` (table if present else planner)."""
    if policy not in ("auto", "table", "balanced"):
        raise ValueError(f"unknown partition policy {policy!r}")
    if policy in ("auto", "table"):
        entry = N_LAYERS_NODES.get(n_nodes, {}).get(n_layer)
        if entry is not None:
            s = entry["N_LAYERS_START"]
            return [s] + [entry.get("N_LAYERS_SECONDARY", 0)] * (n_nodes - 1)
        if policy == "table":
            raise KeyError(f"no reference partition for {n_nodes} nodes x {n_layer} layers")
    head = 0.0
    if config is not None:
        head = config.head_param_count() / max(1, config.block_param_count())
    return balanced_plan(n_nodes, n_layer, head)


# ---- half-layer partitions ---------------------------------------------------------------------------
# A sequential-residual block is two independent residual sub-layers: attention (norm_1 + attn) and MLP
# (norm_2 + mlp).  Cutting the pipeline between them doubles the planner's granularity: 8 stages x
# Llama-3-8B goes from a 5-layer bottleneck stage to ~4.4 layer-equivalents (the reference's chunks
# are whole layers, utils.py:241-385).  Unit u = 2*layer + (0: attention | 1: MLP).
from dataclasses import dataclass


@dataclass(frozen=True)
class HalfStage:
    lo_unit: int            # first half-unit owned (inclusive)
    hi_unit: int            # one past the last half-unit
    first_mlp_only: bool    # first local block holds only its MLP half
    last_attn_only: bool    # last local block holds only its attention half

    @property
    def lo_layer(self) -> int:
        return self.lo_unit // 2

    @property
    def hi_layer(self) -> int:  # one past the last (possibly partial) global layer
        return (self.hi_unit + 1) // 2

    @property
    def n_blocks(self) -> int:
        return self.hi_layer - self.lo_layer


def half_stages(units_per_stage: Sequence[int]) -> List[HalfStage]:
    out, u = [], 0
    for n in units_per_stage:
        if n < 1:
            raise ValueError("every stage needs at least one half-layer unit")
        out.append(HalfStage(u, u + n, first_mlp_only=u % 2 == 1, last_attn_only=(u + n) % 2 == 1))
        u += n
    return out


def decode_unit_costs(config: Config, bytes_per_param: float = 2.0, eff_tbps: float = 6.4, kernel_us: float = 1.5,
                      attn_kernel_us: float = 6.0, sampler_us: float = 35.0) -> Tuple[float, float, float]:
    """Per-token decode cost model in microseconds -> ``(attention unit, MLP unit, output head)``.
    Decode is weight-streaming bound: bytes over the achieved HBM rate plus a per-launch term; defaults
    are the measured B200 values (profiles/README.md)."""
    C = config.n_embd
    bw = eff_tbps * 1e6  # bytes per microsecond
    attn_b = (config.qkv_size + config.attn_out_dim) * C * bytes_per_param
    mlp_mats = 3 if config.mlp_class_name in ("LLaMAMLP", "GemmaMLP", "LLaMAMoE") else 2
    mlp_b = mlp_mats * C * config.intermediate_size * bytes_per_param
    mlp_kernels = 2
    if config.mlp_class_name == "LLaMAMoE":  # a token streams only its chosen experts: router + (gate/up, down) per expert
        mlp_b = config.n_expert_per_token * mlp_b + config.n_expert * C * bytes_per_param
        mlp_kernels = 1 + 2 * config.n_expert_per_token
    head_b = config.padded_vocab_size * C * bytes_per_param
    return (attn_b / bw + 2 * kernel_us + attn_kernel_us, mlp_b / bw + mlp_kernels * kernel_us,
            head_b / bw + kernel_us + sampler_us)


def plan_half_units(n_nodes: int, config: Config, costs: Optional[Tuple[float, float, float]] = None) -> List[int]:
    """Half-units per stage (sum = 2 * n_layer) minimising the slowest stage of the decode ring; stage 0
    additionally carries the output head.  Exact DP over contiguous partitions."""
    if config.parallel_residual:
        raise ValueError("half-layer partitions need sequential-residual blocks")
    U = 2 * config.n_layer
    if n_nodes < 1 or U < n_nodes:
        raise ValueError(f"cannot split {U} half-layers over {n_nodes} nodes")
    ca, cm, ch = costs if costs is not None else decode_unit_costs(config)
    pre = [0.0]
    for u in range(U):
        pre.append(pre[-1] + (ca if u % 2 == 0 else cm))
    seg = lambda a, b, first: pre[b] - pre[a] + (ch if first else 0.0)  # noqa: E731
    INF = float("inf")
    # best[k][u] = minimal bottleneck covering units [0,u) with k stages
    best = [[INF] * (U + 1) for _ in range(n_nodes + 1)]
    arg = [[0] * (U + 1) for _ in range(n_nodes + 1)]
    for u in range(1, U + 1):
        best[1][u] = seg(0, u, True)
    for k in range(2, n_nodes + 1):
        for u in range(k, U + 1):
            for a in range(k - 1, u):
                c = max(best[k - 1][a], seg(a, u, False))
                if c < best[k][u] - 1e-9:
                    best[k][u], arg[k][u] = c, a
    cuts, u = [], U
    for k in range(n_nodes, 1, -1):
        a = arg[k][u]
        cuts.append(u - a)
        u = a
    cuts.append(u)
    return cuts[::-1]


_ATTN_HALF = ("norm_1.", "attn.")


def split_parameters_half(model_params: Dict[str, Any], units_per_stage: Sequence[int]) -> Dict[str, Any]:
    """Like :func:`split_parameters` for a half-unit plan.  A layer cut in two appears in both
    neighbouring chunks: its ``norm_1``/``attn`` tensors as the *last* local block of one stage, its
    ``norm_2``/``mlp`` (and any other) tensors as local block 0 of the next."""
    n_layer = count_transformer_blocks(model_params)
    stages = half_stages(units_per_stage)
    if stages[-1].hi_unit != 2 * n_layer:
        raise ValueError(f"plan {list(units_per_stage)} does not cover {2 * n_layer} half-layers")
    chunks: List[Dict[str, Any]] = [{} for _ in stages]
    for k in [k for k in model_params if k.startswith("transformer.h.")]:
        _, _, li, tail = k.split(".", 3)
        unit = 2 * int(li) + (0 if tail.startswith(_ATTN_HALF) else 1)
        si = next(i for i, st in enumerate(stages) if st.lo_unit <= unit < st.hi_unit)
        chunks[si][f"transformer.h.{int(li) - stages[si].lo_layer}.{tail}"] = model_params.pop(k)
    for k in ("transformer.wte.weight", "transformer.wte.bias", "transformer.wpe.weight", "transformer.ln_f.weight",
              "transformer.ln_f.bias", "lm_head.weight", "lm_head.bias"):
        if k in model_params:
            chunks[0][k] = model_params.pop(k)
    return {"starter": chunks[0], "secondary": chunks[1:]}


# ---- third-layer partitions -------------------------------------------------------------------------------
# A gated MLP is itself two weight passes: gate/up (2 x [I, C]) and down ([C, I]).  Cutting between them makes the
# planner's unit a third of a layer: unit u = 3*layer + (0: attention | 1: gate/up | 2: down).  The boundary after
# a gate/up unit carries [x | h] (n_embd + intermediate_size values per token instead of n_embd) — still nothing
# next to the weight stream it balances: measured decode costs 30 / 34 / 19 us per unit (Llama-3-8B, B200) against
# a 200 us output head put the 8-stage bottleneck at 377 us with thirds vs 396 us with halves.
_THIRD_OF = {"norm_1": 0, "attn": 0, "norm_2": 1, "mlp.fc_1": 1, "mlp.fc_2": 1, "mlp.proj": 2}
_FIRST_PARTS = {0: "both", 1: "mlp", 2: "down"}   # by (first unit % 3)
_LAST_PARTS = {0: "both", 1: "attn", 2: "attn_gu"}  # by (one-past-last unit % 3)


def _third_of(tail: str) -> int:
    for prefix, k in _THIRD_OF.items():
        if tail.startswith(prefix + "."):
            return k
    return 1  # anything else of the MLP half travels with gate/up


def third_stages(units_per_stage: Sequence[int]) -> List[Dict[str, Any]]:
    """Stage shapes of a third-unit plan: ``n_blocks``, ``first_parts`` / ``last_parts`` (see ``Block``),
    ``layer_offset``, ``lo_unit`` / ``hi_unit``."""
    out, u = [], 0
    for n in units_per_stage:
        if n < 1:
            raise ValueError("every stage needs at least one unit")
        lo, hi = u, u + n
        lo_layer, hi_layer = lo // 3, (hi + 2) // 3
        out.append({"n_blocks": hi_layer - lo_layer, "first_parts": _FIRST_PARTS[lo % 3], "last_parts": _LAST_PARTS[hi % 3],
                    "layer_offset": lo_layer, "lo_unit": lo, "hi_unit": hi, "units": n, "unit": "third"})
        u = hi
    return out


def decode_third_costs(config: Config, bytes_per_param: float = 2.0, eff_tbps: float = 6.45, small_us: float = 2.6,
                       attn_kernel_us: float = 10.5, head_extra_us: float = 38.0) -> Tuple[float, float, float, float, float]:
    """``(attention, gate/up, down, head, per-step)`` decode costs in microseconds, fitted to the per-kernel device
    trace of a Llama-3-8B stage on a B200 at ~0.5k context (profiles/README.md): weight bytes at the achieved HBM
    rate plus a latency term per small kernel, the latency-bound attention kernel, and the fixed cost of a step."""
    C, bw = config.n_embd, eff_tbps * 1e6
    qkv = config.qkv_size * C * bytes_per_param / bw + small_us
    o = config.attn_out_dim * C * bytes_per_param / bw + small_us + 1.3
    gu = 2 * C * config.intermediate_size * bytes_per_param / bw - 1.5  # its first chunks are prefetched under the o_proj
    down = C * config.intermediate_size * bytes_per_param / bw + 1.0
    head = config.padded_vocab_size * C * bytes_per_param / bw + head_extra_us
    return qkv + attn_kernel_us + o, gu, down, head, 9.0


def _plan_units(costs: Sequence[float], n_nodes: int, head: float, fixed: float) -> List[int]:
    """Exact DP over contiguous partitions of ``costs`` minimising the slowest stage (stage 0 adds ``head``)."""
    U = len(costs)
    if n_nodes < 1 or U < n_nodes:
        raise ValueError(f"cannot split {U} units over {n_nodes} nodes")
    pre = [0.0]
    for c in costs:
        pre.append(pre[-1] + c)
    INF = float("inf")
    best = [[INF] * (U + 1) for _ in range(n_nodes + 1)]
    arg = [[0] * (U + 1) for _ in range(n_nodes + 1)]
    for u in range(1, U + 1):
        best[1][u] = pre[u] + head + fixed
    for k in range(2, n_nodes + 1):
        for u in range(k, U + 1):
            for a in range(k - 1, u):
                c = max(best[k - 1][a], pre[u] - pre[a] + fixed)
                if c < best[k][u] - 1e-9:
                    best[k][u], arg[k][u] = c, a
    cuts, u = [], U
    for k in range(n_nodes, 1, -1):
        a = arg[k][u]
        cuts.append(u - a)
        u = a
    cuts.append(u)
    return cuts[::-1]


def plan_third_units(n_nodes: int, config: Config, costs: Optional[Tuple[float, float, float, float, float]] = None) -> List[int]:
    """Third-units per stage (sum = 3 * n_layer) minimising the slowest stage of the decode ring."""
    if config.parallel_residual or config.mlp_class_name not in ("LLaMAMLP", "GemmaMLP"):
        raise ValueError("third-layer partitions need sequential-residual blocks with a gated MLP")
    ca, cg, cd, ch, fx = costs if costs is not None else decode_third_costs(config)
    return _plan_units([ca, cg, cd] * config.n_layer, n_nodes, ch, fx)


def split_parameters_units(model_params: Dict[str, Any], specs: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
    """Chunks for a third-unit plan (``specs`` from :func:`third_stages`): every tensor of a layer goes to the stage
    that owns its sub-unit, re-indexed from that stage's first (possibly partial) block."""
    n_layer = count_transformer_blocks(model_params)
    if specs[-1]["hi_unit"] != 3 * n_layer:
        raise ValueError(f"plan does not cover {3 * n_layer} third-layer units")
    chunks: List[Dict[str, Any]] = [{} for _ in specs]
    for k in [k for k in model_params if k.startswith("transformer.h.")]:
        _, _, li, tail = k.split(".", 3)
        unit = 3 * int(li) + _third_of(tail)
        si = next(i for i, sp in enumerate(specs) if sp["lo_unit"] <= unit < sp["hi_unit"])
        chunks[si][f"transformer.h.{int(li) - specs[si]['layer_offset']}.{tail}"] = model_params.pop(k)
    for k in ("transformer.wte.weight", "transformer.wte.bias", "transformer.wpe.weight", "transformer.ln_f.weight",
              "transformer.ln_f.bias", "lm_head.weight", "lm_head.bias"):
        if k in model_params:
            chunks[0][k] = model_params.pop(k)
    return {"starter": chunks[0], "secondary": chunks[1:]}


def stage_shape_from_state_dict(sd: Dict[str, Any]) -> Dict[str, Any]:
    """``{"n_blocks", "first_parts", "last_parts"}`` (+ the half-plan booleans) of a chunk, read off its keys: which
    sub-units of its first / last block are present — so chunk files fully describe sub-layer plans."""
    blocks: Dict[int, set] = {}
    for k in sd:
        if k.startswith("transformer.h."):
            _, _, li, tail = k.split(".", 3)
            blocks.setdefault(int(li), set()).add(_third_of(tail))
    n = len(blocks)
    if n == 0:
        return {"n_blocks": 0, "first_parts": "both", "last_parts": "both", "first_mlp_only": False, "last_attn_only": False}
    first, last = blocks[min(blocks)], blocks[max(blocks)]
    gated = not any(".mlp.fc." in k or ".mlp.experts." in k for k in sd)  # fc + proj (GPT-NeoX) / MoE MLPs are never cut inside
    if not gated:  # two-matrix MLPs (fc + proj) are never cut inside: "mlp present" = units 1 and 2
        first = first | ({1, 2} if first & {1, 2} else set())
        last = last | ({1, 2} if last & {1, 2} else set())
    fp = "both" if 0 in first else ("mlp" if 1 in first else "down")
    lp = "both" if 2 in last else ("attn_gu" if 1 in last else "attn")
    return {"n_blocks": n, "first_parts": fp, "last_parts": lp, "first_mlp_only": fp == "mlp", "last_attn_only": lp == "attn"}


def stage_specs(n_nodes: int, config: Config, policy: str = "auto") -> List[Dict[str, Any]]:
    """Per-stage shape for a partition policy: ``"third"`` / ``"half"`` = sub-layer units (boundaries may fall
    between a layer's attention, gate/up and down passes), else whole layers via :func:`plan_layers`.  Every entry
    carries ``n_blocks``, ``first_parts``, ``last_parts``, ``layer_offset`` and ``layers`` (layer-equivalents)."""
    sub_ok = n_nodes > 1 and not config.parallel_residual
    if policy == "third" and sub_ok and config.mlp_class_name in ("LLaMAMLP", "GemmaMLP"):
        specs = third_stages(plan_third_units(n_nodes, config))
        for sp in specs:
            sp["layers"] = round(sp["units"] / 3, 2)
        return specs
    if policy in ("half", "third") and sub_ok:
        units = plan_half_units(n_nodes, config)
        return [{"n_blocks": h.n_blocks, "first_parts": "mlp" if h.first_mlp_only else "both",
                 "last_parts": "attn" if h.last_attn_only else "both", "first_mlp_only": h.first_mlp_only,
                 "last_attn_only": h.last_attn_only, "layer_offset": h.lo_layer, "units": u, "unit": "half", "layers": u / 2}
                for h, u in zip(half_stages(units), units)]
    pol = "balanced" if policy in ("half", "third") else policy
    plan = plan_layers(n_nodes, config.n_layer, config, policy=pol) if n_nodes > 1 else [config.n_layer]
    out, off = [], 0
    for n in plan:
        out.append({"n_blocks": n, "first_parts": "both", "last_parts": "both", "first_mlp_only": False, "last_attn_only": False,
                    "layer_offset": off, "units": n, "unit": "layer", "layers": float(n)})
        off += n
    return out


def layer_ranges(plan: Sequence[int]) -> List[Tuple[int, int]]:
    out, start = [], 0
    for n in plan:
        out.append((start, start + n))
        start += n
    return out


def count_transformer_blocks(state_dict: Dict[str, Any], base_name_transformer: str = "transformer") -> int:
    """Number of distinct ``transformer.h.<i>`` blocks in a state dict (utils.py:470-492)."""
    prefix = f"{base_name_transformer}.h."
    return len({k[len(prefix):].split(".", 1)[0] for k in state_dict if k.startswith(prefix)})


def split_parameters(
    model_params: Dict[str, Any],
    n_nodes: int,
    plan: Optional[Sequence[int]] = None,
    config: Optional[Config] = None,
    head_on: str = "starter",
) -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """Pop the entries of a full litGPT state dict into per-node chunks.

    ``head_on="finisher"`` gives the first-generation layout (old/nanoGPT/sub/model_dist.py:90-221):
    ``ln_f`` + ``lm_head`` go to the last chunk instead of the starter (a tied head is cloned).

    Returns ``({"starter": sd, "secondary": [sd, ...]}, layers_info)`` where ``layers_info`` has
    the reference's ``N_LAYERS_START`` / ``N_LAYERS_SECONDARY`` keys (the latter is the first
    secondary's count) plus ``"plan"`` with the full per-stage list.  ``model_params`` is
    consumed: whatever is left afterwards was not assigned (the caller warns, utils.py:411-412).
    """
    if n_nodes < 2:
        raise ValueError("There must be at least 2 nodes in the network")
    n_layer = count_transformer_blocks(model_params)
    plan = list(plan) if plan is not None else plan_layers(n_nodes, n_layer, config)
    if len(plan) != n_nodes or sum(plan) != n_layer:
        raise ValueError(f"plan {plan} does not cover {n_layer} layers over {n_nodes} nodes")

    by_layer: Dict[int, List[str]] = {}
    for k in model_params:
        if k.startswith("transformer.h."):
            by_layer.setdefault(int(k.split(".")[2]), []).append(k)

    def take_layers(lo: int, hi: int) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for li in range(lo, hi):
            for k in by_layer.get(li, ()):
                tail = k.split(".", 3)[3]
                out[f"transformer.h.{li - lo}.{tail}"] = model_params.pop(k)
        return out

    ranges = layer_ranges(plan)
    starter: Dict[str, Any] = {}
    for k in ("transformer.wte.weight", "transformer.wte.bias", "transformer.wpe.weight"):
        if k in model_params:
            starter[k] = model_params.pop(k)
    starter.update(take_layers(*ranges[0]))
    if head_on not in ("starter", "finisher"):
        raise ValueError(f"head_on must be 'starter' or 'finisher', got {head_on!r}")
    head: Dict[str, Any] = {}
    for k in ("transformer.ln_f.weight", "transformer.ln_f.bias", "lm_head.weight", "lm_head.bias"):
        if k in model_params:
            head[k] = model_params.pop(k)
    secondary = [take_layers(lo, hi) for lo, hi in ranges[1:]]
    if head_on == "starter":
        starter.update(head)
    else:
        if "lm_head.weight" not in head and "transformer.wte.weight" in starter:  # tied embeddings
            head["lm_head.weight"] = starter["transformer.wte.weight"].clone()
        secondary[-1].update(head)
    info = {"N_LAYERS_START": plan[0], "N_LAYERS_SECONDARY": plan[1], "plan": list(plan)}
    return {"starter": starter, "secondary": secondary}, info


def chunk_dir(ckpt_dir: Union[str, Path], n_nodes: int) -> Path:
    return Path(ckpt_dir) / "chunks" / f"{n_nodes}nodes"


def chunk_file(ckpt_dir: Union[str, Path], n_nodes: int, role: str) -> Path:
    """``role`` = "starter" or "secondary:<i>"."""
    d = chunk_dir(ckpt_dir, n_nodes)
    if role.startswith("starter"):
        return d / "model_starter.pth"
    return d / f"model_secondary{int(role.split(':')[1])}.pth"


def split_and_store(
    model_params: Dict[str, Any],
    n_nodes: int,
    ckpt_dir: Union[str, Path],
    plan: Optional[Sequence[int]] = None,
    config: Optional[Config] = None,
    **kwargs: Any,
) -> Path:
    """Split a state dict and write the chunk files; returns the chunk directory."""
    verb = bool(kwargs.get("verb", False))
    units = kwargs.get("units")
    specs = kwargs.get("specs")
    if specs is not None and specs[0].get("unit") == "third":  # a cut layer's sub-units land in neighbouring chunks
        chunks = split_parameters_units(model_params, specs)
        info = {"plan": [sp["layers"] for sp in specs]}
    elif units is not None:  # half-layer plan
        chunks = split_parameters_half(model_params, units)
        info = {"plan": [u / 2 for u in units]}
    else:
        chunks, info = split_parameters(model_params, n_nodes, plan=plan, config=config,
                                        head_on=kwargs.get("head_on", "starter"))
    if len(model_params):
        warnings.warn(f"{len(model_params)} elements have not been used")
    del model_params
    gc.collect()
    if verb:
        print(f"Using the following split: starter {info['plan'][0]} layers, secondaries {info['plan'][1:]}")
    out = chunk_dir(ckpt_dir, n_nodes)
    os.makedirs(out, exist_ok=True)
    torch.save(chunks["starter"], out / "model_starter.pth")
    for i, sd in enumerate(chunks["secondary"]):
        torch.save(sd, out / f"model_secondary{i}.pth")
    return out


def merge_chunks(starter: Dict[str, Any], secondary: Sequence[Dict[str, Any]]) -> Dict[str, Any]:
    """Inverse of :func:`split_parameters` (used by tests: split ∘ merge = identity)."""
    full: Dict[str, Any] = {}
    n0 = count_transformer_blocks(starter)
    for k, v in starter.items():
        full[k] = v
    base = n0
    for sd in secondary:
        n = count_transformer_blocks(sd)
        for k, v in sd.items():
            parts = k.split(".", 3)
            full[f"transformer.h.{int(parts[2]) + base}.{parts[3]}"] = v
        base += n
    return full
