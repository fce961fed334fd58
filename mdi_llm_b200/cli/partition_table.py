#!/usr/bin/env python3
"""Print the layer-partition table as JSON (the reference ships it as ``src/sub/split_map.json``,
a file no code reads) and, for a given model, the plan the engine would use for any node count."""
from __future__ import annotations

import argparse
import json


def stage_times_us(cfg, specs):
    """Predicted decode time per stage step (microseconds) of a plan: the sub-layer units each stage owns, priced by the
    fitted cost model (gated MLPs: attention | gate/up | down units; others: attention | MLP), plus the fixed cost of a
    step and, on the starter, the output head + sampler."""
    from ..models.gpt import part_units
    from ..models.partition import decode_third_costs, decode_unit_costs

    if cfg.mlp_class_name in ("LLaMAMLP", "GemmaMLP"):
        ca, cg, cd, head, fixed = decode_third_costs(cfg)
        unit = {"attn": ca, "gu": cg, "down": cd}
    else:
        ca, cm, head = decode_unit_costs(cfg)
        unit, fixed = {"attn": ca, "gu": cm, "down": 0.0}, 9.0
    out = []
    for si, sp in enumerate(specs):
        t, nb = fixed + (head if si == 0 else 0.0), sp["n_blocks"]
        for j in range(nb):
            if nb == 1:
                units = [u for u in part_units(sp["first_parts"]) if u in part_units(sp["last_parts"])]
            else:
                units = part_units(sp["first_parts"]) if j == 0 else (part_units(sp["last_parts"]) if j == nb - 1 else ("attn", "gu", "down"))
            t += sum(unit[u] for u in units)
        out.append(t)
    return out


def predict(cfg, max_nodes: int) -> int:
    from ..models.memory import B200_HBM_BYTES, plan_memory
    from ..models.partition import stage_specs

    print(f"{cfg.name}: predicted decode step per stage and ring throughput (n_samples = n_nodes, ~0.5k context, bf16);\n"
          f"GB = the fullest stage's weights + KV slots for n_nodes samples x {min(cfg.block_size, 4096)} positions + hop buffers")
    one_gpu = 1e6 / stage_times_us(cfg, stage_specs(1, cfg, "balanced"))[0]
    for n in range(1, max_nodes + 1):
        for policy in ("table", "balanced", "half", "third"):
            try:
                specs = stage_specs(n, cfg, policy)
            except (KeyError, ValueError):
                continue
            ts = stage_times_us(cfg, specs)
            layers = "/".join(f"{sp['layers']:g}" for sp in specs)
            rate = 1e6 / max(ts)  # a full ring emits one token per step of its slowest stage
            gb = max(m["total"] for m in plan_memory(cfg, specs, n, min(cfg.block_size, 4096))) / 1e9
            fits = "" if gb * 1e9 <= 0.94 * B200_HBM_BYTES else "  [does NOT fit 180 GB]"
            print(f"{n} nodes  {policy:<8} layers {layers:<40} stage us {'/'.join(f'{t:.0f}' for t in ts)}  "
                  f"-> {rate:7.0f} tok/s ({rate / n / one_gpu:.2f} of n x 1 GPU)  {gb:6.1f} GB{fits}")
            if n == 1:
                break
    return 0


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("--model", default=None, help="registry name: show table vs balanced plans for 1..--max-nodes")
    p.add_argument("--max-nodes", type=int, default=8)
    p.add_argument("--predict", action="store_true",
                   help="with --model: per node count and partition policy, the stage times and ring throughput predicted by the "
                        "decode cost model fitted on a B200 (models/partition.py) — the planner's what-if view")
    a = p.parse_args(argv)
    from ..models.config import Config
    from ..models.partition import N_LAYERS_NODES, plan_layers

    if a.model is None:
        print(json.dumps({str(k): {str(l): v for l, v in per.items()} for k, per in N_LAYERS_NODES.items()}, indent=2))
        return 0
    cfg = Config.from_name(a.model)
    if a.predict:
        return predict(cfg, a.max_nodes)
    for n in range(1, a.max_nodes + 1):
        try:
            table = plan_layers(n, cfg.n_layer, cfg, policy="table")
        except KeyError:
            table = None
        print(f"{n} nodes: table={table} balanced={plan_layers(n, cfg.n_layer, cfg, policy='balanced')}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
