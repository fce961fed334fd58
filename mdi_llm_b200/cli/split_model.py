#!/usr/bin/env python3
"""Offline splitter: cut a local checkpoint into per-node chunks without running anything.

Parity: reference ``old/GPT2/split_model.py`` (:1-133: load a checkpoint, ``split_parameters`` for
``--n-nodes``, write one file per node) — in the current generation the same job is the tail of
``prepare_model.py`` (:57-58).  Writes ``<ckpt>/chunks/<N>nodes/model_starter.pth`` and
``model_secondary{i}.pth``.
"""
from __future__ import annotations

import argparse
from pathlib import Path


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("ckpt", type=Path, help="litGPT checkpoint directory (lit_model.pth + model_config.yaml)")
    p.add_argument("--n-nodes", type=int, required=True)
    p.add_argument("--partition", default="auto", choices=["auto", "table", "balanced"])
    p.add_argument("--head-on", default="starter", choices=["starter", "finisher"])
    p.add_argument("-v", "--verb", action="store_true")
    a = p.parse_args(argv)
    from ..models.partition import plan_layers, split_and_store
    from ..utils.checkpoint import load_from_pt

    cfg, sd = load_from_pt(a.ckpt)
    assert sd is not None
    plan = plan_layers(a.n_nodes, cfg.n_layer, cfg, policy=a.partition)
    out = split_and_store(sd, a.n_nodes, a.ckpt, plan=plan, config=cfg, verb=a.verb, head_on=a.head_on)
    print(f"plan {plan} -> {out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
