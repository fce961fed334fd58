#!/usr/bin/env python3
"""Download/convert a model if needed, then split it into per-node chunks.

Parity: reference ``src/prepare_model.py`` (:34-69): ``MODEL [--ckpt-folder --model-name --n-nodes
--hf-token --dtype --device]`` — a local directory is converted when it has no ``lit_model.pth``;
otherwise the model is fetched from the HF hub; then ``split_and_store``.  Extras:
``--random-init`` builds a random checkpoint of a registered architecture (no network on the GPU
box) and ``--partition`` chooses the reference table or the balanced planner.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("MODEL", type=str, help="HF repo id / registry name, or a local checkpoint directory")
    p.add_argument("--ckpt-folder", type=Path, default=Path("checkpoints"))
    p.add_argument("--model-name", type=str, default=None)
    p.add_argument("--n-nodes", type=int, default=None, help="number of nodes to split for (omit: no split)")
    p.add_argument("--hf-token", type=str, default=os.getenv("HF_TOKEN"))
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--device", type=str, default="cpu")
    p.add_argument("--random-init", action="store_true", help="create random weights instead of downloading")
    p.add_argument("--partition", default="auto", choices=["auto", "table", "balanced"])
    p.add_argument("--head-on", default="starter", choices=["starter", "finisher"],
                   help="which node owns ln_f + lm_head ('finisher' = first-generation chain layout)")
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--fit-engine", action="store_true",
                   help="re-parametrise a model whose head size / GQA width is outside the fused sm_100a engine (Phi-2: head size "
                        "80, Falcon-7B: 71 query heads per KV head) into an exactly equivalent one inside it "
                        "(utils/fit_engine.py); written to <checkpoint>-fused/, which is then split")
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    import torch

    from ..models.config import Config
    from ..models.partition import plan_layers, split_and_store
    from ..models.registry import lookup
    from ..utils.checkpoint import load_from_pt, write_random_checkpoint

    local = Path(args.MODEL)
    if args.random_init:
        cfg = Config.from_name(args.model_name or local.name)
        model_path = local if local.is_dir() or os.sep in args.MODEL else args.ckpt_folder / cfg.hf_config.get("org", "custom") / cfg.hf_config.get("name", cfg.name)
        dtype = getattr(torch, args.dtype) if args.dtype else torch.bfloat16
        write_random_checkpoint(model_path, cfg, dtype=dtype, seed=args.seed)
    elif local.is_dir():
        model_path = local
        if not (model_path / "lit_model.pth").is_file():
            from ..utils.convert_hf_checkpoint import convert_hf_checkpoint

            convert_hf_checkpoint(model_path, args.model_name, args.dtype)
    else:
        from ..utils.download import download_from_hub

        repo = args.MODEL
        if "/" not in repo:  # registry name -> org/name
            c = lookup(repo)
            repo = f"{c['hf_config']['org']}/{c['hf_config']['name']}"
        model_path = download_from_hub(repo_id=repo, access_token=args.hf_token, dtype=args.dtype,
                                       checkpoint_dir=args.ckpt_folder, model_name=args.model_name)
    cfg, sd = load_from_pt(model_path, args.device)
    print(f"Model {cfg.name}: {cfg.n_layer} layers, checkpoint at {model_path}")
    if args.fit_engine:
        import shutil

        from ..utils.fit_engine import fit_engine

        new_cfg, sd, notes = fit_engine(cfg, sd)
        if notes:
            fused = model_path.parent / (model_path.name + "-fused")
            fused.mkdir(parents=True, exist_ok=True)
            for f in model_path.iterdir():  # tokenizer, prompt style, generation config travel unchanged
                if f.is_file() and f.name not in ("lit_model.pth", "model_config.yaml"):
                    shutil.copy2(f, fused / f.name)
            new_cfg.save(fused)
            torch.save(sd, fused / "lit_model.pth")
            for n in notes:
                print(f"fit-engine: {n}")
            print(f"fit-engine: equivalent checkpoint written to {fused}")
            cfg, model_path = new_cfg, fused
        else:
            print("fit-engine: the architecture is already inside the fused engine, nothing to do")
    if args.n_nodes and args.n_nodes > 1:
        plan = plan_layers(args.n_nodes, cfg.n_layer, cfg, policy=args.partition)
        out = split_and_store(sd, args.n_nodes, model_path, plan=plan, config=cfg, verb=True, head_on=args.head_on)
        print(f"Chunks written to {out} (layers per node: {plan})")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
