#!/usr/bin/env python3
"""Secondary (worker) node of a model-distributed inference run.

Parity: reference ``src/secondary.py`` — flags (:55-98): ``-d -v -c --chunk --nodes-config PATH
IDX --device --dtype --seed`` plus ``--ckpt`` (chunk path derived from the checkpoint folder).
The node starts its control endpoint and waits for ``POST /init`` from the starter.
"""
from __future__ import annotations

import argparse
from pathlib import Path

from .common import SETTINGS_DIR, seed_everything, setup_debug_log


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Secondary node - MDI")
    p.add_argument("-d", "--debug", action="store_true")
    p.add_argument("-v", "--verb", action="store_true")
    p.add_argument("-c", "--compile", action="store_true", help="accepted for compatibility")
    p.add_argument("--chunk", type=Path, default=None, help="path of the model chunk of this node")
    p.add_argument("--ckpt", type=Path, default=None, help="checkpoint folder (chunk inferred from it)")
    p.add_argument("--nodes-config", type=str, nargs=2, metavar=("CONFIG-PATH", "SECONDARY-INDEX"),
                   default=[str(SETTINGS_DIR / "configuration.json"), "0"],
                   help="JSON node topology and the zero-based index of this secondary")
    p.add_argument("--secondary-config", type=Path, default=None,
                   help="JSON holding only THIS node's entry (addr / communication / inference), alternative to --nodes-config "
                        "(old/GPT2/secondary.py:46-52); the chunk then has to be given with --chunk or arrives with POST /init")
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--seed", type=int, default=10137)
    p.add_argument("--engine", default="auto", choices=["auto", "eager", "cuda"])
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    seed_everything(args.seed)
    if args.debug:
        setup_debug_log("logs_finisher.log")
    print("+---------------------------+\n| Launching secondary node |\n+---------------------------+")
    from ..parallel.distributed import GPTDistributed

    cfg_path, idx = Path(args.nodes_config[0]), int(args.nodes_config[1])
    if args.secondary_config is not None:
        cfg_path, idx = args.secondary_config, 0
    node = GPTDistributed(node_type=f"secondary:{idx}", config_file=cfg_path, ckpt_dir=args.ckpt, chunk_path=args.chunk,
                          device=args.device, dtype=args.dtype, verb=args.verb, compile=args.compile, engine=args.engine)
    node.start()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
