#!/usr/bin/env python3
"""Print the layer-partition table as JSON (the reference ships it as ``src/sub/split_map.json``,
a file no code reads) and, for a given model, the plan the engine would use for any node count."""
from __future__ import annotations

import argparse
import json


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("--model", default=None, help="registry name: show table vs balanced plans for 1..--max-nodes")
    p.add_argument("--max-nodes", type=int, default=8)
    a = p.parse_args(argv)
    from ..models.config import Config
    from ..models.partition import N_LAYERS_NODES, plan_layers

    if a.model is None:
        print(json.dumps({str(k): {str(l): v for l, v in per.items()} for k, per in N_LAYERS_NODES.items()}, indent=2))
        return 0
    cfg = Config.from_name(a.model)
    for n in range(1, a.max_nodes + 1):
        try:
            table = plan_layers(n, cfg.n_layer, cfg, policy="table")
        except KeyError:
            table = None
        print(f"{n} nodes: table={table} balanced={plan_layers(n, cfg.n_layer, cfg, policy='balanced')}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
