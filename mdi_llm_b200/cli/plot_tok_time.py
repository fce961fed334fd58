#!/usr/bin/env python3
"""Overlay tokens-vs-time curves of runs with different node counts.

Parity: reference ``src/plot_tok_time.py`` (:17-93): reads
``logs/tokens_time_samples_{k}nodes_{model}_{n}samples.csv`` for k = 1..N and plots them on one
figure.  Without matplotlib a text table (final tokens/s per node count and speed-up vs 1 node)
is printed instead.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Dict, List, Tuple

from .common import LOGS_DIR, tokens_time_csv_name


def read_points(path: Path) -> List[Tuple[float, int]]:
    pts = []
    for line in path.read_text().splitlines():
        if line.strip():
            t, n = line.split(",")[:2]
            pts.append((float(t), int(float(n))))
    return pts


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("MODEL_DIR", nargs="?", type=Path, default=None,
                   help="checkpoint directory of the model (the reference's positional form: its name selects the CSVs)")
    p.add_argument("--model", default=None, help="model name, instead of MODEL_DIR")
    p.add_argument("-nt", "--no-title", action="store_true", help="if set, don't print the figure title")
    p.add_argument("-n", "--n-samples", type=int, default=3)
    p.add_argument("--max-nodes", type=int, default=8)
    p.add_argument("--logs-dir", type=Path, default=LOGS_DIR)
    p.add_argument("-o", "--out", type=Path, default=None)
    a = p.parse_args(argv)
    if a.model is None:
        if a.MODEL_DIR is None:
            p.error("give the model as MODEL_DIR or --model")
        a.model = a.MODEL_DIR.name
    curves: Dict[int, List[Tuple[float, int]]] = {}
    for k in range(1, a.max_nodes + 1):
        f = a.logs_dir / tokens_time_csv_name(k, a.model, a.n_samples)
        if f.is_file():
            curves[k] = read_points(f)
    if not curves:
        print(f"no tokens_time CSVs for model {a.model!r} in {a.logs_dir}")
        return 1
    rates = {k: (pts[-1][1] / pts[-1][0] if pts[-1][0] > 0 else 0.0) for k, pts in curves.items()}
    base = rates.get(1)
    print(f"{'nodes':>5} {'tokens':>8} {'time (s)':>10} {'tok/s':>10} {'speed-up':>9}")
    for k, pts in sorted(curves.items()):
        print(f"{k:>5} {pts[-1][1]:>8} {pts[-1][0]:>10.3f} {rates[k]:>10.2f} {rates[k] / base if base else float('nan'):>9.2f}")
    from ..utils.plots import have_matplotlib

    if have_matplotlib():
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt

        plt.figure(figsize=(12, 6))
        for k, pts in sorted(curves.items()):
            plt.plot([t for t, _ in pts], [n for _, n in pts], label=f"{k} node{'s' if k > 1 else ''}")
        plt.xlabel("Time (s)"); plt.ylabel("Tokens"); plt.grid(); plt.legend()
        if not a.no_title:
            plt.title(f"{a.model}: generated tokens vs time")
        out = a.out or (a.logs_dir / f"tokens_time_{a.model}_{a.n_samples}samples.png")
        plt.savefig(out)
        print(f"plot saved to {out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
