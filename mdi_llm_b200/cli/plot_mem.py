#!/usr/bin/env python3
"""Overlay the memory-vs-time curves of the nodes of one run.

Parity: reference ``old/GPT2/plot_mem.py`` (:1-106) — one CSV per node written by ``mem_monitor``
(``logs/mem-usage/<model>/*<k>nodes*starter*.csv`` / ``*secondary<i>*.csv``), one line per node,
PNG under ``img/``.  Here the CSVs are this repo's ``mem_monitor`` format (``time_s, rss_mib,
gpu<i>_used_mib``); ``--column`` picks what is plotted (host RSS or one GPU's memory).
"""
from __future__ import annotations

import argparse
import csv
import re
from pathlib import Path
from typing import Dict, List, Tuple

from .common import IMG_DIR, LOGS_DIR

NODE_LABELS = {"starter": "First node"}


def node_label(fname: str) -> str:
    if "starter" in fname:
        return "First node"
    m = re.search(r"secondary(\d+)", fname)
    if m:
        return f"Node {int(m.group(1)) + 2}"
    return Path(fname).stem


def read_curve(path: Path, column: str) -> Tuple[List[float], List[float]]:
    with open(path, newline="") as f:
        rows = list(csv.DictReader(f))
    if rows and column not in rows[0]:
        raise KeyError(f"{path}: no column {column!r} (has {list(rows[0])})")
    return [float(r["time_s"]) for r in rows], [float(r[column]) for r in rows]


def collect(folder: Path, n_nodes: int, column: str) -> Dict[str, Tuple[List[float], List[float]]]:
    tag = "single" if n_nodes == 1 else f"{n_nodes}nodes"
    out = {}
    for p in sorted(folder.glob("*.csv")):
        if tag in p.name or (n_nodes == 1 and "1nodes" in p.name):
            out[node_label(p.name)] = read_curve(p, column)
    return out


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("MODEL", type=str, help="model whose mem_monitor logs are plotted (folder name under --logs)")
    p.add_argument("N_NODES", type=int, help="number of nodes of the run")
    p.add_argument("--logs", type=Path, default=LOGS_DIR / "mem-usage")
    p.add_argument("--column", default="rss_mib")
    p.add_argument("-o", "--out", type=Path, default=None)
    a = p.parse_args(argv)
    folder = a.logs / a.MODEL
    if not folder.is_dir():
        raise SystemExit(f"Error: folder not found {folder}")
    curves = collect(folder, a.N_NODES, a.column)
    if not curves:
        raise SystemExit(f"no CSV for {a.N_NODES} node(s) in {folder}")
    for label, (t, y) in curves.items():
        print(f"{label}: {len(t)} samples, peak {max(y):.0f} MiB")
    from ..utils.plots import have_matplotlib

    if not have_matplotlib():
        print("matplotlib not installed: no figure written")
        return 0
    import matplotlib

    matplotlib.use("Agg")
    import matplotlib.pyplot as plt

    plt.figure(figsize=(6, 5))
    for label, (t, y) in curves.items():
        plt.plot(t, y, label=label, linewidth=2)
    plt.ylabel("Memory usage (MiB)"); plt.xlabel("Time (s)")
    plt.grid(); plt.minorticks_on(); plt.legend()
    plt.title(f"Memory usage, {a.N_NODES} node(s), {a.MODEL}")
    plt.tight_layout()
    out = a.out or IMG_DIR / f"mem_in_time_{a.MODEL}_{a.N_NODES}nodes.png"
    out.parent.mkdir(parents=True, exist_ok=True)
    plt.savefig(out, dpi=300)
    print(out)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
