"""Shared CLI plumbing: paths, logging setup, CSV output of run statistics."""
from __future__ import annotations

import logging
import os
from datetime import datetime
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent
SETTINGS_DIR = PKG_DIR / "settings_distr"
LOGS_DIR = Path(os.environ.get("MDI_LOGS_DIR", Path.cwd() / "logs"))
IMG_DIR = Path(os.environ.get("MDI_IMG_DIR", Path.cwd() / "img"))

CSV_HEADER_STATS = ",".join(["timestamp", "n_samples", "n_layers", "context_size", "gen_time"])


def setup_debug_log(name: str) -> Path:
    """``-d``: DEBUG log of the ``model_dist`` logger into ``logs/<name>`` (starter.py:35-44)."""
    log_file = LOGS_DIR / name
    log_file.parent.mkdir(parents=True, exist_ok=True)
    log = logging.getLogger("model_dist")
    handler = logging.FileHandler(log_file, mode="w")
    handler.setFormatter(logging.Formatter("[%(asctime)s] → %(levelname)s: %(message)s"))
    log.setLevel(logging.DEBUG)
    log.addHandler(handler)
    return log_file


def tokens_time_csv_name(n_nodes: int, model: str, n_samples: int) -> str:
    """File name convention of the reference (starter.py:73)."""
    return f"tokens_time_samples_{n_nodes}nodes_{model}_{n_samples}samples.csv"


def append_run_stats(path: Path, n_samples: int, n_layers: int, context: int, gen_time: float) -> None:
    """``--time-run``: one CSV row per run, header on creation (starter.py:90-105)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    new = not path.exists()
    with open(path, "a") as f:
        if new:
            f.write(CSV_HEADER_STATS + "\n")
        f.write(f"{datetime.now().strftime('%Y-%m-%d %H:%M:%S')},{n_samples},{n_layers},{context},{gen_time}\n")
    print("Stats written to ", path)


def seed_everything(seed: int) -> None:
    import torch

    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
