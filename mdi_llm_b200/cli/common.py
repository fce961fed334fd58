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


def parse_args(train: bool = True, argv=None):
    """The second generation's shared argument parser (reference ``old/GPT2/sub/parser.py:13-140``): one
    function builds either the trainer's or the generation scripts' flags.  The current CLIs have their own
    parsers (same flag names); this keeps scripts written against ``sub.parser.parse_args`` working."""
    import argparse

    from .. import config as C

    p = argparse.ArgumentParser(description="Training" if train else "Generation")
    p.add_argument("-v", "--verb", default=False, action="store_true", help="Enable verbose mode")
    p.add_argument("-d", "--debug", default=False, action="store_true", help="Enable debug mode (file log)")
    p.add_argument("--device", type=str, default=None, help="torch device where to load model and tensors")
    if train:
        p.add_argument("--ckpt", default=None, help="checkpoint folder / file name")
        p.add_argument("--dataset", type=Path, default=None, help="folder with train.bin / val.bin (or a text file)")
        p.add_argument("--batch-size", type=int, default=C.BATCH_SIZE)
        p.add_argument("--init", type=str, default="scratch", choices=["scratch", "resume", "hf"])
        p.add_argument("--max-iters", type=int, default=C.MAX_ITERS)
        p.add_argument("--log-interval", type=int, default=C.LOG_INTERVAL)
        p.add_argument("--ckpt-interval", type=int, default=C.CKPT_INTERVAL)
        p.add_argument("--grad-acc-steps", type=int, default=C.GRADIENT_ACCUMULATION_STEPS)
        p.add_argument("--patience", type=int, default=None)
        p.add_argument("--always-update", default=False, action="store_true")
    else:
        p.add_argument("-p", "--plots", default=False, action="store_true")
        p.add_argument("--ckpt", type=Path, default=None, help="checkpoint folder")
        p.add_argument("--time-run", type=Path, default=None, help="CSV file collecting run statistics")
        p.add_argument("--n-tokens", type=int, default=300)
        p.add_argument("--prompt", type=str, default="\n")
        p.add_argument("--n-samples", type=int, default=1)
    return p.parse_args(argv)
