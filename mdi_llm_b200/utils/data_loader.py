"""Text data set loading and random-crop batching for the trainer.

Parity: reference ``src/sub/utils/data_loader.py`` — ``load_dataset`` (:14-46), ``split_dataset``
(:49-66), ``get_batch`` (:69-126: ``batch_size`` random windows of ``block_size`` tokens from a
tensor or a uint16 ``np.memmap``; pinned + non-blocking H2D on CUDA).
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Tuple, Union

import numpy as np
import torch

__all__ = ["load_dataset", "split_dataset", "get_batch"]

_TEXT_SUFFIXES = (".txt", ".tex", ".md")


def load_dataset(input_path: Union[str, Path], tokenizer: Any, *args: Any, device: str = "cpu", **kwargs: Any) -> torch.Tensor:
    p = Path(input_path)
    if not p.is_file():
        raise ValueError(f"Could not find {p}")
    if not p.name.lower().endswith(_TEXT_SUFFIXES):
        raise ValueError(f"File format not supported!\nSupported formats: {_TEXT_SUFFIXES}")
    text = p.read_text(encoding="utf-8")
    out = tokenizer.encode(text)
    if not isinstance(out, torch.Tensor):
        out = torch.tensor(out, dtype=torch.int)
    return out.to(device)


def split_dataset(data: torch.Tensor, frac_train: float = 0.9) -> Tuple[torch.Tensor, torch.Tensor]:
    """Contiguous split (no shuffling: token order matters)."""
    if not 0 <= frac_train <= 1:
        raise AssertionError("frac_train must be in [0, 1]")
    n = int(frac_train * len(data))
    return data[:n], data[n:]


def get_batch(dataset: Any, batch_size: int, device: str, model_conf: Any) -> Tuple[torch.Tensor, torch.Tensor]:
    """``x[b] = data[i:i+T]``, ``y[b] = data[i+1:i+T+1]`` for random ``i``."""
    T = model_conf.block_size
    hi = len(dataset) - T
    if hi <= 0:
        raise ValueError(f"data set of {len(dataset)} tokens is shorter than block_size {T}")
    starts = torch.randint(hi, (batch_size,)).tolist()
    if isinstance(dataset, torch.Tensor):
        x = torch.stack([dataset[i: i + T] for i in starts]).long()
        y = torch.stack([dataset[i + 1: i + 1 + T] for i in starts]).long()
    elif isinstance(dataset, np.ndarray):  # includes np.memmap
        x = torch.from_numpy(np.stack([dataset[i: i + T] for i in starts]).astype(np.int64))
        y = torch.from_numpy(np.stack([dataset[i + 1: i + 1 + T] for i in starts]).astype(np.int64))
    else:
        raise TypeError(f"Unsupported data type {type(dataset)}")
    if "cuda" in str(device):
        return (x.pin_memory().to(device, non_blocking=True), y.pin_memory().to(device, non_blocking=True))
    return x.to(device), y.to(device)
