"""Small helpers shared by the runtime, the CLIs and the trainer.

Parity: reference ``src/sub/utils/utils.py`` — ``get_obj_size`` (:27-57), ``estimate_loss``
(:60-107), ``get_lr`` (:110-130), ``loading_bar`` (:133-162), ``waiting_animation`` (:165-173),
``remove_prefix`` (:176-183), ``find_eot`` (:185-212), ``detect_stop_tokens`` (:215-225),
``serialize_params`` / ``deserialize_params`` (:441-467).
"""
from __future__ import annotations

import gc
import math
import sys
import threading
from contextlib import nullcontext
from typing import Any, Dict, List, Mapping, Sequence, Tuple, Union

import torch
from torch import nn

from .. import config as C
from .data_loader import get_batch

__all__ = [
    "format_output",
    "get_obj_size", "estimate_loss", "get_lr", "loading_bar", "waiting_animation", "remove_prefix",
    "find_eot", "detect_stop_tokens", "serialize_params", "deserialize_params", "as_id_list",
]


def get_obj_size(obj: Any) -> int:
    """Deep ``sys.getsizeof`` following ``gc`` referents (types excluded)."""
    seen = {id(obj)}
    frontier = [obj]
    total = 0
    while frontier:
        total += sum(sys.getsizeof(o) for o in frontier)
        nxt = {}
        for o in gc.get_referents(*frontier):
            if id(o) not in seen and not isinstance(o, type):
                nxt[id(o)] = o
        seen.update(nxt)
        frontier = list(nxt.values())
    return total


@torch.no_grad()
def estimate_loss(model: nn.Module, train: Any, val: Any, batch_size: int, device: str,
                  *args: Any, **kwargs: Any) -> Dict[str, float]:
    """Mean cross-entropy over ``eval_iters`` random batches of each split."""
    ctx = kwargs.get("ctx", nullcontext())
    iters = int(kwargs.get("eval_iters", C.EVAL_ITERS))
    out: Dict[str, float] = {}
    was_training = model.training
    model.eval()
    cfg = model.module.config if hasattr(model, "module") else model.config
    for split, data in (("train", train), ("val", val)):
        acc = 0.0
        for _ in range(iters):
            x, y = get_batch(data, batch_size, device, cfg)
            with ctx:
                logits = model(x)
                loss = nn.functional.cross_entropy(logits.view(-1, logits.size(-1)).float(), y.view(-1),
                                                   ignore_index=-1)
            acc += float(loss)
        out[split] = acc / max(1, iters)
    model.train(was_training)
    return out


def get_lr(it: int, lr: float = C.LEARNING_RATE, min_lr: float = C.MIN_LR,
           warmup_it: int = C.WARMUP_ITERS, lr_decay_it: int = C.LR_DECAY_ITERS) -> float:
    """Linear warm-up, cosine decay to ``min_lr``, then flat."""
    if it < warmup_it:
        return lr * it / warmup_it
    if it > lr_decay_it:
        return min_lr
    ratio = (it - warmup_it) / max(1, lr_decay_it - warmup_it)
    return min_lr + 0.5 * (1.0 + math.cos(math.pi * ratio)) * (lr - min_lr)


def loading_bar(current_iter: int, tot_iter: int, n_chars: int = 10, ch: str = "=", n_ch: str = " ") -> str:
    done = int(current_iter * n_chars / max(1, tot_iter))
    return "[" + ch * done + n_ch * max(0, n_chars - done - 1) + "]"


def waiting_animation(text: str, stopping: threading.Event, period: float = 0.5) -> None:
    frames = "⠴⠦⠇⠋⠙⠸"
    stopping.clear()
    i = 0
    while not stopping.is_set():
        print(f"{text} {frames[i % len(frames)]}", end="\r")
        i += 1
        stopping.wait(period)
    print("")


def remove_prefix(text: str, prefix: str) -> str:
    return text[len(prefix):] if text.startswith(prefix) else text


def as_id_list(tokens: Union[torch.Tensor, Sequence[int]]) -> List[int]:
    if isinstance(tokens, torch.Tensor):
        return tokens.reshape(-1).tolist()
    return [int(t) for t in tokens]


def find_eot(tokens: torch.Tensor, stop_tokens: Tuple[List[int], ...] = (), prompt_length: int = 0) -> torch.Tensor:
    """Truncate ``tokens`` ``(1, L)`` right before the end of the first stop sequence that
    completes after the prompt (same indexing as the reference: the returned tensor still
    contains the first ``len(stop) - ...`` tokens up to index ``i`` exclusive)."""
    ids = as_id_list(tokens)
    if len(ids) < prompt_length:
        raise AssertionError("Prompt length must be longer than the provided tensor")
    if not stop_tokens:
        return tokens
    start = prompt_length + max(len(s) for s in stop_tokens)
    for i in range(start, len(ids)):
        for s in stop_tokens:
            if ids[i - len(s): i] == list(s):
                return tokens.reshape(1, -1)[:, :i]
    return tokens


def detect_stop_tokens(tokens: Union[torch.Tensor, Sequence[int]], stop_tokens: Tuple[List[int], ...] = ()) -> bool:
    """True when the sequence ends with one of the stop sequences."""
    ids = as_id_list(tokens)
    return any(len(s) <= len(ids) and ids[-len(s):] == list(s) for s in stop_tokens if len(s))


def serialize_params(params: Mapping[str, Any]) -> Dict[str, Any]:
    return {k: (v.tolist() if isinstance(v, torch.Tensor) else v) for k, v in params.items()}


def deserialize_params(params: Dict[str, Any]) -> Dict[str, Any]:
    return {k: (torch.tensor(v) if isinstance(v, list) else v) for k, v in params.items()}


def format_output(text: str, user_tag: str = "<|user|>", assistant_tag: str = "<|assistant|>", color: bool = False) -> str:
    """Pretty-print a chat transcript: every ``<|user|>`` / ``<|assistant|>`` turn on its own paragraph with
    a speaker label (optionally ANSI-coloured).  The reference leaves this as an empty stub
    (``utils.py:228-238``: "isolate the <|user|> and <|assistant|> elements ... maybe format with color")."""
    import re as _re

    parts = _re.split(f"({_re.escape(user_tag)}|{_re.escape(assistant_tag)})", text)
    out, who = [], None
    for part in parts:
        if part == user_tag:
            who = "user"
        elif part == assistant_tag:
            who = "assistant"
        elif part.strip():
            label = {"user": "User", "assistant": "Assistant", None: ""}[who]
            if color and label:
                label = ("\033[36m" if who == "user" else "\033[33m") + label + "\033[0m"
            out.append(f"{label}: {part.strip()}" if label else part.strip())
    return "\n\n".join(out)
