"""Framework-wide constants and run-time defaults.

Parity: the non-registry part of reference ``src/sub/config.py`` (:12-166): default device,
training hyper-parameters, generation defaults (``TOP_K=200``, ``TEMPERATURE=0.8``), the wire
constants of the socket transport (``HEADERLENGTH=16``), dtype mapping and ``TrainingConfig``.
The partition table lives in :mod:`mdi_llm_b200.models.partition` and the model registry in
:mod:`mdi_llm_b200.models.registry`; both are re-exported here for drop-in imports.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from .models.partition import N_LAYERS_NODES  # noqa: F401  (re-export)
from .models.registry import configs, name_to_config  # noqa: F401  (re-export)


def default_device() -> str:
    if torch.cuda.is_available():
        return "cuda"
    if getattr(torch.backends, "mps", None) is not None and torch.backends.mps.is_available():
        return "mps"
    return "cpu"


DEVICE = default_device()

# ---- training defaults (train.py) -------------------------------------------------------------
INIT_FROM = "scratch"  # "scratch" | "resume" | "hf"
BATCH_SIZE = 24
MAX_ITERS = N_ITER_TRAIN = 600_000
GRADIENT_ACCUMULATION_STEPS = 4
CKPT_INTERVAL = 2000
EVAL_ITERS = 200
LOG_INTERVAL = 10
EVAL_ONLY = False
ALWAYS_SAVE_CHECKPOINT = False
WEIGHT_DECAY = 1e-1
BETA1 = 0.9
BETA2 = 0.95
GRAD_CLIP = 1.0
LEARNING_RATE = 3e-4
DECAY_LR = True
WARMUP_ITERS = 2000
LR_DECAY_ITERS = 600_000
MIN_LR = 6e-5

# ---- generation defaults ----------------------------------------------------------------------
TOP_K = 200
TEMPERATURE = 0.8

# ---- socket transport wire format ---------------------------------------------------------------
HEADERLENGTH = 16  # ASCII decimal payload length, left-aligned, space padded
MSGLENGTH = 16 * 2048

# ---- dtypes -------------------------------------------------------------------------------------
DTYPE_TORCH_MAPPING: Dict[str, torch.dtype] = {
    "float32": torch.float32,
    "bfloat16": torch.bfloat16,
    "float16": torch.float16,
    "int8": torch.int8,
}
if hasattr(torch, "float8_e4m3fn"):
    DTYPE_TORCH_MAPPING["float8_e4m3fn"] = torch.float8_e4m3fn


def default_dtype() -> str:
    """bf16 where the device supports it, else fp16 (config.py:104-108)."""
    if torch.cuda.is_available() and torch.cuda.is_bf16_supported():
        return "bfloat16"
    return "float16"


DTYPE = default_dtype()
DTYPE_TORCH = DTYPE_TORCH_MAPPING[DTYPE]
COMPILE = False
BACKEND = "nccl"  # DDP backend for train.py (gloo on CPU-only hosts)

VERB = False
DEBUG = False
PLOTS = False


class TrainingConfig:
    """Mutable bag of trainer settings (config.py:119-162)."""

    tie_embeddings: bool = True
    learning_rate = LEARNING_RATE
    decay_lr = DECAY_LR
    min_lr = MIN_LR
    warmup_iters = WARMUP_ITERS
    lr_decay_iters = LR_DECAY_ITERS
    weight_decay = WEIGHT_DECAY
    beta1 = BETA1
    beta2 = BETA2
    grad_clip = GRAD_CLIP
    eval_only = EVAL_ONLY
    eval_iters = EVAL_ITERS
    batch_size = BATCH_SIZE
    max_iters = MAX_ITERS
    ckpt_interval = CKPT_INTERVAL
    log_interval = LOG_INTERVAL
    gradient_accumulation_steps = GRADIENT_ACCUMULATION_STEPS
    device = DEVICE
    compile = COMPILE
    _dtype_torch = DTYPE_TORCH

    def as_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for attr in dir(self):
            if attr.startswith("__") or attr in ("as_dict", "dtype"):
                continue
            val = getattr(self, attr, None)
            if val is not None and not callable(val):
                out[attr] = val
        return out

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype_torch

    @dtype.setter
    def dtype(self, value: str) -> None:
        if value not in DTYPE_TORCH_MAPPING:
            raise ValueError(f"Supported dtypes are: {list(DTYPE_TORCH_MAPPING)}")
        self._dtype_torch = DTYPE_TORCH_MAPPING[value]


# ---- per-family config lists (the reference's module-level names, config.py:170-1669) --------------
# The registry is table-driven (models/registry.py); these views group it the way the reference's source does
# so that ``from sub.config import llama_3`` style code keeps working.
def _family(*patterns: str):
    import re as _re

    rx = [_re.compile(p) for p in patterns]
    return [c for c in configs if any(r.search(c["name"]) for r in rx)]


stablecode = _family(r"^stablecode", r"^stable-code")
pythia = _family(r"^pythia")
dolly = _family(r"^dolly")
redpajama_incite = _family(r"^RedPajama-INCITE")
falcon = _family(r"^falcon-(7|40)b")
falcon180b = _family(r"^falcon-180B")
open_LLaMA = _family(r"^open_llama")
vicuna = _family(r"^vicuna")
long_chat = _family(r"^longchat")
nous_research = _family(r"^Nous-Hermes")
llama_2 = _family(r"^Llama-2-\d+b(-chat)?-hf$")
llama_3 = _family(r"^Llama-3")
gemma = _family(r"^Gemma")
codegemma = _family(r"^CodeGemma")
danube2 = _family(r"^Danube2")
freewilly_2 = _family(r"^FreeWilly2")
code_llama = _family(r"^CodeLlama")
platypus = _family(r"Platypus")
together_llama2_32k = _family(r"^LLaMA-2-7B-32K")
phi = _family(r"^phi")
mistral = _family(r"^Mistral|^Mixtral")
tiny_llama = _family(r"^tiny-llama")
llama_2_function_calling = _family(r"function-calling")
