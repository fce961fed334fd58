"""Restricted unpickling for everything that arrives over the network.

The reference's wire formats are pickles (control bodies ``model_dist.py:516-533`` /
``gptserver.py:1143``, data-plane frames ``connections.py:207,338``), and ``pickle.loads`` on
bytes from a socket is arbitrary code execution.  The formats are kept (a reference node can still
talk to one of ours) but the *decoder* only resolves the handful of globals those messages really
contain: plain containers, tensors and parameters.  Tensor storages, which the stock pickle
rebuilds through ``torch.load(weights_only=False)``, go through ``weights_only=True`` instead.
Anything else raises :class:`pickle.UnpicklingError`.
"""
from __future__ import annotations

import collections
import io
import pickle
from typing import Any, Callable, Dict, Tuple

import torch

__all__ = ["safe_loads", "UnsafePayload"]


class UnsafePayload(pickle.UnpicklingError):
    pass


def _storage_from_bytes(b: bytes) -> Any:
    return torch.load(io.BytesIO(b), weights_only=True)


def _allowed() -> Dict[Tuple[str, str], Callable[..., Any]]:
    import torch._utils as tu

    table: Dict[Tuple[str, str], Any] = {
        ("collections", "OrderedDict"): collections.OrderedDict,
        ("torch._utils", "_rebuild_tensor_v2"): tu._rebuild_tensor_v2,
        ("torch._utils", "_rebuild_parameter"): tu._rebuild_parameter,
        ("torch.storage", "_load_from_bytes"): _storage_from_bytes,
        ("torch", "Size"): torch.Size,
        ("torch", "device"): torch.device,
        ("builtins", "set"): set,
        ("builtins", "frozenset"): frozenset,
        ("builtins", "complex"): complex,
        ("builtins", "bytearray"): bytearray,
    }
    for name in ("float32", "float16", "bfloat16", "float64", "int8", "uint8", "int16", "int32", "int64", "bool",
                 "float8_e4m3fn", "float8_e5m2"):
        if hasattr(torch, name):
            table[("torch", name)] = getattr(torch, name)
    return table


class _Restricted(pickle.Unpickler):
    _table = None

    def find_class(self, module: str, name: str) -> Any:
        if _Restricted._table is None:
            _Restricted._table = _allowed()
        try:
            return _Restricted._table[(module, name)]
        except KeyError:
            raise UnsafePayload(f"refusing to unpickle global {module}.{name}") from None


def safe_loads(data: bytes) -> Any:
    """``pickle.loads`` for network payloads: containers, scalars, strings, tensors, parameters only."""
    return _Restricted(io.BytesIO(data)).load()
