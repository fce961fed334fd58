"""Stand-in for the two `accelerate` entry points the reference calls (gptserver.py:663,674):
``init_empty_weights`` and ``load_checkpoint_and_dispatch(model, path, dtype=...)``.
The latter loads the chunk state dict (mmap) and assigns it into the already-built module."""
from __future__ import annotations

from contextlib import contextmanager

import torch


@contextmanager
def init_empty_weights():
    with torch.device("meta"):
        yield


def load_checkpoint_and_dispatch(model, checkpoint, dtype=None, **kwargs):
    sd = torch.load(str(checkpoint), map_location="cpu", mmap=True, weights_only=True)
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    model.load_state_dict(sd, assign=True, strict=True)
    return model
