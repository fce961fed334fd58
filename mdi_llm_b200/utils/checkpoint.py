"""Checkpoint I/O in the litGPT / MDI-LLM on-disk layout.

Layout preserved (reference ``utils.py:527-611``)::

    <ckpt>/lit_model.pth  model_config.yaml  tokenizer.*  chunks/<N>nodes/model_starter.pth ...

Parity: ``load_sd`` (utils.py:495-524), ``load_from_pt`` (:527-562), ``save_config`` (:608-611),
``init_from_state_dict`` / ``get_keys_to_submodule`` (:614-662), ``lazy_load`` and
``incremental_save`` (litgpt_utils.py:14-343).  The lazy/incremental machinery is rebuilt on
mmap (``torch.load(mmap=True)`` and file-backed tensors) instead of a custom unpickler.
Additionally :func:`write_random_checkpoint` creates a random-init checkpoint of any registered
architecture — the GPU box has no network, so benchmarks and tests build their weights locally.
"""
from __future__ import annotations

import os
import tempfile
import warnings
from contextlib import contextmanager
from pathlib import Path
from typing import Any, Dict, Iterator, Optional, Tuple, Union

import torch
import torch.nn as nn

from ..models.config import Config

__all__ = [
    "load_sd", "load_from_pt", "load_from_hf", "save_config", "init_from_state_dict", "get_keys_to_submodule",
    "lazy_load", "incremental_save", "IncrementalSaver", "write_random_checkpoint",
    "random_state_dict", "materialize_stage",
]

PathLike = Union[str, Path]


def load_sd(model_path: PathLike, device: Optional[Union[torch.device, str]] = "cpu", **kwargs: Any) -> Dict[str, Any]:
    """``torch.load`` of a state dict with a CPU fallback on device OOM."""
    kwargs.setdefault("weights_only", True)
    try:
        return torch.load(model_path, map_location=device, **kwargs)
    except Exception as e:  # noqa: BLE001
        if "out of memory" in str(e) and str(device) != "cpu":
            warnings.warn(f"Unable to fit model ckpt in {device} memory! Retrying with cpu")
            return torch.load(model_path, map_location="cpu", **kwargs)
        raise


def lazy_load(path: PathLike) -> Dict[str, Any]:
    """State dict whose tensors are memory-mapped: bytes are paged in when first touched.
    Functional equivalent of litGPT's ``lazy_load`` (litgpt_utils.py:14-179)."""
    return torch.load(str(path), map_location="cpu", mmap=True, weights_only=True)


def load_from_pt(
    model_path: PathLike,
    device: Optional[Union[torch.device, str]] = "cpu",
    config_only: bool = False,
) -> Tuple[Config, Optional[Dict[str, Any]]]:
    """Read ``model_config.yaml`` (+ ``lit_model.pth``) from a checkpoint directory."""
    model_dir = Path(model_path)
    if not model_dir.is_dir():
        raise NotADirectoryError(f"Unable to find model checkpoint at {model_dir}")
    config = Config.from_file(model_dir / "model_config.yaml")
    if config_only:
        return config, None
    return config, load_sd(model_dir / "lit_model.pth", device)


def load_from_hf(
    repo_id: str,
    access_token: Optional[str] = None,
    dtype: Optional[str] = None,
    checkpoint_dir: PathLike = Path("checkpoints"),
    model_name: Optional[str] = None,
    device: Optional[Union[torch.device, str]] = "cpu",
    config_only: bool = False,
) -> Tuple[Config, Optional[Dict[str, Any]]]:
    """Download ``repo_id`` from the HF hub into ``checkpoint_dir/repo_id``, convert it to the
    litGPT layout and load it (reference utils.py:565-605).  Needs network + ``huggingface_hub``."""
    from .download import download_from_hub

    download_from_hub(repo_id=repo_id, access_token=access_token, dtype=dtype,
                      checkpoint_dir=Path(checkpoint_dir), model_name=model_name)
    return load_from_pt(Path(checkpoint_dir) / repo_id, device, config_only=config_only)


def save_config(config: Config, checkpoint_dir: PathLike) -> None:
    config.save(checkpoint_dir)


def get_keys_to_submodule(model: nn.Module) -> Dict[str, Tuple[nn.Module, str]]:
    """Map every parameter key of ``model`` to ``(owning leaf module, attribute name)``."""
    out: Dict[str, Tuple[nn.Module, str]] = {}
    for mod_name, mod in model.named_modules():
        for p_name, _ in mod.named_parameters(recurse=False):
            key = f"{mod_name}.{p_name}" if mod_name else p_name
            out[key] = (mod, p_name)
    return out


def init_from_state_dict(model: nn.Module, state_dict: Dict[str, Any], strict: bool = True) -> nn.Module:
    """Install the tensors of ``state_dict`` as the parameters of ``model`` *by reference*
    (no copy) — the way to fill a model built on the meta device without a 2x memory spike
    (utils.py:614-641).  Tied parameters (same object under two keys) stay tied."""
    targets = get_keys_to_submodule(model)
    installed: Dict[int, nn.Parameter] = {}  # id(old parameter object) -> its replacement
    pending = []
    for key, (mod, attr) in targets.items():
        old = getattr(mod, attr)
        if key in state_dict:
            new = nn.Parameter(state_dict[key], requires_grad=False)
            installed.setdefault(id(old), new)
            setattr(mod, attr, new)
        else:
            pending.append((key, mod, attr, old))
    missing = []
    for key, mod, attr, old in pending:  # keys absent from the dict: fine if tied to a present one
        if id(old) in installed:
            setattr(mod, attr, installed[id(old)])
        else:
            missing.append(key)
    if strict and missing:
        raise KeyError(f"state dict is missing keys: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    return model


def materialize_stage(model: nn.Module, state_dict: Dict[str, Any], device: Union[str, torch.device],
                      dtype: Optional[torch.dtype] = None) -> nn.Module:
    """Cast/move chunk tensors one at a time onto ``device`` and install them in a meta-built
    stage.  Peak host memory = one tensor; replaces accelerate's
    ``load_checkpoint_and_dispatch`` (gptserver.py:674-676)."""
    moved: Dict[str, Any] = {}
    for k in list(state_dict.keys()):
        t = state_dict.pop(k)
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        moved[k] = t.to(device)
    init_from_state_dict(model, moved)
    # non-persistent buffers (rope tables) were built on cpu by the RopeMixin setter
    for name, buf in list(model.named_buffers()):
        if buf.device != torch.device(device) and buf.device.type != "meta":
            owner = model
            *path, leaf = name.split(".")
            for p in path:
                owner = getattr(owner, p)
            setattr(owner, leaf, buf.to(device))
    return model


class IncrementalSaver:
    """Stream tensors to disk while a large state dict is being assembled.

    ``store_early(t)`` parks ``t`` in a file-backed tensor (RAM stays flat); ``save(sd)``
    writes the final ``.pth``.  Same contract as litGPT's ``incremental_save``
    (litgpt_utils.py:304-343) on a simpler mechanism.
    """

    def __init__(self, name: PathLike) -> None:
        self.name = str(name)
        self._tmp = tempfile.TemporaryDirectory(prefix="mdi_incsave_", dir=os.path.dirname(self.name) or ".")
        self._n = 0

    def store_early(self, tensor: torch.Tensor) -> torch.Tensor:
        if tensor.numel() == 0:
            return tensor
        t = tensor.detach().contiguous().cpu()
        fn = os.path.join(self._tmp.name, f"t{self._n}.bin")
        self._n += 1
        view_dtype = torch.uint8
        nbytes = t.numel() * t.element_size()
        backing = torch.from_file(fn, shared=True, size=nbytes, dtype=view_dtype)
        backing.copy_(t.view(torch.uint8).reshape(-1) if t.dim() else t.reshape(1).view(torch.uint8))
        return backing.view(t.dtype).reshape(t.shape)

    def save(self, obj: Any) -> None:
        torch.save(obj, self.name)

    def close(self) -> None:
        self._tmp.cleanup()


@contextmanager
def incremental_save(name: PathLike) -> Iterator[IncrementalSaver]:
    saver = IncrementalSaver(name)
    try:
        yield saver
    finally:
        saver.close()


def random_state_dict(config: Config, dtype: torch.dtype = torch.bfloat16, seed: int = 1234,
                      std: float = 0.02) -> Dict[str, torch.Tensor]:
    """litGPT-keyed random weights for ``config`` without instantiating modules."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape: int) -> torch.Tensor:
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    c, v = config.n_embd, config.padded_vocab_size
    sd: Dict[str, torch.Tensor] = {"transformer.wte.weight": rnd(v, c)}
    if config.pos_embedding == "learned":
        sd["transformer.wpe.weight"] = rnd(config.block_size, c)
    layer_norm = config.norm_class_name == "LayerNorm"

    def norm(prefix: str) -> None:
        sd[f"{prefix}.weight"] = (torch.ones(c) + 0.1 * torch.randn(c, generator=g)).to(dtype)
        if layer_norm:
            sd[f"{prefix}.bias"] = rnd(c)

    def linear(prefix: str, out_f: int, in_f: int, bias: bool) -> None:
        sd[f"{prefix}.weight"] = rnd(out_f, in_f)
        if bias:
            sd[f"{prefix}.bias"] = rnd(out_f)

    for l in range(config.n_layer):
        p = f"transformer.h.{l}"
        norm(f"{p}.norm_1")
        linear(f"{p}.attn.attn", config.qkv_size, c, config.bias)
        linear(f"{p}.attn.proj", c, config.attn_out_dim, config.bias)
        if not config.shared_attention_norm:
            norm(f"{p}.norm_2")
        i = config.intermediate_size
        if config.mlp_class_name in ("LLaMAMLP", "GemmaMLP"):
            linear(f"{p}.mlp.fc_1", i, c, config.bias)
            linear(f"{p}.mlp.fc_2", i, c, config.bias)
            linear(f"{p}.mlp.proj", c, i, config.bias)
        elif config.mlp_class_name == "LLaMAMoE":
            linear(f"{p}.mlp.gate", config.n_expert, c, False)
            for e in range(config.n_expert):
                linear(f"{p}.mlp.experts.{e}.fc_1", i, c, config.bias)
                linear(f"{p}.mlp.experts.{e}.fc_2", i, c, config.bias)
                linear(f"{p}.mlp.experts.{e}.proj", c, i, config.bias)
        else:
            linear(f"{p}.mlp.fc", i, c, config.bias)
            linear(f"{p}.mlp.proj", c, i, config.bias)
    norm("transformer.ln_f")
    if config.tie_embeddings:
        sd["lm_head.weight"] = sd["transformer.wte.weight"]
    else:
        linear("lm_head", v, c, config.lm_head_bias)
    if config.tie_embeddings and config.lm_head_bias:
        sd["lm_head.bias"] = rnd(v)
    return sd


def write_random_checkpoint(
    checkpoint_dir: PathLike,
    config: Union[Config, str],
    dtype: torch.dtype = torch.bfloat16,
    seed: int = 1234,
    **config_overrides: Any,
) -> Path:
    """Create ``<dir>/lit_model.pth`` + ``model_config.yaml`` with random-init weights."""
    if isinstance(config, str):
        config = Config.from_name(config, **config_overrides)
    out = Path(checkpoint_dir)
    out.mkdir(parents=True, exist_ok=True)
    torch.save(random_state_dict(config, dtype=dtype, seed=seed), out / "lit_model.pth")
    config.save(out)
    return out


def _key_seed(seed: int, key: str) -> int:
    import zlib

    return (int(seed) * 1000003 + zlib.crc32(key.encode())) & (2 ** 62 - 1)


def random_init_stage_(model: nn.Module, device: Union[str, torch.device], dtype: torch.dtype = torch.bfloat16,
                       seed: int = 1234, std: float = 0.02, layer_offset: int = 0) -> nn.Module:
    """Materialise a meta-built stage directly on ``device`` with random weights (no checkpoint
    I/O — benchmarks on the network-less GPU box).  Norm weights ~ 1 ± 0.1, the rest N(0, std).

    Every tensor is drawn from its own generator seeded by ``(seed, global parameter name)`` — the local
    block index shifted by ``layer_offset`` — so the SAME model comes out whatever the partition: the N
    stages of a pipeline and a single-stage copy built with the same seed hold identical weights (what the
    token-exactness checks of ``bench.py`` and the tests rely on)."""
    g = torch.Generator(device=device)
    for key, (mod, attr) in get_keys_to_submodule(model).items():
        old = getattr(mod, attr)
        if old.device.type != "meta":
            continue
        gkey = key
        if key.startswith("transformer.h."):
            _, _, li, tail = key.split(".", 3)
            gkey = f"transformer.h.{int(li) + layer_offset}.{tail}"
        g.manual_seed(_key_seed(seed, gkey))
        if "norm" in key or "ln_f" in key:
            t = (1.0 + 0.1 * torch.randn(old.shape, generator=g, device=device, dtype=torch.float32))
            if key.endswith(".bias"):
                t = t - 1.0
        else:
            t = torch.randn(old.shape, generator=g, device=device, dtype=torch.float32) * std
        setattr(mod, attr, nn.Parameter(t.to(dtype), requires_grad=False))
    if getattr(getattr(model, "config", None), "tie_embeddings", False) and hasattr(model, "lm_head"):
        model.lm_head.weight = model.transformer.wte.weight
    for name, buf in list(model.named_buffers()):
        owner = model
        *path, leaf = name.split(".")
        for p_ in path:
            owner = getattr(owner, p_)
        if buf.device.type != "meta":
            setattr(owner, leaf, buf.to(device))
    return model
