"""Tokenizer front-end over several backends.

Parity: reference ``src/sub/tokenizer.py:11-149`` — SentencePiece (``tokenizer.model``) or HF
``tokenizers`` (``tokenizer.json``), ``force_backend``, BOS/EOS discovery from
``tokenizer_config.json`` / ``generation_config.json``, ``encode(str, device, bos, eos,
max_length) -> int32 tensor``, ``decode``, ``token_to_id``, ``vocab_size``.

Added backends (the GPU box has no network, so real vocabularies are often absent):
``"bytes"`` — a tokenizer-free UTF-8 byte vocabulary (256 bytes + BOS/EOS/PAD) selected by a
``tokenizer_bytes.json`` marker, and ``"char"`` / ``"bpe"`` — the character and byte-pair
tokenizers of the legacy trees (``old/GPT2/sub/char_tokenizer.py``, ``bpe_tokenizer.py``)
rebuilt in :mod:`mdi_llm_b200.text.simple_tokenizers`.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import torch

__all__ = ["Tokenizer", "write_bytes_tokenizer"]

_BACKENDS = ("sentencepiece", "huggingface", "bytes", "char", "bpe")


class _SentencePiece:
    name = "sentencepiece"

    def __init__(self, d: Path) -> None:
        from sentencepiece import SentencePieceProcessor

        self.sp = SentencePieceProcessor(model_file=str(d / "tokenizer.model"))
        self.bos_id: Optional[int] = self.sp.bos_id()
        self.eos_id: Optional[int] = self.sp.eos_id()

    @staticmethod
    def present(d: Path) -> bool:
        return (d / "tokenizer.model").is_file()

    def vocab_size(self) -> int:
        return self.sp.vocab_size()

    def token_to_id(self, token: str) -> Optional[int]:
        return self.sp.piece_to_id(token)

    def encode(self, s: str) -> List[int]:
        return self.sp.encode(s)

    def decode(self, ids: Sequence[int]) -> str:
        return self.sp.decode(list(ids))


class _HuggingFace:
    name = "huggingface"

    def __init__(self, d: Path) -> None:
        from tokenizers import Tokenizer as HFTokenizer

        self.tk = HFTokenizer.from_file(str(d / "tokenizer.json"))
        self.bos_id = self.eos_id = None
        cfg = _read_json(d / "tokenizer_config.json")
        for attr, key in (("bos_id", "bos_token"), ("eos_id", "eos_token")):
            tok = cfg.get(key)
            if isinstance(tok, dict):  # {"content": "<s>", ...} form
                tok = tok.get("content")
            if tok is not None:
                setattr(self, attr, self.tk.token_to_id(tok))
        gen = _read_json(d / "generation_config.json")
        if self.bos_id is None:
            self.bos_id = gen.get("bos_token_id")
        if self.eos_id is None:
            eos = gen.get("eos_token_id")
            self.eos_id = eos[0] if isinstance(eos, list) else eos

    @staticmethod
    def present(d: Path) -> bool:
        return (d / "tokenizer.json").is_file()

    def vocab_size(self) -> int:
        return self.tk.get_vocab_size(with_added_tokens=False)

    def token_to_id(self, token: str) -> Optional[int]:
        return self.tk.token_to_id(token)

    def encode(self, s: str) -> List[int]:
        return self.tk.encode(s).ids

    def decode(self, ids: Sequence[int]) -> str:
        return self.tk.decode(list(ids))


class _Bytes:
    """UTF-8 bytes as tokens: ids 0..255 are bytes, then BOS, EOS, PAD."""

    name = "bytes"
    BOS, EOS, PAD = 256, 257, 258

    def __init__(self, d: Optional[Path] = None) -> None:
        self.bos_id, self.eos_id = self.BOS, self.EOS

    @staticmethod
    def present(d: Path) -> bool:
        return (d / "tokenizer_bytes.json").is_file()

    def vocab_size(self) -> int:
        return 259

    def token_to_id(self, token: str) -> Optional[int]:
        special = {"<bos>": self.BOS, "<eos>": self.EOS, "<pad>": self.PAD}
        if token in special:
            return special[token]
        b = token.encode("utf-8")
        return b[0] if len(b) == 1 else None

    def encode(self, s: str) -> List[int]:
        return list(s.encode("utf-8"))

    def decode(self, ids: Sequence[int]) -> str:
        return bytes(i for i in ids if 0 <= i < 256).decode("utf-8", errors="replace")


def _read_json(p: Path) -> Dict[str, Any]:
    if not p.is_file():
        return {}
    with open(p, encoding="utf-8") as fp:
        return json.load(fp)


def write_bytes_tokenizer(checkpoint_dir: Union[str, Path]) -> Path:
    """Drop the marker that makes :class:`Tokenizer` use the byte vocabulary for this dir."""
    p = Path(checkpoint_dir) / "tokenizer_bytes.json"
    p.parent.mkdir(parents=True, exist_ok=True)
    p.write_text(json.dumps({"type": "bytes", "vocab_size": 259, "bos_id": 256, "eos_id": 257}))
    return p


class Tokenizer:
    def __init__(self, checkpoint_dir: Union[Path, str], force_backend: Optional[str] = None) -> None:
        if force_backend is not None and force_backend not in _BACKENDS:
            raise AssertionError(f"Unsupported backend: {force_backend}")
        d = Path(checkpoint_dir)
        if not d.is_dir():
            d = d.parent
        if not d.exists():
            raise NotADirectoryError(f"The checkpoint directory does not exist: {str(d)}")
        self.use_bos = self.check_if_bos_token_used(d)

        from .simple_tokenizers import BPEBackend, CharBackend

        # `.model` takes precedence over `.json` when both exist (reference behaviour)
        order = [_SentencePiece, _HuggingFace, _Bytes, CharBackend, BPEBackend]
        chosen = None
        for cls in order:
            if force_backend is not None and cls.name != force_backend:
                continue
            if cls.present(d):
                chosen = cls(d)
                break
        if chosen is None:
            if force_backend:
                raise FileNotFoundError("Unable to find the configuration for the desired tokenizer")
            raise NotImplementedError("No supported tokenizer found")
        self.processor = chosen
        self.backend = chosen.name
        self.bos_id, self.eos_id = chosen.bos_id, chosen.eos_id
        if self.backend == "bytes":
            self.use_bos = True

    @property
    def vocab_size(self) -> int:
        return self.processor.vocab_size()

    def token_to_id(self, token: str) -> int:
        id_ = self.processor.token_to_id(token)
        if id_ is None:
            raise ValueError(f"token {token!r} not found in the collection.")
        return id_

    def check_if_bos_token_used(self, checkpoint_dir: Path) -> bool:
        cfg = _read_json(checkpoint_dir / "tokenizer_config.json")
        if not cfg:
            return False
        if "add_bos_token" in cfg:
            return bool(cfg["add_bos_token"])
        return cfg.get("tokenizer_class") == "LlamaTokenizer"

    def encode(
        self,
        string: str,
        device: Optional[torch.device] = None,
        bos: Optional[bool] = None,
        eos: bool = False,
        max_length: int = -1,
    ) -> torch.Tensor:
        tokens = list(self.processor.encode(string))
        if bos or (bos is None and self.use_bos):
            if self.bos_id is None:
                raise NotImplementedError("This tokenizer does not have a defined a bos token")
            tokens = [self.bos_id] + tokens
        if eos:
            tokens = tokens + [self.eos_id]
        if max_length > 0:
            tokens = tokens[:max_length]
        return torch.tensor(tokens, dtype=torch.int, device=device)

    def decode(self, tensor: torch.Tensor) -> str:
        ids = tensor.reshape(-1).tolist() if isinstance(tensor, torch.Tensor) else list(tensor)
        return self.processor.decode(ids)
