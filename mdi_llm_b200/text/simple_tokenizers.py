"""Self-contained tokenizers that need no external vocabulary files.

Parity with the legacy trees of the reference: ``CharacterTokenizer``
(``old/GPT2/sub/char_tokenizer.py:6-60``: ``tokenize`` builds the char<->id maps, ``encode``,
``decode``) and ``BPETokenizer`` (``old/GPT2/sub/bpe_tokenizer.py:104-300``: byte-level BPE
trained with a GPT-style pre-tokenisation regex; ``tokenize(text, out_vocab_size)``,
``store_tokenizer_info`` / ``load_tokenizer_info``, ``encode``, ``decode``, ``trained``).

Both also plug into :class:`mdi_llm_b200.text.tokenizer.Tokenizer` as backends, detected by
``tokenizer_char.json`` / ``tokenizer_bpe.json`` in the checkpoint directory, which lets the
trainer (``cli/train.py``) and the prep scripts run end to end with no downloaded assets.
"""
from __future__ import annotations

import json
from collections import Counter
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

try:  # `regex` supports \p{L}; fall back to `re` with an ASCII-ish approximation
    import regex as _re

    _SPLIT = _re.compile(r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")
except ImportError:  # pragma: no cover
    import re as _re

    _SPLIT = _re.compile(r"""'(?:[sdmt]|ll|ve|re)| ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+(?!\S)|\s+""")

__all__ = ["CharacterTokenizer", "BPETokenizer", "CharBackend", "BPEBackend"]


class CharacterTokenizer:
    def __init__(self, stoi: Optional[Dict[str, int]] = None, itos: Optional[Dict[int, str]] = None) -> None:
        self.stoi: Dict[str, int] = dict(stoi or {})
        self.itos: Dict[int, str] = {int(k): v for k, v in (itos or {}).items()}
        if self.stoi and not self.itos:
            self.itos = {i: s for s, i in self.stoi.items()}
        if self.itos and not self.stoi:
            self.stoi = {s: i for i, s in self.itos.items()}

    @property
    def vocab_size(self) -> int:
        return len(self.stoi)

    def tokenize(self, text: Union[str, Iterable[str]]) -> None:
        """Build the vocabulary from all distinct characters of ``text``."""
        chars = sorted(set(text if isinstance(text, str) else "".join(text)))
        self.stoi = {c: i for i, c in enumerate(chars)}
        self.itos = dict(enumerate(chars))

    def encode(self, in_str: str) -> List[int]:
        return [self.stoi[c] for c in in_str if c in self.stoi]

    def decode(self, line: Iterable[int]) -> str:
        return "".join(self.itos[int(i)] for i in line if int(i) in self.itos)

    def save(self, d: Union[str, Path]) -> Path:
        p = Path(d) / "tokenizer_char.json"
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(json.dumps({"stoi": self.stoi}, ensure_ascii=False))
        return p

    @classmethod
    def load(cls, d: Union[str, Path]) -> "CharacterTokenizer":
        return cls(stoi=json.loads((Path(d) / "tokenizer_char.json").read_text())["stoi"])


class BPETokenizer:
    """Byte-level BPE.  Ids 0..255 are raw bytes; id 256+k is the k-th learned merge."""

    def __init__(self) -> None:
        self.merges: Dict[Tuple[int, int], int] = {}
        self.vocab: Dict[int, bytes] = {i: bytes([i]) for i in range(256)}
        self.n_vocab = 256

    def trained(self) -> bool:
        return bool(self.merges)

    @property
    def vocab_size(self) -> int:
        return self.n_vocab

    @staticmethod
    def _words(text: str) -> List[List[int]]:
        return [list(w.encode("utf-8")) for w in _SPLIT.findall(text)]

    @staticmethod
    def _merge(word: List[int], pair: Tuple[int, int], new_id: int) -> List[int]:
        out: List[int] = []
        i, n = 0, len(word)
        while i < n:
            if i + 1 < n and word[i] == pair[0] and word[i + 1] == pair[1]:
                out.append(new_id)
                i += 2
            else:
                out.append(word[i])
                i += 1
        return out

    def tokenize(self, text: str, out_vocab_size: int = 500) -> None:
        """Learn merges until the vocabulary reaches ``out_vocab_size``."""
        if out_vocab_size < 256:
            raise ValueError("out_vocab_size must be >= 256")
        # work on distinct pre-tokens weighted by frequency
        freq = Counter(tuple(w) for w in self._words(text))
        words: Dict[Tuple[int, ...], int] = dict(freq)
        self.merges.clear()
        self.vocab = {i: bytes([i]) for i in range(256)}
        next_id = 256
        while next_id < out_vocab_size:
            stats: Counter = Counter()
            for w, c in words.items():
                for pair in zip(w, w[1:]):
                    stats[pair] += c
            if not stats:
                break
            pair = max(stats.items(), key=lambda kv: (kv[1], -kv[0][0], -kv[0][1]))[0]
            self.merges[pair] = next_id
            self.vocab[next_id] = self.vocab[pair[0]] + self.vocab[pair[1]]
            merged: Dict[Tuple[int, ...], int] = {}
            for w, c in words.items():
                nw = tuple(self._merge(list(w), pair, next_id)) if len(w) > 1 else w
                merged[nw] = merged.get(nw, 0) + c
            words = merged
            next_id += 1
        self.n_vocab = next_id

    def build_mapping(self) -> None:
        self.vocab = {i: bytes([i]) for i in range(256)}
        for (a, b), idx in sorted(self.merges.items(), key=lambda kv: kv[1]):
            self.vocab[idx] = self.vocab[a] + self.vocab[b]
        self.n_vocab = 256 + len(self.merges)

    def encode(self, text: str) -> List[int]:
        out: List[int] = []
        for word in self._words(text):
            while len(word) > 1:
                cand = [(self.merges.get(p, 1 << 60), p) for p in zip(word, word[1:])]
                rank, pair = min(cand)
                if rank == 1 << 60:
                    break
                word = self._merge(word, pair, rank)
            out.extend(word)
        return out

    def decode(self, ids: Iterable[int]) -> str:
        return b"".join(self.vocab.get(int(i), b"") for i in ids).decode("utf-8", errors="replace")

    def store_tokenizer_info(self, info_dir: Union[str, Path], overwrite: bool = False) -> Path:
        p = Path(info_dir) / "tokenizer_bpe.json"
        if p.exists() and not overwrite:
            raise FileExistsError(f"{p} exists (pass overwrite=True)")
        p.parent.mkdir(parents=True, exist_ok=True)
        merges = [[a, b, idx] for (a, b), idx in sorted(self.merges.items(), key=lambda kv: kv[1])]
        p.write_text(json.dumps({"merges": merges}))
        return p

    def load_tokenizer_info(self, vocab_path: Union[str, Path], meta_path: Optional[Union[str, Path]] = None) -> None:
        p = Path(vocab_path)
        if p.is_dir():
            p = p / "tokenizer_bpe.json"
        self.merges = {(a, b): idx for a, b, idx in json.loads(p.read_text())["merges"]}
        self.build_mapping()


class CharBackend:
    name = "char"
    bos_id: Optional[int] = None
    eos_id: Optional[int] = None

    def __init__(self, d: Path) -> None:
        self.tk = CharacterTokenizer.load(d)
        self.eos_id = self.tk.stoi.get("\n")

    @staticmethod
    def present(d: Path) -> bool:
        return (d / "tokenizer_char.json").is_file()

    def vocab_size(self) -> int:
        return self.tk.vocab_size

    def token_to_id(self, token: str) -> Optional[int]:
        return self.tk.stoi.get(token)

    def encode(self, s: str) -> List[int]:
        return self.tk.encode(s)

    def decode(self, ids: Sequence[int]) -> str:
        return self.tk.decode(ids)


class BPEBackend:
    name = "bpe"
    bos_id: Optional[int] = None
    eos_id: Optional[int] = None

    def __init__(self, d: Path) -> None:
        self.tk = BPETokenizer()
        self.tk.load_tokenizer_info(d)
        self.eos_id = ord("\n")

    @staticmethod
    def present(d: Path) -> bool:
        return (d / "tokenizer_bpe.json").is_file()

    def vocab_size(self) -> int:
        return self.tk.vocab_size

    def token_to_id(self, token: str) -> Optional[int]:
        b = token.encode("utf-8")
        for idx, v in self.tk.vocab.items():
            if v == b:
                return idx
        return None

    def encode(self, s: str) -> List[int]:
        return self.tk.encode(s)

    def decode(self, ids: Sequence[int]) -> str:
        return self.tk.decode(ids)
