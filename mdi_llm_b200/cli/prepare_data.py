#!/usr/bin/env python3
"""Tokenise a text file into ``train.bin`` / ``val.bin`` (uint16 memmaps) for the trainer.

Parity: reference ``src/prepare_data.py`` (:18-51): ``DATA --tokenizer <dir>`` → 90/10 split.
``--tokenizer`` may be ``char`` / ``bpe[:vocab]`` to *train* one of the built-in tokenizers on the
data (legacy ``old/GPT2/prepare_data.py`` behaviour) and save it next to the bins.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("DATA", type=Path, help="text file (.txt/.md/.tex), or a directory holding one .txt file (old/GPT2/prepare_data.py:24)")
    p.add_argument("-t", "--tokenizer", type=str, default="character",
                   help="tokenizer directory, or 'char' / 'character' / 'bpe[:vocab_size]' / 'bytes'")
    p.add_argument("--vocab-size", type=int, default=500, help="vocabulary size of a custom BPE tokenizer ('-t bpe')")
    p.add_argument("--frac-train", type=float, default=0.9)
    p.add_argument("--out-dir", type=Path, default=None)
    a = p.parse_args(argv)
    from ..text.simple_tokenizers import BPETokenizer, CharacterTokenizer
    from ..text.tokenizer import Tokenizer, write_bytes_tokenizer
    from ..utils.data_loader import load_dataset, split_dataset

    if a.DATA.is_dir():  # the older generations pass the data-set folder: its only .txt file is the corpus
        txt = sorted(a.DATA.glob("*.txt"))
        if not txt:
            raise FileNotFoundError(f"no .txt file in {a.DATA}")
        a.DATA = txt[0]
    out = a.out_dir or a.DATA.parent
    out.mkdir(parents=True, exist_ok=True)
    spec = a.tokenizer
    if spec.lower() == "gpt2" and not Path(spec).is_dir():
        raise SystemExit("'-t gpt2' needs the GPT-2 tokenizer files: pass the directory that holds them (tokenizer.json / vocab)")
    if spec.lower() in ("char", "character"):
        t = CharacterTokenizer()
        t.tokenize(a.DATA.read_text(encoding="utf-8"))
        t.save(out)
        tok = Tokenizer(out, force_backend="char")
    elif spec.startswith("bpe"):
        vocab = int(spec.split(":")[1]) if ":" in spec else a.vocab_size
        t = BPETokenizer()
        t.tokenize(a.DATA.read_text(encoding="utf-8"), vocab)
        t.store_tokenizer_info(out, overwrite=True)
        tok = Tokenizer(out, force_backend="bpe")
    elif spec == "bytes":
        write_bytes_tokenizer(out)
        tok = Tokenizer(out, force_backend="bytes")
    else:
        tok = Tokenizer(Path(spec))
    data = load_dataset(a.DATA, tok)
    if tok.vocab_size > 65535:
        raise ValueError("vocabulary does not fit uint16 bins")
    train, val = split_dataset(data, a.frac_train)
    np.asarray(train.cpu(), dtype=np.uint16).tofile(out / "train.bin")
    np.asarray(val.cpu(), dtype=np.uint16).tofile(out / "val.bin")
    print(f"train: {len(train)} tokens, val: {len(val)} tokens, vocab {tok.vocab_size} -> {out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
