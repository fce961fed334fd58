#!/usr/bin/env python3
"""OpenWebText → ``train.bin`` / ``val.bin`` with multi-process tokenisation.

Parity: reference ``src/prepare_owt.py`` (:20-70): HF ``datasets`` ``openwebtext``, 0.05 % held-out
split, tokenise with ``--tokenizer`` in ``--num-proc`` workers, concatenate into uint16 memmaps in
1024 shards.  ``datasets`` and network access are required (neither exists on the GPU box):
``--text-dir`` tokenises local ``.txt`` files through the same sharded writer instead.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

import numpy as np


def write_sharded(token_lists, total_len: int, path: Path, n_shards: int = 1024) -> None:
    arr = np.memmap(path, dtype=np.uint16, mode="w+", shape=(total_len,))
    idx = 0
    n = len(token_lists)
    n_shards = max(1, min(n_shards, n))
    for s in range(n_shards):
        chunk = token_lists[s * n // n_shards: (s + 1) * n // n_shards]
        if not chunk:
            continue
        flat = np.concatenate([np.asarray(c, dtype=np.uint16) for c in chunk])
        arr[idx: idx + len(flat)] = flat
        idx += len(flat)
    arr.flush()


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("TOKENIZER_PATH", nargs="?", type=Path, default=None,
                   help="directory containing the tokenizer (the reference's positional form, prepare_owt.py:76)")
    p.add_argument("--tokenizer", type=Path, default=None, help="the same, as an option")
    p.add_argument("--out-dir", "--data-path", dest="out_dir", type=Path, default=Path("data/openwebtext"))
    p.add_argument("--num-proc", "--nproc", dest="num_proc", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    p.add_argument("--device", default=None, help="accepted for compatibility (tokenisation runs on the CPU)")
    p.add_argument("--text-dir", type=Path, default=None, help="tokenise local .txt files instead of downloading")
    p.add_argument("--val-fraction", type=float, default=0.0005)
    a = p.parse_args(argv)
    a.tokenizer = a.tokenizer or a.TOKENIZER_PATH
    if a.tokenizer is None:
        p.error("give the tokenizer directory (positional TOKENIZER_PATH or --tokenizer)")
    from ..text.tokenizer import Tokenizer

    tok = Tokenizer(a.tokenizer)
    a.out_dir.mkdir(parents=True, exist_ok=True)

    def encode(text: str):
        ids = tok.encode(text, eos=tok.eos_id is not None).tolist()
        return ids

    if a.text_dir is not None:
        docs = [f.read_text(encoding="utf-8") for f in sorted(a.text_dir.glob("*.txt"))]
        n_val = max(1, int(len(docs) * a.val_fraction)) if len(docs) > 1 else 0
        splits = {"val": docs[:n_val], "train": docs[n_val:]}
        for name, texts in splits.items():
            toks = [encode(t) for t in texts]
            write_sharded(toks, sum(map(len, toks)), a.out_dir / f"{name}.bin")
            print(f"{name}: {sum(map(len, toks))} tokens")
        return 0
    try:
        from datasets import load_dataset
    except ImportError as e:
        raise SystemExit("the `datasets` package (and network access) is needed for OpenWebText; use --text-dir for local text") from e
    ds = load_dataset("openwebtext", num_proc=a.num_proc)
    split = ds["train"].train_test_split(test_size=a.val_fraction, seed=2357, shuffle=True)
    split["val"] = split.pop("test")
    tokenised = split.map(lambda ex: {"ids": encode(ex["text"]), "len": 0}, remove_columns=["text"], num_proc=a.num_proc)
    for name, d in tokenised.items():
        lists = d["ids"]
        write_sharded(lists, sum(map(len, lists)), a.out_dir / f"{name}.bin")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
