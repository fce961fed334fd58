#!/usr/bin/env python3
"""Train a small BPE (or char) tokenizer on a text file and show an encode/decode round trip.

Parity: reference ``old/GPT2/test_tokenizer.py`` / ``old/nanoGPT/test_tokenizer.py`` (train the
repo's own ``BPETokenizer`` on Shakespeare with 500 merges, encode a line, decode it back).
"""
from __future__ import annotations

import argparse
from pathlib import Path

DATA = Path(__file__).resolve().parents[1] / "data" / "sonnets.txt"


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("PATH", nargs="?", type=Path, default=None,
                   help="folder containing tokenizer files: show the special tokens of a checkpoint's tokenizer and a round "
                        "trip instead (what the reference's scripts/test_tok.py:6-13 sets out to do)")
    p.add_argument("--text", type=Path, default=DATA)
    p.add_argument("--kind", choices=["bpe", "char"], default="bpe")
    p.add_argument("--vocab-size", type=int, default=500)
    p.add_argument("--sentence", default="O, that this too too solid flesh would melt")
    p.add_argument("--save", type=Path, default=None, help="store the trained tokenizer here")
    a = p.parse_args(argv)
    if a.PATH is not None:
        from ..text.tokenizer import Tokenizer

        ck = Tokenizer(a.PATH)
        for label, tid in (("Beginning of sentence", ck.bos_id), ("End of sentence", ck.eos_id)):
            print(f"{label}: {tid} -> {ck.decode([tid]) if tid is not None else None!r}")
        ids = ck.encode(a.sentence)
        back = ck.decode(ids)
        print(f"backend {ck.backend}, vocab {ck.vocab_size}\n{a.sentence!r} -> {ids.tolist()} -> {back!r}")
        return 0 if back.strip() == a.sentence.strip() else 1
    from ..text.simple_tokenizers import BPETokenizer, CharacterTokenizer

    text = a.text.read_text(encoding="utf-8")
    tok = BPETokenizer() if a.kind == "bpe" else CharacterTokenizer()
    if a.kind == "bpe":
        tok.tokenize(text, out_vocab_size=a.vocab_size)
    else:
        tok.tokenize(text)
    print(f"Encoding the string:\n{a.sentence}")
    enc = tok.encode(a.sentence)
    print(f"Encoded sequence:\n {enc}")
    dec = tok.decode(enc)
    print(f"\nDecoding\n    {dec}")
    if a.save:
        a.save.parent.mkdir(parents=True, exist_ok=True)
        tok.store_tokenizer_info(a.save) if a.kind == "bpe" else tok.save(a.save)
    return 0 if dec == a.sentence else 1


if __name__ == "__main__":
    raise SystemExit(main())
