#!/usr/bin/env python3
"""Sanity-check a litGPT checkpoint directory.

Parity: reference ``src/scripts/inspect_lit.py`` (:46-99): print the config and dtype, verify the
number of transformer blocks found in ``lit_model.pth`` equals ``config.n_layer`` (raises
otherwise), dump the key list; also lists existing chunk splits.  ``--tokenizer`` exercises the
tokenizer round trip (what ``scripts/test_tok.py`` intends to do; that script is broken upstream).
"""
from __future__ import annotations

import argparse
from pathlib import Path


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("ckpt", type=Path, nargs="?", default=None)
    p.add_argument("--model", type=str, default=None,
                   help="the reference's form (inspect_lit.py:112): a local checkpoint folder, or a Hugging Face repo id that is "
                        "downloaded and converted into --ckpt-folder first")
    p.add_argument("--ckpt-folder", type=Path, default=Path("checkpoints"))
    p.add_argument("--device", type=str, default="cpu", help="accepted for compatibility (tensors are memory-mapped)")
    p.add_argument("-s", "--save", action="store_true", help="write the parameter names to tmp/<model>_params_keys_lit.txt")
    p.add_argument("--keys", action="store_true", help="print every parameter key and shape")
    p.add_argument("--tokenizer", type=str, default=None, help="encode/decode this text with the checkpoint's tokenizer")
    a = p.parse_args(argv)
    from ..models.partition import count_transformer_blocks
    from ..utils.checkpoint import lazy_load, load_from_pt

    if a.ckpt is None:
        if a.model is None:
            p.error("give the checkpoint directory (positional) or --model")
        a.ckpt = Path(a.model)
        if not a.ckpt.is_dir():  # a hub id: fetch + convert like the reference's load_from_hf
            from ..utils.checkpoint import load_from_hf

            print(f"Loading pretrained model {a.model} from Huggingface")
            load_from_hf(a.model, checkpoint_dir=a.ckpt_folder)
            a.ckpt = a.ckpt_folder / a.model
    cfg, _ = load_from_pt(a.ckpt, config_only=True)
    print("Model config:")
    for k, v in cfg.asdict().items():
        print(f"  {k}: {v}")
    sd = lazy_load(a.ckpt / "lit_model.pth")
    n = count_transformer_blocks(sd)
    dtypes = sorted({str(v.dtype) for v in sd.values()})
    n_params = sum(v.numel() for v in sd.values())
    print(f"\n{len(sd)} tensors, {n_params / 1e6:.1f} M parameters, dtypes {dtypes}, {n} transformer blocks")
    if n != cfg.n_layer:
        raise ValueError(f"The number of detected layers ({n}) is different from the config ({cfg.n_layer})")
    if a.keys:
        for k, v in sd.items():
            print(f"  {k}: {tuple(v.shape)}")
    if a.save:
        keys_file = Path("tmp") / f"{a.ckpt.name}_params_keys_lit.txt"
        keys_file.parent.mkdir(parents=True, exist_ok=True)
        keys_file.write_text("".join(f"{k}\n" for k in sd))
        print(f"Writing keys of Lit model to {keys_file}")
    chunks = a.ckpt / "chunks"
    if chunks.is_dir():
        for d in sorted(chunks.iterdir()):
            parts = {f.name: count_transformer_blocks(lazy_load(f)) for f in sorted(d.glob("*.pth"))}
            print(f"  split {d.name}: {parts}")
    if a.tokenizer is not None:
        from ..text.tokenizer import Tokenizer

        tok = Tokenizer(a.ckpt)
        ids = tok.encode(a.tokenizer)
        print(f"tokenizer backend {tok.backend}, vocab {tok.vocab_size}, bos {tok.bos_id}, eos {tok.eos_id}")
        print(f"  encode -> {ids.tolist()}\n  decode -> {tok.decode(ids)!r}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
