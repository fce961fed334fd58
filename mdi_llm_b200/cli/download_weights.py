#!/usr/bin/env python3
"""Download weights (or only the tokenizer) of a model from the Hugging Face hub.

Parity: reference ``src/download_weights.py`` (:10-67): ``MODEL [--dtype --hf-token --ckpt-dir
--saved-name --tokenizer-only --no-convert]``.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("MODEL", nargs="?", default=None, help="HF repo id (omit to list the supported ones)")
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--hf-token", type=str, default=os.getenv("HF_TOKEN"))
    p.add_argument("--ckpt-dir", type=Path, default=Path("checkpoints"))
    p.add_argument("--saved-name", "--model-name", dest="saved_name", type=str, default=None)
    p.add_argument("--tokenizer-only", action="store_true")
    p.add_argument("--no-convert", action="store_true")
    a = p.parse_args(argv)
    from ..utils.download import download_from_hub

    download_from_hub(repo_id=a.MODEL, access_token=a.hf_token, tokenizer_only=a.tokenizer_only,
                      convert_checkpoint=not a.no_convert, dtype=a.dtype, checkpoint_dir=a.ckpt_dir, model_name=a.saved_name)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
