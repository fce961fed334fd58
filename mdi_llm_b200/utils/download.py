"""Download a model (or just its tokenizer) from the Hugging Face Hub and convert it to litGPT.

Parity: reference ``src/sub/utils/download.py:15-181`` — ``download_from_hub(repo_id,
access_token, tokenizer_only, convert_checkpoint, dtype, checkpoint_dir, model_name)``: picks
``.bin`` vs ``.safetensors`` files, optional ``hf_transfer``, maps gated-repo errors to a readable
message, then runs the HF→lit conversion.  Safetensors shards are read directly by the converter
here (the reference rewrites them as ``.bin`` first).  There is no network on the B200 box: the
function raises a clear error when the hub is unreachable.
"""
from __future__ import annotations

import os
from contextlib import contextmanager
from pathlib import Path
from typing import Iterator, List, Optional

from .imports import RequirementCache

__all__ = ["download_from_hub", "find_weight_files", "gated_repo_catcher"]

_HUB = RequirementCache("huggingface_hub")
_HF_TRANSFER = RequirementCache("hf_transfer")


def find_weight_files(repo_id: str, access_token: Optional[str]) -> tuple:
    from huggingface_hub import repo_info
    from huggingface_hub.utils import filter_repo_objects

    with gated_repo_catcher(repo_id, access_token):
        info = repo_info(repo_id, token=access_token)
    names = [f.rfilename for f in info.siblings]
    return list(filter_repo_objects(items=names, allow_patterns=["*.bin*"])), \
        list(filter_repo_objects(items=names, allow_patterns=["*.safetensors*"]))


@contextmanager
def gated_repo_catcher(repo_id: str, access_token: Optional[str]) -> Iterator[None]:
    try:
        yield
    except OSError as e:
        err = str(e)
        if "Repository Not Found" in err:
            raise ValueError(f"Repository at https://huggingface.co/api/models/{repo_id} not found. "
                             "Please make sure you specified the correct `repo_id`.") from None
        if "gated repo" in err:
            if not access_token:
                raise ValueError(f"https://huggingface.co/{repo_id} requires authentication, please set the `HF_TOKEN=your_token`"
                                 " environment variable or pass `--access_token=your_token`. You can find your token by visiting"
                                 " https://huggingface.co/settings/tokens.") from None
            raise ValueError(f"https://huggingface.co/{repo_id} requires authentication. The access token provided by `HF_TOKEN=your_token`"
                             " environment variable or `--access_token=your_token` may not have sufficient access rights. Please"
                             f" visit https://huggingface.co/{repo_id} for more information.") from None
        raise


def download_from_hub(
    repo_id: Optional[str] = None,
    access_token: Optional[str] = os.getenv("HF_TOKEN"),
    tokenizer_only: bool = False,
    convert_checkpoint: bool = True,
    dtype: Optional[str] = None,
    checkpoint_dir: Path = Path("checkpoints"),
    model_name: Optional[str] = None,
) -> Path:
    """Fetch ``repo_id`` into ``checkpoint_dir/repo_id`` and (by default) convert it."""
    from ..models.registry import configs

    if repo_id is None:
        options = [f"{c['hf_config']['org']}/{c['hf_config']['name']}" for c in configs]
        print("Please specify --repo_id <repo_id>. Available values:")
        print("\n".join(sorted(options, key=lambda x: x.lower())))
        return Path(checkpoint_dir)
    if not _HUB:
        raise ModuleNotFoundError(str(_HUB))
    from huggingface_hub import snapshot_download

    patterns: List[str] = ["tokenizer*", "generation_config.json", "config.json"]
    if not tokenizer_only:
        bins, safetensors = find_weight_files(repo_id, access_token)
        if bins:
            patterns += ["*.bin", "*.bin.index.json"]  # covers .bin.index.json as well
        elif safetensors:
            patterns += ["*.safetensors", "*.safetensors.index.json"]
        else:
            raise ValueError(f"Couldn't find weight files for {repo_id}")
    import huggingface_hub._snapshot_download as dl
    import huggingface_hub.constants as constants

    previous = constants.HF_HUB_ENABLE_HF_TRANSFER
    if _HF_TRANSFER and not previous:
        print("Setting HF_HUB_ENABLE_HF_TRANSFER=1")
        constants.HF_HUB_ENABLE_HF_TRANSFER = True
        dl.HF_HUB_ENABLE_HF_TRANSFER = True
    directory = Path(checkpoint_dir) / repo_id
    try:
        with gated_repo_catcher(repo_id, access_token):
            snapshot_download(repo_id, local_dir=directory, allow_patterns=patterns, token=access_token)
    finally:
        constants.HF_HUB_ENABLE_HF_TRANSFER = previous
        dl.HF_HUB_ENABLE_HF_TRANSFER = previous
    if convert_checkpoint and not tokenizer_only:
        from .convert_hf_checkpoint import convert_hf_checkpoint

        print("Converting checkpoint files to the litGPT format.")
        convert_hf_checkpoint(checkpoint_dir=directory, dtype=dtype, model_name=model_name)
    return directory
