#!/usr/bin/env python3
"""Inspect (and optionally fix / split-check) a checkpoint of any generation.

Parity: the reference's legacy tooling — ``old/GPT2/inspect_pretrained.py`` (key lists of a local
``.pt`` or of an HF GPT-2 flavour, optional save), ``old/GPT2/model_surgery.py`` (repair the
``config["DATASET"]`` field from the folder name) and ``old/nanoGPT/test_checkpoint.py`` (size in
RAM, split the parameters and round-trip the chunks through the wire serialisation).

Understands: a litGPT directory / ``lit_model.pth``; this repo's trainer output
(``ckpt_model.pth`` + pickled ``ckpt_state.pkl``); legacy single-file training checkpoints
(dict with ``model`` / ``model_args`` / ``config``); an HF ``gpt2*`` directory (converted on the fly).
"""
from __future__ import annotations

import argparse
import pickle
from pathlib import Path
from typing import Any, Dict

import torch


def load_any(path: Path) -> Dict[str, Any]:
    """-> {"kind", "state_dict", "meta"}"""
    if path.is_dir():
        if (path / "lit_model.pth").is_file():
            from ..utils.checkpoint import load_from_pt

            cfg, sd = load_from_pt(path)
            return {"kind": "litgpt", "state_dict": sd, "meta": {"config": cfg.asdict()}}
        if (path / "config.json").is_file():  # HF layout
            from ..utils.convert_hf_checkpoint import convert_hf_checkpoint
            from ..utils.checkpoint import load_from_pt

            convert_hf_checkpoint(checkpoint_dir=path)
            cfg, sd = load_from_pt(path)
            return {"kind": "hf-converted", "state_dict": sd, "meta": {"config": cfg.asdict()}}
        raise FileNotFoundError(f"{path}: neither lit_model.pth nor config.json")
    if path.suffix == ".pkl":
        with open(path, "rb") as f:
            st = pickle.load(f)
        return {"kind": "trainer-state", "state_dict": {}, "meta": st}
    obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict):
        meta = {k: v for k, v in obj.items() if k != "model"}
        return {"kind": "legacy-training", "state_dict": obj["model"], "meta": meta, "raw": obj}
    return {"kind": "state-dict", "state_dict": obj, "meta": {}}


def surgery_fix_dataset(path: Path, loaded: Dict[str, Any]) -> bool:
    """``config["DATASET"]`` must name the dataset folder the checkpoint sits in (``<data>/<set>/out/ckpt.pt``)."""
    raw = loaded.get("raw")
    if not raw or "config" not in raw or "DATASET" not in raw["config"]:
        return False
    want = path.resolve().parent.parent.name
    if raw["config"]["DATASET"] == want:
        return False
    raw["config"]["DATASET"] = want
    torch.save(raw, path)
    return True


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description=__doc__)
    p.add_argument("model", type=Path, help="checkpoint file or directory")
    p.add_argument("--keys", action="store_true", help="print every parameter name / shape / dtype")
    p.add_argument("--save-keys", type=Path, default=None, help="write the key list to this file")
    p.add_argument("--split", type=int, default=None, metavar="N", help="split for N nodes and round-trip the chunks")
    p.add_argument("--fix-dataset", action="store_true", help="model surgery: repair config['DATASET']")
    a = p.parse_args(argv)

    from ..models.partition import count_transformer_blocks, split_parameters
    from ..utils.misc import deserialize_params, get_obj_size, serialize_params

    loaded = load_any(a.model)
    sd = loaded["state_dict"]
    print(f"kind: {loaded['kind']}")
    for k, v in loaded["meta"].items():
        if isinstance(v, dict):
            print(f"{k}:")
            for kk, vv in v.items():
                print(f"\t{kk}: {vv if not isinstance(vv, dict) else '{...}'}")
        else:
            print(f"{k}: {v if not hasattr(v, 'keys') else '{...}'}")
    if sd:
        n_par = sum(int(t.numel()) for t in sd.values() if hasattr(t, "numel"))
        dtypes = sorted({str(t.dtype) for t in sd.values() if hasattr(t, "dtype")})
        print(f"{len(sd)} tensors, {n_par / 1e6:.2f} M parameters, dtypes {dtypes}, {get_obj_size(sd) / 2**20:.1f} MiB in RAM")
        try:
            print(f"transformer blocks: {count_transformer_blocks(sd)}")
        except Exception:  # noqa: BLE001  (legacy key names)
            pass
        lines = [f"{k}\t{tuple(v.shape)}\t{v.dtype}" for k, v in sd.items() if hasattr(v, "shape")]
        if a.keys:
            print("\n".join(lines))
        if a.save_keys:
            a.save_keys.parent.mkdir(parents=True, exist_ok=True)
            a.save_keys.write_text("\n".join(lines) + "\n")
    if a.fix_dataset:
        print("Ckpt was updated (fixed dataset name)!" if surgery_fix_dataset(a.model, loaded) else "dataset name ok")
    if a.split:
        chunks, info = split_parameters(dict(sd), a.split)
        print(f"split plan {info['plan']}")
        for name, c in [("starter", chunks["starter"])] + [(f"secondary{i}", c) for i, c in enumerate(chunks["secondary"])]:
            wire = serialize_params(c)
            back = deserialize_params(wire)
            ok = all(torch.equal(back[k], c[k]) for k in c)
            print(f"  {name}: {len(c)} tensors, {get_obj_size(c) / 2**20:.1f} MiB, wire round-trip {'ok' if ok else 'MISMATCH'}")
            if not ok:
                return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
