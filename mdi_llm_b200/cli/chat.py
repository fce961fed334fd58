#!/usr/bin/env python3
"""Interactive chat REPL on a single device.

Parity: reference ``src/chat.py`` (:57-200): prompt-style aware REPL, streaming through
``GPT.generate_chat`` (stop-sequence hold-back buffer) and incremental re-decoding so that
tokenizers that merge spaces print correctly (:36-54).

On a CUDA device with an architecture the fused kernels cover (``--engine auto|cuda``) the reply is produced by
the same path as everything else: a one-stage :class:`~mdi_llm_b200.parallel.pipeline.DevicePipeline` in
host-fed mode (tcgen05 prefill, CUDA-graph decode steps, device sampler), one token read back per step and
pushed through the same hold-back logic.  ``--engine eager`` keeps the plain PyTorch model.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path
from typing import Iterator, List

from .common import seed_everything


def decode_stream(tokenizer, token_stream: Iterator, out=sys.stdout) -> int:
    """Print tokens as they arrive.  SentencePiece drops leading spaces when decoding a single
    token, so the whole suffix is re-decoded and only the new characters are written."""
    ids: List[int] = []
    printed = ""
    n = 0
    for t in token_stream:
        ids.append(int(t))
        text = tokenizer.decode(__import__("torch").tensor(ids))
        if len(text) > len(printed) and not text.endswith("�"):
            out.write(text[len(printed):])
            out.flush()
            printed = text
        n += 1
    return n


def stream_device(pipe, ids, max_new_tokens: int, stop_tokens) -> Iterator[int]:
    """``GPT.generate_chat`` on the fused engine: yields generated token ids in chunks of the longest stop sequence's
    length and returns, without the pending chunk, as soon as a stop sequence completes — the reference's buffering rule
    (model.py:526-573: the completing token is never printed; earlier tokens of the sequence may be)."""
    pipe.prepare([ids.cpu()], max_new_tokens)
    pipe.prefill()
    produced: List[int] = []
    emitted = 0
    hold = max((len(s) for s in stop_tokens), default=1)
    for t in range(1, max_new_tokens + 1):
        got: List[int] = []
        pipe.decode_rounds_host(1, on_token=lambda slot, pos, tok: got.append(tok))
        if not got:
            break
        produced.append(got[0])
        tail = produced[-hold:]
        if any(len(s) <= len(produced) and tail[-len(s):] == list(s) for s in stop_tokens):
            return
        if t - emitted >= hold:
            yield from produced[emitted:t]
            emitted = t
    yield from produced[emitted:]


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Chat with a litGPT-format model")
    p.add_argument("--ckpt", type=Path, required=True)
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--top-k", type=int, default=200)
    p.add_argument("--temperature", type=float, default=0.8)
    p.add_argument("--max-new-tokens", type=int, default=256)
    p.add_argument("--sequence-length", "--context-length", "--block-size", dest="sequence_length", type=int, default=None)
    p.add_argument("--seed", type=int, default=10137)
    p.add_argument("--once", type=str, default=None, help="answer this single prompt and exit (non-interactive)")
    p.add_argument("--engine", default="auto", choices=["auto", "eager", "cuda"], help="fused sm_100a engine or eager PyTorch")
    p.add_argument("--top-p", type=float, default=1.0)
    p.add_argument("-c", "--compile", action="store_true",
                   help="accepted for compatibility (chat.py:215): the fused engine needs no compilation step")
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    import torch

    from .. import config as C
    from ..models.gpt import GPT
    from ..text.prompts import PromptStyle, has_prompt_style, load_prompt_style
    from ..text.tokenizer import Tokenizer, write_bytes_tokenizer
    from ..utils.checkpoint import load_from_pt

    seed_everything(args.seed)
    device = args.device or C.default_device()
    dtype = C.DTYPE_TORCH_MAPPING[args.dtype or (C.default_dtype() if "cuda" in device else "float32")]
    cfg, sd = load_from_pt(args.ckpt)
    pipe = model = None
    if "cuda" in device and args.engine in ("auto", "cuda"):
        from ..parallel.engine import engine_supports

        if engine_supports(cfg, dtype):
            from ..models.stage import build_stage
            from ..parallel.pipeline import DevicePipeline
            from ..parallel.scheduler import SamplingParams
            from ..utils.checkpoint import materialize_stage

            stage = build_stage(cfg, "starter", cfg.n_layer, meta=True)
            materialize_stage(stage, sd, device, dtype)
            pipe = DevicePipeline(stage, 0, 1, n_samples=1, max_seq_length=args.sequence_length or cfg.block_size,
                                  sampling=SamplingParams(temperature=args.temperature, top_k=args.top_k, top_p=args.top_p,
                                                          seed=args.seed))
        elif args.engine == "cuda":
            raise RuntimeError(f"the fused engine does not support {cfg.name} / {dtype}")
    if pipe is None:
        model = GPT(cfg)
        model.load_state_dict(sd, strict=not cfg.tie_embeddings)
        model = model.to(device=device, dtype=dtype).eval()
        if args.sequence_length:
            model.max_seq_length = args.sequence_length
    max_seq = pipe.stage.S if pipe is not None else model.max_seq_length
    try:
        tok = Tokenizer(args.ckpt)
    except (NotImplementedError, FileNotFoundError):
        write_bytes_tokenizer(args.ckpt)
        tok = Tokenizer(args.ckpt, force_backend="bytes")
    style = load_prompt_style(args.ckpt) if has_prompt_style(args.ckpt) else PromptStyle.from_config(cfg)
    try:
        stop_tokens = style.stop_tokens(tok)
    except ValueError:
        stop_tokens = ([tok.eos_id],)
    print(f"Now chatting with {cfg.name}.\nTo exit, press 'Enter' on an empty prompt.\n")
    while True:
        try:
            prompt = args.once if args.once is not None else input(">> Prompt: ")
        except (KeyboardInterrupt, EOFError):
            break
        if not prompt:
            break
        ids = tok.encode(style.apply(prompt), device=torch.device(device))
        budget = min(max_seq, ids.numel() + args.max_new_tokens)
        t0 = time.perf_counter()
        print(">> Reply: ", end="")
        if pipe is not None:
            n = decode_stream(tok, stream_device(pipe, ids, budget - ids.numel(), stop_tokens))
        else:
            model.set_kv_cache(1)
            n = decode_stream(tok, model.generate_chat(ids, budget, temperature=args.temperature, top_k=args.top_k,
                                                       top_p=args.top_p, stop_tokens=stop_tokens))
            model.clear_kv_cache()
        dt = time.perf_counter() - t0
        print(f"\nTime for inference: {dt:.02f} sec total, {n / max(dt, 1e-9):.02f} tokens/sec", file=sys.stderr)
        print()
        if args.once is not None:
            break
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
