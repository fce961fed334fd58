#!/usr/bin/env python3
"""Single-device generation — the "1 node" baseline of the tokens-vs-time plots.

Parity: reference ``src/sample.py`` (:27-276; flags :287-356): load config + ``lit_model.pth``,
dtype inference, generate ``--n-samples`` sequentially (no batching, "would not be fair compared
to MDI", :131-133) with per-token timestamps, CSV/PNG with ``-p``, cProfile with ``-d``.
On CUDA with a supported architecture the decode runs on the fused sm_100a kernels
(``DevicePipeline`` with one stage); ``--engine eager`` forces plain PyTorch.
"""
from __future__ import annotations

import argparse
import cProfile
import pstats
import time
from pathlib import Path

from .common import IMG_DIR, LOGS_DIR, append_run_stats, seed_everything, tokens_time_csv_name


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Single-device sampling")
    p.add_argument("--ckpt", type=Path, required=True, help="checkpoint folder (lit_model.pth + model_config.yaml)")
    p.add_argument("--prompt", type=str, default="Who are you?")
    p.add_argument("--n-samples", type=int, default=1)
    p.add_argument("--n-tokens", type=int, default=300)
    p.add_argument("--sequence-length", "--context-length", "--block-size", dest="sequence_length", type=int, default=None)
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--dtype", type=str, default=None)
    p.add_argument("--temperature", type=float, default=None)
    p.add_argument("--top-k", type=int, default=None)
    p.add_argument("--greedy", action="store_true")
    p.add_argument("--engine", default="auto", choices=["auto", "eager", "cuda"])
    p.add_argument("--seed", type=int, default=10137)
    p.add_argument("--time-run", type=Path, default=None)
    p.add_argument("-p", "--plots", action="store_true")
    p.add_argument("-v", "--verb", action="store_true")
    p.add_argument("-d", "--debug", action="store_true", help="profile with cProfile -> logs/sample_profile.prof")
    p.add_argument("-c", "--compile", action="store_true", help="accepted for compatibility")
    return p


def run(args) -> int:
    import torch

    from .. import config as C
    from ..models.gpt import GPT
    from ..text.prompts import PromptStyle, get_user_prompt, has_prompt_style, load_prompt_style
    from ..text.tokenizer import Tokenizer, write_bytes_tokenizer
    from ..utils.checkpoint import load_from_pt
    from ..utils.misc import find_eot
    from ..utils.plots import plot_tokens_per_time, write_points_csv

    seed_everything(args.seed)
    device = args.device or C.default_device()
    dtype_name = args.dtype or (C.default_dtype() if "cuda" in device else "float32")
    dtype = C.DTYPE_TORCH_MAPPING[dtype_name]
    cfg, sd = load_from_pt(args.ckpt)
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
    prompts = [tok.encode(t) for t in get_user_prompt(args.prompt, args.n_samples, style)]
    greedy = args.greedy
    temperature = 0.0 if greedy else (args.temperature if args.temperature is not None else C.TEMPERATURE)
    top_k = None if greedy else (args.top_k if args.top_k is not None else C.TOP_K)
    seq_len = args.sequence_length or cfg.block_size
    if any(p.numel() + args.n_tokens > seq_len for p in prompts):
        raise ValueError(f"Cannot generate {args.n_tokens} tokens - would exceed block size!")

    use_cuda_engine = False
    if "cuda" in device and args.engine in ("auto", "cuda"):
        from ..parallel.engine import engine_supports

        use_cuda_engine = engine_supports(cfg, dtype)
        if args.engine == "cuda" and not use_cuda_engine:
            raise RuntimeError(f"the fused engine does not support {cfg.name} / {dtype_name}")
    tok_time, outputs = [], []
    t_start = time.time()
    if use_cuda_engine:
        from ..models.stage import build_stage
        from ..parallel.pipeline import DevicePipeline
        from ..parallel.scheduler import SamplingParams
        from ..utils.checkpoint import materialize_stage

        stage = build_stage(cfg, "starter", cfg.n_layer, meta=True)
        materialize_stage(stage, sd, device, dtype)
        sampling = SamplingParams.greedy() if greedy else SamplingParams(temperature=temperature, top_k=top_k, seed=args.seed)
        pipe = DevicePipeline(stage, 0, 1, n_samples=1, max_seq_length=seq_len, sampling=sampling)
        n_done = 0
        for p in prompts:  # sequential on purpose (reference sample.py:131-133)
            pipe.prepare([p], args.n_tokens)
            pipe.prefill()
            pipe.decode_rounds_host(args.n_tokens, on_token=lambda s, pos, t: tok_time.append((len(tok_time) + 1, time.time() - t_start)))
            outputs.append(pipe.result_tokens()[0])
            n_done += args.n_tokens
    else:
        model = GPT(cfg)
        model.load_state_dict(sd, strict=not cfg.tie_embeddings)
        model = model.to(device=device, dtype=dtype).eval()
        model.max_seq_length = seq_len
        model.set_kv_cache(1)
        for p in prompts:
            local: list = []
            out = model.generate(p.to(device), p.numel() + args.n_tokens, temperature=temperature, top_k=top_k,
                                 top_p=0.0 if greedy else 1.0, tok_time=local)
            base = tok_time[-1] if tok_time else (0, 0.0)
            tok_time.extend((base[0] + n + 1, base[1] + t) for n, t in local)
            outputs.append(out.cpu())
            model.kv_pool.reset()
    total = time.time() - t_start
    for i, (p, out) in enumerate(zip(prompts, outputs)):
        text = tok.decode(find_eot(out, stop_tokens, p.numel()))
        print("-------------------------------------------------")
        print(f"Sample {i + 1}:\n{text}\n")
    n_gen = args.n_tokens * len(prompts)
    print(f"Total generation time: {total:.3f} s ({n_gen / total:.1f} tokens/s, engine: {'cuda' if use_cuda_engine else 'eager'})")
    if args.plots and tok_time:
        name = tokens_time_csv_name(1, Path(args.ckpt).name, args.n_samples)
        write_points_csv(tok_time, LOGS_DIR / name)
        plot_tokens_per_time(tok_time, out_path=IMG_DIR / name.replace("tokens_time_samples_", "tokens_time_").replace(".csv", ".png"))
    if args.time_run is not None:
        append_run_stats(args.time_run, args.n_samples, cfg.n_layer, cfg.block_size, total)
    return 0


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.debug:
        LOGS_DIR.mkdir(parents=True, exist_ok=True)
        prof = cProfile.Profile()
        prof.enable()
        try:
            return run(args)
        finally:
            prof.disable()
            prof.dump_stats(str(LOGS_DIR / "sample_profile.prof"))
            pstats.Stats(prof).sort_stats("cumulative").print_stats(15)
    return run(args)


if __name__ == "__main__":
    raise SystemExit(main())
