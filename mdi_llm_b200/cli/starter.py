#!/usr/bin/env python3
"""Starter node (node 0) of a model-distributed inference run.

Parity: reference ``src/starter.py`` — same flags (:109-196): ``-d -v -p -c --ckpt --chunk
--nodes-config --device --prompt --n-samples --n-tokens --sequence-length|--context-length|
--block-size --dtype --time-run --seed``; writes the tokens/time CSV (+PNG when matplotlib is
available) with ``-p`` and appends run statistics with ``--time-run``.  Extra flags select the
B200 execution path: ``--engine`` (auto|eager|cuda), ``--greedy``, ``--temperature``, ``--top-k``,
``--partition``.
"""
from __future__ import annotations

import argparse
from pathlib import Path

from .common import IMG_DIR, LOGS_DIR, SETTINGS_DIR, append_run_stats, seed_everything, setup_debug_log, tokens_time_csv_name


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Starter node - MDI")
    p.add_argument("-d", "--debug", action="store_true", help="enable debug mode (file log)")
    p.add_argument("-v", "--verb", action="store_true", help="enable verbose mode")
    p.add_argument("-p", "--plots", action="store_true", help="write the tokens-vs-time CSV/plot")
    p.add_argument("-c", "--compile", action="store_true", help="accepted for compatibility (CUDA graphs replace torch.compile)")
    p.add_argument("--ckpt", type=Path, default=Path("checkpoint"), help="folder containing the model files")
    p.add_argument("--chunk", type=Path, default=None, help="path of the model chunk")
    p.add_argument("--nodes-config", type=Path, default=SETTINGS_DIR / "configuration.json",
                   help="JSON node topology (default: settings_distr/configuration.json)")
    p.add_argument("--device", type=str, default=None, help="torch device where to load model and tensors")
    p.add_argument("--prompt", type=str, default="Who are you?",
                   help="prompt for all samples, or FILE:<path> with one paragraph per sample")
    p.add_argument("--n-samples", type=int, default=3, help="number of samples to generate")
    p.add_argument("--n-tokens", type=int, default=300, help="tokens to generate per sample")
    p.add_argument("--sequence-length", "--context-length", "--block-size", dest="sequence_length", type=int,
                   default=None, help="truncate the context (smaller KV caches)")
    p.add_argument("--dtype", type=str, default=None, help="float32 | float16 | bfloat16")
    p.add_argument("--time-run", default=None, type=Path, help="CSV file collecting run statistics")
    p.add_argument("--seed", type=int, default=10137, help="random seed")
    # --- extensions ---
    p.add_argument("--engine", default="auto", choices=["auto", "eager", "cuda"], help="stage executor")
    p.add_argument("--greedy", action="store_true", help="arg-max decoding")
    p.add_argument("--temperature", type=float, default=None)
    p.add_argument("--top-k", type=int, default=None)
    p.add_argument("--no-kv-cache", action="store_true",
                   help="GPT-2 generation protocol: no KV caches, the whole context travels the ring every step")
    p.add_argument("--head-on", default="starter", choices=["starter", "finisher"],
                   help="'finisher': first-generation chain, the last node owns ln_f + lm_head and returns logits")
    p.add_argument("--partition", default="auto", choices=["auto", "table", "balanced"], help="layer partition policy")
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    seed_everything(args.seed)
    print("+------------------------+\n| Launching starter node |\n+------------------------+")
    if args.debug:
        setup_debug_log("logs_starter.log")

    from .. import config as C
    from ..parallel.distributed import GPTDistributed
    from ..parallel.scheduler import SamplingParams
    from ..utils.plots import plot_tokens_per_time, write_points_csv

    if args.greedy:
        sampling = SamplingParams.greedy()
    else:
        sampling = SamplingParams(temperature=args.temperature if args.temperature is not None else C.TEMPERATURE,
                                  top_k=args.top_k if args.top_k is not None else C.TOP_K, seed=args.seed)
    gpt_distr = GPTDistributed(
        node_type="starter", config_file=args.nodes_config, ckpt_dir=args.ckpt, chunk_path=args.chunk,
        device=args.device, dtype=args.dtype, model_seq_length=args.sequence_length, verb=args.verb, plots=args.plots,
        compile=args.compile, engine=args.engine, sampling=sampling, partition=args.partition,
        use_kv_cache=not args.no_kv_cache, head_on=args.head_on)
    gen_times = gpt_distr.start(n_samples=args.n_samples, tokens_per_sample=args.n_tokens, prompt=args.prompt)

    if args.plots and gen_times:
        name = tokens_time_csv_name(gpt_distr.n_nodes, gpt_distr.full_model_name, args.n_samples)
        write_points_csv(gen_times, LOGS_DIR / name)
        plot_tokens_per_time(gen_times, out_path=IMG_DIR / name.replace("tokens_time_samples_", "tokens_time_").replace(".csv", ".png"))
    if args.time_run is not None and gen_times:
        cfg = gpt_distr.model_config
        append_run_stats(args.time_run, args.n_samples, cfg.n_layer, cfg.block_size, gen_times[-1][1])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
