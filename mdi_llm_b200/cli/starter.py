#!/usr/bin/env python3
"""Starter node (node 0) of a model-distributed inference run.

Parity: reference ``src/starter.py`` — same flags (:109-196): ``-d -v -p -c --ckpt --chunk
--nodes-config --device --prompt --n-samples --n-tokens --sequence-length|--context-length|
--block-size --dtype --time-run --seed``; writes the tokens/time CSV (+PNG when matplotlib is
available) with ``-p`` and appends run statistics with ``--time-run``.  Extra flags select the
B200 execution path: ``--engine`` (auto|eager|cuda), ``--greedy``, ``--temperature``, ``--top-k``, ``--top-p``,
``--partition`` (incl. ``half``), ``--transport/--hop`` (auto|socket|p2p|nccl), ``--weights`` (bf16|fp8),
``--decode-mode`` (device|host), ``--random-init SEED``.  With ``--transport auto`` and every node on a GPU of this
box the run uses the device ring — the benchmarked path — through exactly this CLI.
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
    p.add_argument("--push-chunks", action="store_true",
                   help="after splitting the model on the fly, send every secondary its chunk inside POST /init (the reference's "
                        "behaviour) instead of expecting it to read the chunk file; a secondary that answers that it has no "
                        "parameters gets its chunk this way in any case")
    p.add_argument("--partition", default="auto", choices=["auto", "table", "balanced", "half", "third"],
                   help="layer partition policy (half / third: boundaries may fall between a layer's attention, gate/up and down passes)")
    p.add_argument("--top-p", type=float, default=None, help="nucleus sampling threshold")
    p.add_argument("--transport", "--hop", dest="transport", default="auto", choices=["auto", "socket", "p2p", "nccl"],
                   help="inter-node data plane: p2p = fused NVLink stores + flags (device ring, one box), nccl = same "
                        "kernels with NCCL send/recv hops, socket = the reference's TCP + pickle; auto picks p2p when possible")
    p.add_argument("--weights", default="bf16", choices=["bf16", "fp8"], help="serve bf16 or block-scaled fp8 weights (device ring)")
    p.add_argument("--decode-mode", default="device", choices=["device", "host"],
                   help="device ring: device-driven steps (fastest) or host-fed steps with every token read back as it appears")
    p.add_argument("--random-init", type=int, default=None, metavar="SEED",
                   help="synthetic weights instead of checkpoint chunks (benchmarks; needs only model_config.yaml in --ckpt)")
    p.add_argument("--max-prompt-len", type=int, default=None, help="size of the prefill hop buffers (default: context length)")
    p.add_argument("--watchdog", type=float, default=None, help="seconds a stage waits for its neighbour before aborting the ring")
    p.add_argument("--tokens-out", type=Path, default=None,
                   help="write the generated token ids, the data plane used and the tokens/time points as JSON")
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
                                  top_k=args.top_k if args.top_k is not None else C.TOP_K,
                                  top_p=args.top_p if args.top_p is not None else 1.0, seed=args.seed)
    gpt_distr = GPTDistributed(
        node_type="starter", config_file=args.nodes_config, ckpt_dir=args.ckpt, chunk_path=args.chunk,
        device=args.device, dtype=args.dtype, model_seq_length=args.sequence_length, verb=args.verb, plots=args.plots,
        compile=args.compile, engine=args.engine, sampling=sampling, partition=args.partition, push_chunks=args.push_chunks,
        use_kv_cache=not args.no_kv_cache, head_on=args.head_on, transport=args.transport, weights=args.weights,
        decode_mode=args.decode_mode, random_init=args.random_init, max_prompt_len=args.max_prompt_len,
        watchdog_s=args.watchdog)
    gen_times = gpt_distr.start(n_samples=args.n_samples, tokens_per_sample=args.n_tokens, prompt=args.prompt)

    if args.tokens_out is not None:
        import json

        res = gpt_distr.gpt_serv.last_result
        args.tokens_out.parent.mkdir(parents=True, exist_ok=True)
        args.tokens_out.write_text(json.dumps({
            "transport": getattr(gpt_distr, "transport", None), "n_nodes": gpt_distr.n_nodes,
            "tokens": {str(i): t.view(-1).tolist() for i, t in sorted(res.samples.items())} if res else {},
            "prompt_lengths": {str(i): n for i, n in res.prompt_lengths.items()} if res else {},
            "tok_time": gen_times or []}))
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
