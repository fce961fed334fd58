#!/usr/bin/env python3
"""Headline benchmark: generated tokens/sec of Llama-3-8B recurrent-pipeline decode on N B200s.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (N > 1: launched under
``torch.distributed.run`` with one rank per GPU).  Prints ONE JSON line on rank 0.

* One *step* = one decode round of the recurrent pipeline: every one of the ``n_samples = N``
  concurrent samples advances by one token (BASELINE.json configs: N GPUs ↔ N samples), so a step
  generates N tokens and ``value = K * N / t`` is the whole-box aggregate.  Weak scaling: the
  number of in-flight samples grows with N while each GPU holds 1/N of the layers.
* Timed region = exactly K rounds after W warm-up rounds (prefill excluded: the metric is
  *decode*), bracketed by barrier + ``torch.cuda.synchronize()``; device time from CUDA events,
  MAX over ranks.  Inputs larger than L2: each round streams the stage's weights
  (16 GB / N ≫ 126 MB L2) so no L2 flush is needed between iterations.
* ``e2e``: the same rounds through the public host-fed API (``DevicePipeline`` mode="host"):
  every step copies its descriptor H2D from pinned memory and reads the sampled token back D2H.
* ``--impl reference`` runs the unmodified reference (baseline/_ref) — see baseline/run_reference.py.

Synthetic prompts, random-init weights of the Llama-3-8B architecture (no network on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from typing import Any, Dict, List, Optional

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "generated tokens/sec (whole box, device-timed, max over ranks) Llama-3-8B recurrent-pipeline decode"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (the recipe's clocks line).
    NVML in-process (a sample every 2 ms, so even a 100 ms region gets dozens); falls back to one
    ``nvidia-smi`` poller when NVML is not importable."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, n_gpus: int) -> None:
        self.n = n_gpus
        self.proc: Optional[subprocess.Popen] = None
        self.lines: List[str] = []
        self.sm: List[float] = []
        self.mx: List[float] = []
        self.reasons: set = set()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._nvml = None

    def start(self) -> None:
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(min(self.n, pynvml.nvmlDeviceGetCount()))]
            self._thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self._thread.start()
            return
        except Exception:  # noqa: BLE001
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll_nvml(self) -> None:
        nv = self._nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            for h in self._handles:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def _pump(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self) -> Dict[str, Any]:
        if self._nvml is not None:
            self._stop.set()
            if self._thread is not None:
                self._thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                if int(f[0]) >= self.n:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# the only numbers the reference publishes (Jetson TX2 testbed, read off its plots); nothing for
# Llama-3-8B, so the headline config reports null
PUBLISHED_TOK_S = {("tiny-llama-1.1b", 2): 13.3, ("tiny-llama-1.1b", 3): 17.9, ("NanoLlama", 1): 17.9, ("NanoLlama", 2): 25.8,
                   ("NanoLlama", 3): 30.4}


def _vs_published(model: str, n_nodes: int, value: float) -> Optional[float]:
    ref = PUBLISHED_TOK_S.get((model, n_nodes))
    return round(value / ref, 2) if ref else None


def parse_args() -> argparse.Namespace:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="Llama-3-8B")
    ap.add_argument("--prompt-len", type=int, default=64)
    ap.add_argument("--seq-len", type=int, default=0, help="KV/context budget (0 = prompt + all rounds)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="rounds of the host-fed e2e measurement (0 = min(steps, 64))")
    ap.add_argument("--n-samples", type=int, default=0, help="concurrent samples (0 = number of GPUs)")
    ap.add_argument("--partition", default="half", choices=["auto", "table", "balanced", "half"],
                    help="table: the reference's N_LAYERS_NODES; balanced: whole layers, head-aware; half: attention|MLP half-layer units")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--ctas-per-sm", type=int, default=4)
    ap.add_argument("--weights", default="bf16", choices=["bf16", "fp8"],
                    help="bf16 (headline) or fp8-e4m3 block-scaled weights (BASELINE config #5; NOT the headline dtype)")
    ap.add_argument("--hop", default="p2p", choices=["p2p", "nccl"],
                    help="inter-stage hop: fused peer stores + flags (product) or NCCL send/recv (baseline midpoint)")
    ap.add_argument("--variant", type=int, default=-1, help="decode linear path: 0 LDG, 1 bulk-copy x4 stages, 2 bulk-copy x2")
    ap.add_argument("--temperature", type=float, default=0.8)
    ap.add_argument("--top-k", type=int, default=200)
    ap.add_argument("--tiny", action="store_true", help="tiny model (CI smoke of the harness; NOT a valid bench number)")
    return ap.parse_args()


def run_ours(args: argparse.Namespace) -> Dict[str, Any]:
    import torch
    import torch.distributed as dist

    from mdi_llm_b200.models.config import Config
    from mdi_llm_b200.models.partition import plan_layers
    from mdi_llm_b200.models.stage import build_stage
    from mdi_llm_b200.parallel.pipeline import DevicePipeline
    from mdi_llm_b200.parallel.scheduler import SamplingParams
    from mdi_llm_b200.utils.checkpoint import random_init_stage_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run --nproc-per-node {args.gpus}")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier() -> None:
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    if args.variant >= 0:
        from mdi_llm_b200 import ops as _ops

        _ops.set_linear_variant(args.variant)
    if args.tiny:
        cfg = Config.from_name("tiny-llama-1.1b", n_layer=8, n_embd=512, n_head=8, n_query_groups=2,
                               intermediate_size=1024, vocab_size=2000, padded_vocab_size=2048, block_size=2048)
    else:
        cfg = Config.from_name(args.model)
    n_samples = args.n_samples or world
    e2e_rounds = args.e2e_steps or min(args.steps, 64)
    rounds_total = args.warmup + args.steps + 1
    seq_len = args.seq_len or min(cfg.block_size, ((args.prompt_len + max(rounds_total, e2e_rounds + 4) + 64) // 64) * 64)
    role = "starter" if rank == 0 else f"secondary:{rank - 1}"
    if args.partition == "half" and world > 1 and not cfg.parallel_residual:
        # half-layer units (attention | MLP): pipeline boundaries may fall inside a layer
        from mdi_llm_b200.models.partition import half_stages, plan_half_units

        units = plan_half_units(world, cfg)
        hs = half_stages(units)[rank]
        plan = [u / 2 for u in units]
        stage = build_stage(cfg, role, hs.n_blocks, meta=True, first_mlp_only=hs.first_mlp_only,
                            last_attn_only=hs.last_attn_only)
    else:
        policy = "balanced" if args.partition == "half" else args.partition
        plan = plan_layers(world, cfg.n_layer, cfg, policy=policy) if world > 1 else [cfg.n_layer]
        stage = build_stage(cfg, role, plan[rank], meta=True)
    random_init_stage_(stage, device, torch.bfloat16, seed=1234 + rank)
    sampling = SamplingParams(temperature=args.temperature, top_k=args.top_k, seed=2024)
    pipe = DevicePipeline(stage, rank, world, n_samples=n_samples, max_seq_length=seq_len, sampling=sampling,
                          max_prompt_len=args.prompt_len, use_pdl=not args.no_pdl, ctas_per_sm=args.ctas_per_sm, hop=args.hop,
                          weight_dtype=args.weights, free_bf16=args.weights == "fp8")
    pipe.connect_distributed()
    g = torch.Generator().manual_seed(7)
    prompts = [torch.randint(0, cfg.vocab_size, (args.prompt_len,), generator=g, dtype=torch.int32) for _ in range(n_samples)]

    # ---------------- device-driven, device-timed ----------------
    pipe.prepare(prompts, rounds_total)
    barrier()
    pf0, pf1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pf0.record()
    pipe.prefill()  # round 0: every prompt through every stage (tcgen05 GEMMs + attention, fused hop)
    pf1.record()
    barrier()
    prefill_ms = torch.tensor([pf0.elapsed_time(pf1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(prefill_ms, op=dist.ReduceOp.MAX)
    pipe.decode_rounds(args.warmup)
    barrier()
    sampler = ClockSampler(world)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = pipe.n_graph_launches
    wait0 = pipe.stage.wait_cycles()
    ev0.record()
    pipe.decode_rounds(args.steps)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    graph_nodes = next(iter(pipe.stage._graphs.values())).n_nodes if pipe.stage._graphs else 0
    launches = torch.tensor([(pipe.n_graph_launches - launches0) * graph_nodes], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    tokens = args.steps * n_samples
    value = tokens / (ms_total / 1e3)
    status = int(pipe.stage.status[0].item())
    # exposed wait per stage step (cycles CTA 0 spun on the incoming hop flag), gathered from all ranks
    sm_clock_khz = torch.cuda.get_device_properties(device).clock_rate if hasattr(torch.cuda.get_device_properties(device), "clock_rate") else 1_965_000
    wait_us = (pipe.stage.wait_cycles() - wait0) / (sm_clock_khz / 1e3) / max(1, args.steps * n_samples)
    waits = torch.zeros(world, device=device, dtype=torch.float64)
    waits[rank] = wait_us
    if world > 1:
        dist.all_reduce(waits, op=dist.ReduceOp.SUM)

    # ---------------- end to end through the host-fed public API ----------------
    if args.hop == "nccl":  # comparison midpoint only: no separate e2e measurement
        e2e_value, h2d, d2h, steps_e2e = None, 0, 0, 1
    else:
        pinned = [p.pin_memory() for p in prompts]  # inputs start in pinned host memory
        pipe.prepare(pinned, e2e_rounds + 2)
        barrier()
        pipe.prefill()
        pipe.decode_rounds_host(1)  # one untimed round (graph capture of the host-fed variant)
        barrier()
        t0 = time.perf_counter()
        _, h2d, d2h = pipe.decode_rounds_host(e2e_rounds)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = e2e_rounds * n_samples / float(e2e_s.item())
        steps_e2e = e2e_rounds * n_samples

    out = {
        "metric": METRIC.replace("Llama-3-8B", cfg.name) if cfg.name != "Llama-3-8B" else METRIC, "value": round(value, 3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": _vs_published(cfg.name, world, value),
        "dtype": "bf16" if args.weights == "bf16" else "fp8-e4m3 block-scaled weights (128), bf16 activations, fp32 accumulate",
        "data": "synthetic prompts, random-init weights", "impl": "ours",
        "config": {"model": cfg.name, "n_layer": cfg.n_layer, "global_batch": n_samples, "seq_len": seq_len,
                   "prompt_len": args.prompt_len, "parallelism": f"pp{world} recurrent pipeline, plan {plan}",
                   "tokens_per_step": n_samples, "l2_policy": "inputs (stage weights) larger than L2, no flush",
                   "sampling": {"temperature": args.temperature, "top_k": args.top_k}, "pdl": not args.no_pdl,
                   "linear_variant": args.variant, "ctas_per_sm": args.ctas_per_sm,
                   "hop": ("NCCL send/recv (baseline midpoint)" if args.hop == "nccl" else "fused P2P store + flag (NVLink)") if world > 1 else "local (standalone ring)",
                   "timing": "CUDA events, max over ranks"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3) if e2e_value is not None else None, "unit": "tokens/s", "h2d_bytes_per_step": h2d // max(1, steps_e2e),
                "d2h_bytes_per_step": d2h // max(1, steps_e2e), "rounds": e2e_rounds,
                "how": "host-fed steps: pinned ctx H2D + sampled-token D2H every step, wall clock, max over ranks"},
        "prefill_ms_all_samples": round(float(prefill_ms.item()), 3), "gpu_launches": int(launches.item()),
        "hop_watchdog_status": status,
        "stage_wait_us_per_step": [round(x, 2) for x in waits.tolist()],
        "stage_busy_us_per_step": [round(ms_total * 1e3 / (args.steps * n_samples) - x, 2) for x in waits.tolist()],
    }
    if args.tiny:
        out["config"]["WARNING"] = "tiny smoke model — not the BASELINE config"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else {}


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        from baseline.run_reference import run_reference

        # the reference prints its generated samples and progress spinners on stdout: keep stdout for
        # the one JSON line by pointing fd 1 at stderr while it runs
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            out = run_reference(args)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        out = run_ours(args)
    if out:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
