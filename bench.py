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
* The run goes THROUGH THE NODE API: rank 0 builds ``GPTDistributed("starter")``, rank i ``GPTDistributed("secondary:i-1")``
  (the objects behind the ``starter`` / ``secondary`` CLIs); HTTP control plane, CUDA-IPC handles exchanged at ``POST /init``,
  fused NVLink hops.  ``open_session().run(rounds)`` is how exactly K rounds get timed (CUDA events on every node, max).
* ``e2e``: the same API with ``decode_mode="host"``: every step copies its descriptor H2D from pinned memory and reads the
  sampled token back D2H; wall clock on the starter.
* after the timed regions: ``tokens_match`` (greedy N-node tokens == one-stage pipeline tokens, and near-arg-max of the eager
  PyTorch model), ``long_run`` (>= 256 timed rounds), and at N=8 BASELINE config #5 (fp8, 2048-token context).
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-table", "reference-nccl"],
                    help="reference = byte-stock reference; reference-table = + injected partition-table entry for node counts it lacks; "
                         "reference-nccl = its data plane swapped for torch.distributed send/recv (baseline/nccl_connections.py)")
    ap.add_argument("--model", default="Llama-3-8B")
    ap.add_argument("--prompt-len", type=int, default=448,
                    help="prompt tokens per sample; the timed rounds then run at a realistic context (~0.5k positions) in both arms")
    ap.add_argument("--seq-len", type=int, default=0, help="KV/context budget (0 = prompt + all rounds)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="rounds of the host-fed e2e measurement (0 = max(steps, 64))")
    ap.add_argument("--n-samples", type=int, default=0, help="concurrent samples (0 = number of GPUs)")
    ap.add_argument("--partition", default="third", choices=["auto", "table", "balanced", "half", "third"],
                    help="table: the reference's N_LAYERS_NODES; balanced: whole layers, head-aware; half: attention|MLP units; third: attention|gate-up|down units")
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
    ap.add_argument("--direct", action="store_true", help="drive DevicePipeline directly under torch.distributed (round-1 harness, A/B only)")
    ap.add_argument("--no-check", action="store_true", help="skip the greedy token check that follows the timed regions")
    ap.add_argument("--long-steps", type=int, default=256, help="rounds of the additional long run reported as long_run (0 = off)")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="N=8: skip the extra fp8 / 2048-context job (BASELINE config #5) that follows the headline job")
    return ap.parse_args()


def run_direct(args: argparse.Namespace) -> Dict[str, Any]:
    """Round-1 harness kept for A/B: ``DevicePipeline`` driven directly under torch.distributed (no node API)."""
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
    e2e_rounds = args.e2e_steps or max(args.steps, 64)
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
        "data": "synthetic prompts, random-init weights", "impl": "ours-direct",
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


def _bench_topology(world: int, job: int) -> Dict[str, Any]:
    """Loopback node JSON (reference schema); ports derived from the rendezvous port so that every rank computes
    the same file without talking and concurrent / stale runs never collide."""
    base = 30000 + (int(os.environ.get("MASTER_PORT", "29500")) % 1000) * 30 + job * 1000

    def node(i: int) -> Dict[str, Any]:
        return {"addr": "127.0.0.1", "communication": {"port": base + i, "starter_addr": "127.0.0.1"},
                "inference": {"port_in": base + 10 + 2 * i, "port_out": base + 11 + 2 * i}, "device": f"cuda:{i}"}

    return {"nodes": {"starter": node(0), "secondary": [node(i) for i in range(1, world)]}}


def run_job(args: argparse.Namespace, job: int = 0, light: bool = False) -> Dict[str, Any]:
    """One benchmark job THROUGH THE NODE API: rank 0 is a ``GPTDistributed("starter")``, rank i a
    ``GPTDistributed("secondary:i-1")`` — the objects behind the ``starter`` / ``secondary`` CLIs — on
    ``cuda:LOCAL_RANK``; HTTP control plane, CUDA-IPC handle exchange at ``/init``, fused NVLink hops.
    torch.distributed is not used by the p2p data plane at all (``--hop nccl`` initialises NCCL for its hops)."""
    import tempfile

    import torch

    from mdi_llm_b200.models.config import Config
    from mdi_llm_b200.parallel.distributed import GPTDistributed
    from mdi_llm_b200.parallel.scheduler import SamplingParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run --nproc-per-node {args.gpus}")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if args.hop == "nccl" and world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    if args.variant >= 0:
        from mdi_llm_b200 import ops as _ops

        _ops.set_linear_variant(args.variant)
    if args.tiny:
        cfg = Config.from_name("tiny-llama-1.1b", n_layer=8, n_embd=512, n_head=8, n_query_groups=2,
                               intermediate_size=1024, vocab_size=2000, padded_vocab_size=2048, block_size=2048)
    else:
        cfg = Config.from_name(args.model)
    n_samples = args.n_samples or world
    e2e_rounds = args.e2e_steps or max(args.steps, 64)
    check_rounds = 16
    rounds_total = args.warmup + args.steps + 1
    # same context budget formula as the reference arm (baseline/run_reference.py): room for the timed rounds only;
    # the e2e and check sessions that follow fit themselves into it
    seq_len = args.seq_len or min(cfg.block_size, ((args.prompt_len + rounds_total + 64) // 64) * 64)
    e2e_rounds = max(1, min(e2e_rounds, seq_len - args.prompt_len - 4))
    ckpt = os.path.join(tempfile.gettempdir(), f"mdi_bench_{os.environ.get('MASTER_PORT', '0')}_{job}_r{rank}", "custom", cfg.name)
    os.makedirs(ckpt, exist_ok=True)
    cfg.save(ckpt)  # model_config.yaml: all the API needs next to synthetic weights
    topo = _bench_topology(world, job)
    sampling = SamplingParams(temperature=args.temperature, top_k=args.top_k, seed=2024)
    common = dict(ckpt_dir=ckpt, device=f"cuda:{local_rank}", dtype="bfloat16", transport=args.hop if world > 1 else "p2p",
                  weights=args.weights, engine="cuda")
    if rank > 0:  # ---- worker nodes: exactly what `secondary.py` does -------------------------------------------
        node = GPTDistributed(f"secondary:{rank - 1}", topo, **common)
        node.start()  # serves POST /init, POST /ring ... until the starter's PUT /stop
        return {}

    import warnings

    warnings.filterwarnings("ignore", message="No tokenizer files")
    gd = GPTDistributed("starter", topo, model_seq_length=seq_len, partition=args.partition, random_init=1234, sampling=sampling,
                        max_prompt_len=args.prompt_len, **common)
    g = torch.Generator().manual_seed(7)
    prompts = [torch.randint(0, cfg.vocab_size, (args.prompt_len,), generator=g, dtype=torch.int32) for _ in range(n_samples)]
    out: Dict[str, Any] = {}
    try:
        # ---------------- device-driven, device-timed (CUDA events on every node, max over nodes) ----------------
        sess = gd.open_session(n_samples, rounds_total, prompts, mode="device")
        plan = [sp["layers"] for sp in gd.specs] if gd.specs else [cfg.n_layer]
        warm = sess.run(args.warmup)  # prefill of every prompt through every stage + W warm-up rounds
        sampler = ClockSampler(world)
        sampler.start()
        timed = sess.run(args.steps)
        clocks = sampler.stop()
        sess.close()
        ms_total = timed["decode_ms"]
        tokens = args.steps * n_samples
        value = tokens / (ms_total / 1e3)
        launches = sum(r["kernel_launches"] for r in timed["per_node"])
        sm_khz = 1_965_000
        waits = [r["wait_cycles"] / (sm_khz / 1e3) / max(1, tokens) for r in sorted(timed["per_node"], key=lambda r: r["rank"])]

        # ---------------- end to end: host-fed steps through the same API objects ----------------------------------
        serv = gd.gpt_serv
        if light or (args.hop == "nccl" and world > 1):
            e2e = {"value": None, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        else:
            pinned = [p.pin_memory() for p in prompts]  # inputs start in pinned host memory
            s2 = serv.open_ring_session(n_samples, pinned, e2e_rounds + 2, mode="host")
            s2.run(1)  # prefill + one untimed round (graph capture of the host-fed variant)
            t0 = time.perf_counter()
            r2 = s2.run(e2e_rounds)
            e2e_s = time.perf_counter() - t0  # wall clock on the starter around the API call; nodes' GPUs are done when it returns
            s2.close()
            steps_e2e = e2e_rounds * n_samples
            st_local = r2["per_node"][0]
            e2e = {"value": round(steps_e2e / e2e_s, 3), "unit": "tokens/s", "h2d_bytes_per_step": st_local["h2d"] // steps_e2e,
                   "d2h_bytes_per_step": st_local["d2h"] // steps_e2e, "rounds": e2e_rounds,
                   "how": "GPTDistributed/GPTServer API, decode_mode=host: pinned step descriptor H2D + sampled token D2H every "
                          "step on the starter; wall clock around the call (includes the HTTP fan-out to the nodes)"}

        # ---------------- correctness, outside the timed regions ---------------------------------------------------
        check = _token_check(serv, cfg, prompts, n_samples, min(check_rounds, seq_len - args.prompt_len - 1), seq_len, world, args) \
            if not (args.no_check or (light and args.weights == "bf16")) else {}

        out = {
            "metric": METRIC.replace("Llama-3-8B", cfg.name) if cfg.name != "Llama-3-8B" else METRIC, "value": round(value, 3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": _vs_published(cfg.name, world, value),
            "dtype": "bf16" if args.weights == "bf16" else "fp8-e4m3 block-scaled weights (128), bf16 activations, fp32 accumulate",
            "data": "synthetic prompts, random-init weights", "impl": "ours",
            # `config`: the benchmark configuration proper — the SAME keys and values in both arms (see baseline/run_reference.py);
            # everything that describes how this arm runs it is under `details`
            "config": {"model": cfg.name, "global_batch": n_samples, "seq_len": seq_len, "prompt_len": args.prompt_len,
                       "parallelism": f"pp{world}", "tokens_per_step": n_samples,
                       "sampling": {"temperature": args.temperature, "top_k": args.top_k},
                       "l2_policy": "inputs (stage weights) larger than L2, no flush"},
            "details": {"n_layer": cfg.n_layer, "plan_layers_per_stage": plan, "partition": args.partition,
                        "api": "GPTDistributed(starter) + GPTDistributed(secondary:i) per GPU, HTTP control plane, "
                               "CUDA-IPC handles exchanged at POST /init",
                        "hop": ("NCCL send/recv (baseline midpoint)" if args.hop == "nccl" else "fused P2P store + flag (NVLink)") if world > 1 else "local (standalone ring)",
                        "timing": "CUDA events around each node's K decode rounds, max over nodes"},
            "clocks": clocks, "e2e": e2e,
            "prefill_ms_all_samples": round(warm["prefill_ms"], 3), "gpu_launches": int(launches),
            "hop_watchdog_status": max(max(r["status"]) for r in timed["per_node"]),
            "stage_wait_us_per_step": [round(x, 2) for x in waits],
            "stage_busy_us_per_step": [round(ms_total * 1e3 / tokens - x, 2) for x in waits],
            "per_node_decode_ms": [round(r["decode_ms"], 3) for r in sorted(timed["per_node"], key=lambda r: r["rank"])],
        }
        out.update(check)
        if args.tiny:
            out["config"]["WARNING"] = "tiny smoke model — not the BASELINE config"
    finally:
        gd.stop_nodes()
        gd.gpt_serv.shutdown()
    return out


def _token_check(serv: Any, cfg: Any, prompts: List[Any], n_samples: int, rounds: int, seq_len: int, world: int,
                 args: argparse.Namespace) -> Dict[str, Any]:
    """Greedy decoding through the N-node API must give the SAME tokens as a one-stage pipeline holding the same
    model (weights are seeded per parameter name, so any partition is the same model), and every token must be the
    (near-)arg-max of the eager PyTorch model's teacher-forced logits."""
    import torch

    from mdi_llm_b200.models.stage import build_stage
    from mdi_llm_b200.parallel.pipeline import DevicePipeline
    from mdi_llm_b200.parallel.scheduler import SamplingParams
    from mdi_llm_b200.utils.checkpoint import random_init_stage_

    dev = serv.torch_model_device
    s = serv.open_ring_session(n_samples, prompts, rounds, mode="device", sampling=SamplingParams.greedy())
    s.run()
    got = s.tokens()
    s.close()
    res: Dict[str, Any] = {"token_check_rounds": rounds}
    full = build_stage(cfg, "starter", cfg.n_layer, meta=True)
    random_init_stage_(full, dev, torch.bfloat16, seed=1234)
    if world > 1 or args.weights != "bf16":
        single = DevicePipeline(full, 0, 1, n_samples=n_samples, max_seq_length=seq_len, sampling=SamplingParams.greedy(),
                                weight_dtype=args.weights)
        ref = single.generate(prompts, rounds)
        res["tokens_match"] = all(torch.equal(got[i], ref[i]) for i in range(n_samples))
        res["tokens_match_what"] = f"{world}-node API run vs one-stage pipeline, greedy, {rounds} tokens x {n_samples} samples"
        del single
    if args.weights == "bf16":  # eager oracle: same weights through stock PyTorch ops
        full.max_seq_length = seq_len
        full.set_kv_cache(1, device=dev, dtype=torch.bfloat16)
        worst = 0.0
        with torch.inference_mode():
            for i in range(min(n_samples, 2)):
                toks = got[i].to(dev)
                T = toks.shape[1]
                h = full(toks[:, :-1].long(), torch.arange(T - 1, device=dev), slot=0)
                logits = full.head(h).float()[0]
                P = prompts[i].numel()
                rows = logits[P - 1:]
                chosen = rows.gather(1, toks[0, P:].long().view(-1, 1)).squeeze(1)
                worst = max(worst, float((rows.max(dim=1).values - chosen).max()))
        res["tokens_vs_eager_max_logit_gap"] = round(worst, 4)
        res["tokens_near_argmax_of_eager"] = worst <= 0.25
        if world == 1:
            res["tokens_match"] = res["tokens_near_argmax_of_eager"]
            res["tokens_match_what"] = "1-node API run vs eager PyTorch model: every token within 0.25 logit of the oracle's arg-max"
    del full
    torch.cuda.empty_cache()
    return res


def run_ours(args: argparse.Namespace) -> Dict[str, Any]:
    """Headline job, then (same ranks, fresh nodes) the BASELINE-length jobs that the 20-step headline cannot show."""
    import copy
    import gc

    out = run_job(args, 0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    extras = []
    if args.extras and not args.tiny and args.long_steps > 0:  # >= 256 timed rounds at the same dtype / prompt length
        a = copy.copy(args)
        a.steps, a.warmup, a.seq_len = args.long_steps, 8, 0
        extras.append(("long_run", a))
    force5 = os.environ.get("MDI_BENCH_FORCE_CFG5", "") not in ("", "0")  # exercise the extra job below at any N (harness test)
    if args.extras and (world == 8 or force5) and not args.tiny and args.weights == "bf16" and args.model == "Llama-3-8B":
        a = copy.copy(args)  # BASELINE config #5: fp8 block-scaled weights, 1024-token prompts, 2048-token context
        a.weights, a.prompt_len, a.seq_len, a.steps, a.warmup = "fp8", 1024, 2048, 256, 8
        extras.append(("baseline_config_5_fp8_2048ctx", a))
    for job, (key, a) in enumerate(extras, start=1):
        import torch

        gc.collect()
        torch.cuda.empty_cache()
        try:
            r = run_job(a, job, light=True)
        except Exception as e:  # noqa: BLE001  (never lose the headline line to an extra)
            r = {"error": repr(e)[:300]}
        if out:
            out[key] = {k: r[k] for k in ("value", "unit", "dtype", "steps", "warmup", "ms_per_step", "prefill_ms_all_samples",
                                          "tokens_match", "tokens_match_what", "config", "details", "error") if k in r}
    return out


def main() -> None:
    args = parse_args()
    if args.impl.startswith("reference"):
        from baseline.run_reference import run_reference

        # the reference prints its generated samples and progress spinners on stdout: keep stdout for
        # the one JSON line by pointing fd 1 at stderr while it runs
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            out = run_reference(args, args.impl)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        # stdout carries exactly ONE line (the JSON): anything the framework prints while it runs goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            out = run_direct(args) if args.direct else run_ours(args)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if out:
        print(json.dumps(out), flush=True)
    if args.impl == "reference-nccl":  # RX threads may sit in a posted NCCL recv: leave without tearing the communicators down
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
