"""Recurrent-pipeline scheduler (host-driven form).

The reference keeps ``>= n_nodes`` independent samples in flight so that every stage is always
busy with a different sample; each sample's hidden state travels the ring
starter -> sec0 -> ... -> starter and only the newest token is forwarded once the per-sample KV
caches are warm (``README.md:193-246``; loops: ``gptserver.py:788-1019`` starter,
``:1021-1110`` secondary).  This module reproduces that protocol on top of two abstractions:

* a :class:`StageRunner` — "run my slice of the model for sample *i*" (eager PyTorch here,
  the CUDA stage executor in :mod:`mdi_llm_b200.parallel.engine`);
* a :class:`~mdi_llm_b200.parallel.transport.base.Transport`.

Protocol details preserved: FIFO processing; per sample ``iter 0`` is the prefill of the whole
prompt, later iterations carry one token; exactly ``max_new_tokens`` tokens are produced per
sample; a finished sample emits a ``stop`` marker that travels the ring and the starter ends
when the first marker returns (gptserver.py:919-921, 985-994); secondaries create per-sample
state lazily on first sight (:1083-1088).

Two earlier generations of the protocol are kept as options:

* ``use_kv_cache=False`` — the GPT-2 generation: no KV caches, every message carries the whole
  growing context ``(1,T,C)`` cropped to ``block_size`` (old/GPT2/sub/model_dist.py:959-972);
* ``head_remote=True`` — the nanoGPT generation's starter / intermediate / finisher chain: the last
  node owns ``ln_f`` + ``lm_head`` and returns logits to the starter, which only samples
  (old/nanoGPT/sub/model_dist.py:90-221).

The device-driven form of the same schedule (static round-robin order, no host in the loop,
fused P2P hops) lives in ``engine.py``; both produce identical tokens under greedy sampling.
"""
from __future__ import annotations

import threading
import time
import warnings
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ..models.gpt import sample as sample_logits
from ..models.stage import StageModule, StarterNode
from .transport.base import Message, Transport, build_msg

__all__ = ["SamplingParams", "StageRunner", "EagerStageRunner", "starter_loop", "secondary_loop",
           "GenerationResult"]


@dataclass
class SamplingParams:
    temperature: float = 0.8
    top_k: Optional[int] = 200
    top_p: float = 1.0
    seed: Optional[int] = None

    @classmethod
    def greedy(cls) -> "SamplingParams":
        return cls(temperature=0.0, top_k=None, top_p=0.0)


class StageRunner:
    """What the scheduler needs from a stage."""

    role: str = "stage"
    device: torch.device

    def begin_sample(self, sample_id: int) -> None:  # allocate per-sample state (KV slot)
        raise NotImplementedError

    def forward(self, sample_id: int, data: torch.Tensor, input_pos: Optional[torch.Tensor]) -> torch.Tensor:
        """starter: ``data`` = token ids ``[1,T]``; secondary: hidden state ``[1,T,C]``.
        ``input_pos=None`` = cache-less causal forward over the whole of ``data``."""
        raise NotImplementedError

    def head(self, hidden: torch.Tensor) -> torch.Tensor:
        """starter only: ``ln_f`` + ``lm_head`` on the last position -> logits ``[1,1,V]``."""
        raise NotImplementedError

    def to_device(self, data: Any) -> Any:
        return data.to(self.device) if isinstance(data, torch.Tensor) else data


class EagerStageRunner(StageRunner):
    """Stage runner on plain PyTorch modules (CPU or GPU) with a slot KV pool."""

    def __init__(self, model: StageModule, n_slots_hint: int = 1) -> None:
        self.model = model.eval()
        self.role = model.role
        p = next(model.parameters())
        self.device, self.dtype = p.device, p.dtype
        self.slots: Dict[int, int] = {}
        self._hint = max(1, n_slots_hint)

    def begin_sample(self, sample_id: int) -> None:
        if sample_id in self.slots:
            return
        self.slots[sample_id] = len(self.slots)
        self.model.ensure_slots(max(self._hint, len(self.slots)))

    @torch.inference_mode()
    def forward(self, sample_id: int, data: torch.Tensor, input_pos: Optional[torch.Tensor]) -> torch.Tensor:
        slot = self.slots.get(sample_id, 0) if input_pos is None else self.slots[sample_id]
        if self.role == "starter":
            return self.model(data.long(), input_pos, slot=slot)
        return self.model(data.to(self.dtype), input_pos, slot=slot)

    @torch.inference_mode()
    def head(self, hidden: torch.Tensor) -> torch.Tensor:
        assert isinstance(self.model, StarterNode)
        return self.model.head(hidden[:, -1:].to(self.dtype))


@dataclass
class GenerationResult:
    samples: Dict[int, torch.Tensor]  # sample id -> (1, prompt + generated) token ids (cpu)
    prompt_lengths: Dict[int, int]
    tok_time: List[Tuple[int, float]] = field(default_factory=list)
    n_tokens: int = 0
    elapsed: float = 0.0


def starter_loop(
    runner: StageRunner,
    transport: Transport,
    prompts: Sequence[torch.Tensor],
    max_new_tokens: int,
    sampling: SamplingParams,
    running: threading.Event,
    n_nodes: int = 1,
    record_times: bool = True,
    on_token: Optional[Callable[[int, int], None]] = None,
    recv_timeout: float = 2.0,
    watchdog_s: Optional[float] = None,
    use_kv_cache: bool = True,
    head_remote: bool = False,
    block_size: Optional[int] = None,
) -> GenerationResult:
    """Generation loop of node 0.  ``prompts[i]`` is a 1-D tensor of token ids."""
    n_samples = len(prompts)
    if n_samples < 1:
        raise ValueError("Cannot generate less than 1 sample!")
    if n_samples < n_nodes:
        warnings.warn(f"Generating less samples ({n_samples}) than nodes ({n_nodes}) will not be efficient!")
    gen = None
    if sampling.seed is not None:
        gen = torch.Generator(device="cpu").manual_seed(sampling.seed)

    samples: Dict[int, torch.Tensor] = {}
    prompt_len = {i: int(p.numel()) for i, p in enumerate(prompts)}
    iter_ind = {i: 0 for i in range(n_samples)}
    input_pos: Dict[int, torch.Tensor] = {}
    # the reference seeds its own input queue with one prefill message per sample
    pending: List[Message] = [build_msg(p.view(1, -1), i) for i, p in enumerate(prompts)]
    tok_time: List[Tuple[int, float]] = [(0, 0.0)] if record_times else []
    n_tokens = 0
    t0 = time.time()
    last_progress = time.time()

    while running.is_set():
        if pending:
            msg: Optional[Message] = pending.pop(0)
        else:
            msg = transport.recv(timeout=recv_timeout)
        if msg is None:
            if watchdog_s is not None and time.time() - last_progress > watchdog_s:
                running.clear()
                raise TimeoutError(f"no message from the ring for {watchdog_s} s (dead node?)")
            continue
        last_progress = time.time()
        if msg.get("stop"):
            break  # first stop marker completed the ring: every sample is done (FIFO order)
        sid = msg["sample_index"]
        data = runner.to_device(msg["data"])
        if iter_ind[sid] >= 1:
            logits = data if head_remote else runner.head(data)
            nxt = sample_logits(logits.float().cpu() if gen is not None else logits.float(),
                                temperature=sampling.temperature, top_k=sampling.top_k,
                                top_p=sampling.top_p, generator=gen)
            nxt = nxt.view(1, 1).to(samples[sid].device, samples[sid].dtype)
            samples[sid] = torch.cat((samples[sid], nxt), dim=1)
            input_pos[sid] = input_pos[sid][-1:] + 1
            n_tokens += 1
            if record_times:
                tok_time.append((n_tokens, time.time() - t0))
            if on_token is not None:
                on_token(sid, int(nxt))
        else:
            samples[sid] = data.view(1, -1)
            if use_kv_cache:
                runner.begin_sample(sid)
            input_pos[sid] = torch.arange(0, prompt_len[sid], device=runner.device)

        if iter_ind[sid] < max_new_tokens:
            if use_kv_cache:
                idx_cond = samples[sid] if iter_ind[sid] == 0 else samples[sid][:, -1:]
                out = build_msg(runner.forward(sid, idx_cond, input_pos[sid]), sid)
            else:  # whole context again, cropped to the model's window
                idx_cond = samples[sid] if block_size is None else samples[sid][:, -block_size:]
                out = build_msg(runner.forward(sid, idx_cond, None), sid)
        else:
            out = build_msg("", sid, stop=True)
        iter_ind[sid] += 1
        transport.send(out)

    return GenerationResult(
        samples={i: s.detach().cpu() for i, s in samples.items()},
        prompt_lengths=prompt_len, tok_time=tok_time, n_tokens=n_tokens, elapsed=time.time() - t0,
    )


def secondary_loop(
    runner: StageRunner,
    transport: Transport,
    running: threading.Event,
    n_samples: Optional[int] = None,
    recv_timeout: float = 2.0,
    use_kv_cache: bool = True,
) -> int:
    """Worker loop of nodes 1..N-1: forward every hidden state, relay stop markers.  Returns the
    number of forwards executed.  Ends when ``running`` is cleared (PUT /stop)."""
    input_pos: Dict[int, torch.Tensor] = {}
    n_fwd = 0
    while running.is_set():
        msg = transport.recv(timeout=recv_timeout)
        if msg is None:
            continue
        sid = msg["sample_index"]
        if msg.get("stop"):
            transport.send(msg)
            continue
        data = runner.to_device(msg["data"])
        if not use_kv_cache:
            transport.send(build_msg(runner.forward(sid, data, None), sid))
            n_fwd += 1
            continue
        if sid not in input_pos:
            if n_samples is not None and n_fwd >= n_samples:
                raise AssertionError("Should have seen this sample already...")
            runner.begin_sample(sid)
            input_pos[sid] = torch.arange(0, data.size(1), device=runner.device)
        out = runner.forward(sid, data, input_pos[sid])
        transport.send(build_msg(out, sid))
        input_pos[sid] = input_pos[sid][-1:] + 1
        n_fwd += 1
    return n_fwd
