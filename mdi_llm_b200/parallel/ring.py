"""The device ring behind the node API: ``GPTServer`` / ``GPTDistributed`` / ``starter`` / ``secondary``
driving :class:`~.pipeline.DevicePipeline` (fused NVLink hops) instead of sockets.

In the reference the API *is* the data path: ``GPTDistributed.start()`` → ``configure_nodes``
(``src/sub/model_dist.py:402-484``) → every node opens its sockets at init
(``src/sub/gptserver.py:540-583,1123-1203``) and the generation loops push pickles through them.
Here the same calls set up peer-mapped buffers instead:

1. ``POST /init`` — a secondary builds its stage *and* its :class:`RingBackend` (hop buffers are
   ``cudaMalloc``'ed so that they can be exported) and answers with the 64-byte CUDA-IPC handles of
   ``hidden_in``+``flags`` and of the prefill buffer;
2. ``POST /ring {"op": "connect"}`` — the starter tells every node the handles of its successor;
   the node maps them (``cudaIpcOpenMemHandle``) — from then on stage *i*'s last kernel stores
   straight into stage *i+1*'s memory and releases its flag, no host, socket or NCCL call on the path;
3. ``POST /ring {"op": "prepare"}`` / ``{"op": "run"}`` — per generation: reset the per-sample state,
   then enqueue the prefill and the decode rounds (CUDA-graph replays; the device derives every
   step descriptor itself).  A ``run`` may cover only some rounds, which is how ``bench.py`` times
   exactly K rounds through this API, and returns the node's device-side time of the segment.

Secondaries always run device-driven.  The starter runs device-driven too (``decode_mode="device"``,
tokens and their device timestamps are read back once at the end) or host-fed
(``decode_mode="host"``: every step's descriptor goes H2D from pinned memory and every sampled token
comes back D2H before the next step — the streaming mode used by ``chat`` and by the end-to-end
measurement).
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch

from .control import call_node
from .scheduler import SamplingParams

__all__ = ["RingBackend", "RingSession", "RingError", "ring_capable"]


class RingError(RuntimeError):
    pass


def ring_capable(model_device: str, dtype: torch.dtype, config: Any) -> bool:
    """Can this node run the fused device ring?  (CUDA device, kernels present, architecture covered)"""
    if not str(model_device).startswith("cuda") or not torch.cuda.is_available():
        return False
    from .engine import engine_supports

    return engine_supports(config, dtype)


def _spin_until(t: Optional[float]) -> None:
    """Common start time for a segment on every node of the box (same wall clock): the HTTP fan-out
    reaches the nodes a few hundred microseconds apart, the segment starts together."""
    if t is None:
        return
    while time.time() < t:
        pass


class RingBackend:
    """One node's end of the ring: its :class:`DevicePipeline` plus the per-generation bookkeeping."""

    def __init__(self, model: Any, rank: int, world: int, n_samples: int, max_seq_length: int,
                 sampling: Optional[SamplingParams] = None, max_prompt_len: int = 0, hop: str = "p2p",
                 weight_dtype: str = "bf16", wait_max_cycles: int = 20_000_000_000, **pipe_kw: Any) -> None:
        from .pipeline import DevicePipeline

        self.pipe = DevicePipeline(model, rank, world, n_samples=n_samples, max_seq_length=max_seq_length,
                                   sampling=sampling, max_prompt_len=max_prompt_len or max_seq_length, hop=hop,
                                   weight_dtype=weight_dtype, free_bf16=weight_dtype == "fp8",
                                   wait_max_cycles=wait_max_cycles, **pipe_kw)
        self.rank, self.world, self.n_samples = rank, world, n_samples
        self.device = self.pipe.device
        self.lock = threading.Lock()
        self.prepared = False

    # ---- wiring ---------------------------------------------------------------------------------------
    def handles(self) -> Dict[str, Any]:
        return self.pipe.export_handles()

    def connect(self, nxt: Optional[Dict[str, Any]]) -> None:
        if self.world == 1 or nxt is None:
            return
        if self.pipe.hop == "nccl":
            self.pipe.connect_distributed()
        else:
            self.pipe.connect_ipc(nxt)

    # ---- one generation -------------------------------------------------------------------------------
    def prepare(self, prompt_lens: Sequence[int], max_new_tokens: int,
                prompts: Optional[Sequence[torch.Tensor]] = None) -> None:
        if prompts is None:  # secondaries only need the lengths
            prompts = [torch.zeros(int(n), dtype=torch.int32) for n in prompt_lens]
        with self.lock:
            self.pipe.prepare(prompts, max_new_tokens)
            torch.cuda.synchronize(self.device)
            self.prepared = True

    def run(self, prefill: bool, rounds: int, start_at: Optional[float] = None, mode: str = "device",
            on_token: Optional[Callable[[int, int, int], None]] = None) -> Dict[str, Any]:
        """Enqueue (prefill and) ``rounds`` decode rounds, wait for the GPU, report the segment."""
        if not self.prepared:
            raise RingError("run before prepare")
        p = self.pipe
        with self.lock, torch.cuda.device(self.device):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            launches0, kern0, wait0 = p.n_graph_launches, p.n_kernel_launches, p.stage.wait_cycles()
            h2d = d2h = 0
            _spin_until(start_at)
            ev[0].record()
            if prefill:
                p.prefill()
            ev[1].record()
            if rounds > 0:
                if mode == "host" and p.is_starter:
                    _, h2d, d2h = p.decode_rounds_host(rounds, on_token=on_token)
                else:
                    p.decode_rounds(rounds)
            ev[2].record()
            torch.cuda.synchronize(self.device)
            status = p.stage.status[:2].tolist()
            return {"rank": self.rank, "prefill_ms": ev[0].elapsed_time(ev[1]), "decode_ms": ev[1].elapsed_time(ev[2]),
                    "steps": p.n_graph_launches - launches0, "kernel_launches": p.n_kernel_launches - kern0,
                    "wait_cycles": p.stage.wait_cycles() - wait0, "status": status, "h2d": h2d, "d2h": d2h,
                    "round": p.round}

    def abort(self) -> None:
        """Host-initiated abort: poison this node's incoming flags — its queued steps stop waiting, mark
        themselves aborted and pass the poison on, so the whole ring drains (see ``hop_wait``)."""
        self.pipe.poison()

    def tokens(self) -> Dict[int, torch.Tensor]:
        return self.pipe.result_tokens()

    def token_times(self) -> List[float]:
        return self.pipe.token_times()

    def close(self) -> None:
        self.pipe.close()

    # ---- control-plane entry point (secondaries) ----------------------------------------------------------
    def handle(self, msg: Dict[str, Any]) -> Dict[str, Any]:
        op = msg.get("op")
        if op == "connect":
            self.connect(msg.get("next"))
            return {"ok": True}
        if op == "prepare":
            self.prepare(msg["prompt_lens"], int(msg["max_new_tokens"]))
            return {"ok": True}
        if op == "run":
            return self.run(bool(msg.get("prefill", False)), int(msg.get("rounds", 0)), msg.get("start_at"))
        if op == "abort":
            self.abort()
            return {"ok": True}
        if op == "stats":
            return {"status": self.pipe.stage.status[:2].tolist(), "round": self.pipe.round}
        raise RingError(f"unknown ring op {op!r}")


class RingSession:
    """Starter-side driver of one generation over the ring (local backend + secondaries over HTTP)."""

    def __init__(self, backend: RingBackend, secondaries: Sequence[Dict[str, Any]], prompts: Sequence[torch.Tensor],
                 max_new_tokens: int, mode: str = "device", timeout: float = 3600.0) -> None:
        if mode not in ("device", "host"):
            raise ValueError("decode mode must be 'device' or 'host'")
        self.backend, self.mode, self.timeout = backend, mode, timeout
        self.urls = [f"http://{s['addr']}:{s['communication']['port']}/ring" for s in secondaries]
        self.pool = ThreadPoolExecutor(max_workers=max(1, len(self.urls))) if self.urls else None
        self.prompt_lens = [int(p.numel()) for p in prompts]
        self.max_new = int(max_new_tokens)
        self.n_samples = len(prompts)
        self.rounds_done = 0
        self.prefilled = False
        self.t0_host = 0.0
        # 1) reset everywhere (flags must be zero before any producer may signal), 2) then run
        futs = self._fan({"op": "prepare", "prompt_lens": self.prompt_lens, "max_new_tokens": self.max_new})
        pinned = [p.to(torch.int32).pin_memory() if not p.is_cuda and torch.cuda.is_available() else p for p in prompts]
        self.backend.prepare(self.prompt_lens, self.max_new, prompts=pinned)  # H2D of the prompts from pinned memory
        self._join(futs)

    def _post(self, url: str, msg: Dict[str, Any], timeout: Optional[float] = None) -> Dict[str, Any]:
        status, body = call_node("post", url, msg, max_n_requests=1,  # ring ops are not idempotent: never re-post
                                 timeout=self.timeout if timeout is None else timeout)
        if status != 200 or not isinstance(body, dict):
            raise RingError(f"node {url} answered {status}: {body}")
        return body

    def _fan(self, msg: Dict[str, Any]) -> List[Any]:
        if not self.pool:
            return []
        return [self.pool.submit(self._post, u, msg) for u in self.urls]

    @staticmethod
    def _join(futs: Sequence[Any]) -> List[Dict[str, Any]]:
        return [f.result() for f in futs]

    def run(self, rounds: Optional[int] = None, on_token: Optional[Callable[[int, int, int], None]] = None) -> Dict[str, Any]:
        """Prefill (first call) + ``rounds`` decode rounds (default: all that remain) on every node.
        Returns per-node device times of the segment (``decode_ms`` = CUDA events around the node's
        decode launches) and their maximum."""
        left = self.max_new - self.rounds_done
        rounds = left if rounds is None else min(int(rounds), left)
        prefill = not self.prefilled
        # device-driven segments start together on every node (their CUDA-event times are then comparable); a
        # host-fed segment is paced by the starter's own token read-backs, so nobody waits for a common start
        start_at = time.time() + (0.004 if self.urls and self.mode == "device" else 0.0)
        futs = self._fan({"op": "run", "prefill": prefill, "rounds": rounds,
                          "start_at": start_at if self.mode == "device" else None})
        if prefill:
            self.t0_host = start_at
        local = self.backend.run(prefill, rounds, start_at if self.urls and self.mode == "device" else None, mode=self.mode,
                                 on_token=on_token)
        per_node = [local] + self._join(futs)
        self.prefilled = True
        self.rounds_done += rounds
        bad = [r["rank"] for r in per_node if (r["status"][0] & 3) or r["status"][1]]  # bit 4 (sampler overflow) is handled exactly
        if bad:
            raise RingError(f"pipeline aborted: hop watchdog / abort flag set on node(s) {bad} "
                            f"(status words {[r['status'] for r in per_node]})")
        return {"rounds": rounds, "tokens": rounds * self.n_samples, "per_node": per_node,
                "decode_ms": max(r["decode_ms"] for r in per_node), "prefill_ms": max(r["prefill_ms"] for r in per_node)}

    def abort(self) -> None:
        """Poison every node's incoming flags.  Meant to be called from another thread while :meth:`run` is blocked, so
        the posts go out on their OWN threads: the session's pool is busy with the very `run` requests to be aborted."""
        self.backend.abort()
        if not self.urls:
            return
        with ThreadPoolExecutor(max_workers=len(self.urls)) as urgent:
            for f in [urgent.submit(self._post, u, {"op": "abort"}, 5.0) for u in self.urls]:  # a frozen node must not hold up the others
                try:
                    f.result()
                except Exception:  # noqa: BLE001  (a dead node cannot be told)
                    pass

    def tokens(self) -> Dict[int, torch.Tensor]:
        return self.backend.tokens()

    def close(self) -> None:
        if self.pool:
            self.pool.shutdown(wait=False)
