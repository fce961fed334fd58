"""Transport interface of the inter-stage data plane.

Every transport moves *messages* ``{"sample_index": int, "data": Tensor | "", "stop": bool}``
(the reference's message format, ``src/sub/connections.py:21`` / ``gptserver.py:585-597``) from
one pipeline stage to the next one in the ring.  Implementations:

* :mod:`.socket_transport` — TCP + pickle, wire-compatible with the reference;
* :mod:`.inproc`           — loop-back queue (standalone, 1 node: ``gptserver.py:276-278``);
* :mod:`.nccl_p2p`         — ``torch.distributed`` isend/irecv on a side stream (baseline).

The product data plane is not a ``Transport`` at all: on one NVSwitch box the activations never pass through the
host, so there is nothing to ``send``/``recv`` — the hop is fused into the kernels and wired by
:mod:`mdi_llm_b200.parallel.ring` (CUDA-IPC handles exchanged over the control plane, peer stores + flags issued by
the last kernel of a stage, see ``parallel/pipeline.py`` and ``ops/csrc/common.cuh``).

A :class:`ChaosPolicy` can be attached to any host-side transport to delay or drop messages —
the fault-injection hook the reference lacks.
"""
from __future__ import annotations

import random
import threading
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Any, Deque, Dict, Optional

__all__ = ["Message", "build_msg", "Transport", "MessageQueue", "ChaosPolicy", "TransportError"]

Message = Dict[str, Any]


class TransportError(RuntimeError):
    pass


def build_msg(data: Any, sample_index: int, stop: bool = False) -> Message:
    return {"sample_index": int(sample_index), "data": data, "stop": bool(stop)}


@dataclass
class ChaosPolicy:
    """Fault injection for tests: each message is delayed by ``delay_s`` with probability
    ``p_delay`` and dropped with probability ``p_drop`` (stop markers are never dropped unless
    ``drop_stop``)."""

    p_delay: float = 0.0
    delay_s: float = 0.0
    p_drop: float = 0.0
    drop_stop: bool = False
    seed: int = 0
    dropped: int = 0
    delayed: int = 0
    _rng: random.Random = field(default_factory=random.Random, repr=False)

    def __post_init__(self) -> None:
        self._rng.seed(self.seed)

    def admit(self, msg: Message) -> bool:
        if self.p_delay and self._rng.random() < self.p_delay:
            self.delayed += 1
            time.sleep(self.delay_s)
        if self.p_drop and (self.drop_stop or not msg.get("stop")) and self._rng.random() < self.p_drop:
            self.dropped += 1
            return False
        return True


class MessageQueue:
    """FIFO + condition variable.  Unlike the reference's ``deque`` + ``Event`` pair
    (gptserver.py:112-118) emptiness test and wait are atomic, so no 2 s poll is needed."""

    def __init__(self) -> None:
        self._q: Deque[Message] = deque()
        self._cv = threading.Condition()
        self._closed = False

    def put(self, msg: Message) -> None:
        with self._cv:
            self._q.append(msg)
            self._cv.notify()

    def get(self, timeout: Optional[float] = None) -> Optional[Message]:
        with self._cv:
            if not self._q:
                if self._closed:
                    return None
                self._cv.wait(timeout)
            if self._q:
                return self._q.popleft()
            return None

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()

    def __len__(self) -> int:
        with self._cv:
            return len(self._q)


class Transport:
    """One direction-pair endpoint of a node: ``send`` goes to the next node of the ring,
    ``recv`` yields what the previous node sent."""

    name = "transport"

    def __init__(self, chaos: Optional[ChaosPolicy] = None) -> None:
        self.chaos = chaos
        self.running = threading.Event()  # per instance (the reference shares one class-level Event)
        self.stats = {"sent": 0, "received": 0, "bytes_sent": 0, "bytes_received": 0}

    def launch(self) -> None:
        self.running.set()

    def send(self, msg: Message) -> None:
        raise NotImplementedError

    def recv(self, timeout: Optional[float] = None) -> Optional[Message]:
        raise NotImplementedError

    def shutdown(self) -> None:
        self.running.clear()
