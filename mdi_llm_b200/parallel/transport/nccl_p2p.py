"""``torch.distributed`` point-to-point transport (NCCL on GPUs, gloo on CPU).

The *baseline* data plane the product is compared against ("a path that only calls NCCL for the
inter-stage hop is the baseline, not the product"): the reference has no such path at all — its
only transport is pickle over TCP (``src/sub/connections.py``) — so this is the midpoint between
the socket transport and the fused P2P-store hop of ``parallel/pipeline.py``.

Messages keep the reference's schema.  Wire form: an int64 header ``[sample_index, stop, ndim,
d0, d1, d2, dtype_code]`` followed (unless ``stop``) by the payload tensor.  Every ring edge
``i -> i+1`` gets its own process group, hence its own NCCL communicator and stream: with a single
communicator a 2-node ring deadlocks (each rank's ``recv`` is queued behind its own blocked ``send``).
"""
from __future__ import annotations

import threading
from typing import Any, List, Optional

import torch
import torch.distributed as dist

from .base import ChaosPolicy, Message, MessageQueue, Transport, build_msg

__all__ = ["TorchDistTransport", "make_edge_groups"]

_DTYPES = [torch.float32, torch.bfloat16, torch.float16, torch.int32, torch.int64]


def make_edge_groups(world: int) -> List[Any]:
    """One group per directed ring edge ``i -> (i+1) % world`` (collective: call on every rank)."""
    if world == 1:
        return [None]
    return [dist.new_group([i, (i + 1) % world] if i != (i + 1) % world else [i]) for i in range(world)]


class TorchDistTransport(Transport):
    name = "torch.distributed"

    def __init__(self, rank: int, world: int, device: torch.device, edge_groups: Optional[List[Any]] = None,
                 chaos: Optional[ChaosPolicy] = None) -> None:
        super().__init__(chaos)
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.prev, self.next = (rank - 1) % world, (rank + 1) % world
        groups = edge_groups if edge_groups is not None else make_edge_groups(world)
        self.g_out, self.g_in = groups[rank], groups[self.prev]
        self.in_q = MessageQueue()
        self.thread: Optional[threading.Thread] = None
        self._pending: List[Any] = []

    def launch(self) -> None:
        super().launch()
        self.thread = threading.Thread(target=self._rx_loop, name="dist-rx", daemon=True)
        self.thread.start()

    def _rx_loop(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        while self.running.is_set():
            header = torch.zeros(7, dtype=torch.int64, device=self.device)
            try:
                dist.recv(header, src=self.prev, group=self.g_in)
            except Exception:  # noqa: BLE001  (group destroyed at shutdown)
                break
            h = header.tolist()
            if h[0] < 0:  # shutdown sentinel
                break
            if h[1]:
                self.in_q.put(build_msg("", h[0], stop=True))
                continue
            shape = [int(x) for x in h[3:3 + h[2]]]
            data = torch.empty(shape, dtype=_DTYPES[h[6]], device=self.device)
            dist.recv(data, src=self.prev, group=self.g_in)
            self.stats["bytes_received"] += data.numel() * data.element_size() + 56
            self.in_q.put(build_msg(data, h[0]))
        self.in_q.close()

    def send(self, msg: Message) -> None:
        if self.chaos is not None and not self.chaos.admit(msg):
            return
        data = msg.get("data")
        stop = bool(msg.get("stop"))
        h = [int(msg["sample_index"]), int(stop), 0, 0, 0, 0, 0]
        if not stop:
            data = data.to(self.device).contiguous()
            h[2] = data.dim()
            h[3:3 + data.dim()] = list(data.shape)
            h[6] = _DTYPES.index(data.dtype)
        self._pending = [w for w in self._pending if not w.is_completed()]
        self._pending.append(dist.isend(torch.tensor(h, dtype=torch.int64, device=self.device), dst=self.next, group=self.g_out))
        if not stop:
            self._pending.append(dist.isend(data, dst=self.next, group=self.g_out))
            self.stats["bytes_sent"] += data.numel() * data.element_size() + 56
        self.stats["sent"] += 1

    def recv(self, timeout: Optional[float] = None) -> Optional[Message]:
        msg = self.in_q.get(timeout)
        if msg is not None:
            self.stats["received"] += 1
        return msg

    def shutdown(self) -> None:
        if self.running.is_set():
            super().shutdown()
            try:  # unblock the next rank's RX thread
                dist.isend(torch.tensor([-1, 0, 0, 0, 0, 0, 0], dtype=torch.int64, device=self.device),
                           dst=self.next, group=self.g_out).wait()
            except Exception:  # noqa: BLE001
                pass
            for w in self._pending:
                try:
                    w.wait()
                except Exception:  # noqa: BLE001
                    pass
            if self.thread is not None:
                self.thread.join(timeout=5)
