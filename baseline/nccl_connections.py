"""Drop-in replacement of the reference's ``sub.connections`` for the *reference NCCL-p2p build*
(BASELINE.md "What therefore has to be measured"): the reference's own node runtime, generation loops,
stage modules and eager PyTorch ops, with ONLY the data-plane classes swapped — ``torch.distributed``
send/recv over NCCL instead of pickle over TCP.  The reference has no such build; this is the
minimal variant: same class names, constructor signatures, queue + event hand-off and thread structure as
``src/sub/connections.py:57-363``, so ``sub.gptserver`` runs unmodified on top of it.

Wire format per message: an int64 header ``[sample_index, stop, T, C, dtype_code]`` then (unless ``stop``)
the activation tensor ``[1, T, C]`` — both sent device-to-device.  One process group per ring edge, so the RX
and TX threads of a node never share a communicator (with two nodes both edges connect the same pair of ranks
and a shared communicator would serialise a send behind the matching recv: deadlock).
"""
from __future__ import annotations

import threading
from collections import deque
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

_DTYPES = [torch.bfloat16, torch.float16, torch.float32]
_groups: Optional[list] = None
_lock = threading.Lock()


def _edge_groups() -> list:
    """edge i = rank i -> rank (i+1) % world; created once, collectively, in rank order."""
    global _groups
    with _lock:
        if _groups is None:
            world = dist.get_world_size()
            _groups = [dist.new_group(ranks=sorted({i, (i + 1) % world})) for i in range(world)]
        return _groups


class NodeConnection:
    msg_format = {"sample_index": 0, "data": None, "stop": False}
    name = "connection"
    verb = False

    def __init__(self, **kwargs: Any) -> None:
        self.running = threading.Event()
        self.verb = bool(kwargs.get("verb", False))
        # CUDA + NCCL on the GPU box; CPU + gloo only for the protocol test of this shim
        self.device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def run(self) -> None:
        pass

    def launch(self, **kwargs: Any) -> None:
        self.running.set()
        self.running_thread = threading.Thread(target=self.run, name=self.name, daemon=True, kwargs=kwargs)
        self.running_thread.start()


class InputNodeConnection(NodeConnection):
    def __init__(self, config: Dict, prev_node: Dict[str, Any], queue: deque, event_callback: threading.Event,
                 max_tries: int = 30, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.name = "input_queue"
        self.message_queue, self.queue_not_empty = queue, event_callback
        self.queue_not_empty.clear()
        self.src = (self.rank - 1) % self.world
        self.group = _edge_groups()[self.src]

    def recv_msg(self) -> Optional[Dict[str, Any]]:
        hdr = torch.zeros(5, dtype=torch.int64, device=self.device)
        dist.recv(hdr, src=self.src, group=self.group)
        sample, stop, T, C, code = hdr.tolist()  # host sync: the reference's loop needs the message on the host side too
        if sample < 0:
            return None  # the sender closed its end
        if stop:
            return {"sample_index": sample, "data": "", "stop": True}
        data = torch.empty(1, T, C, dtype=_DTYPES[code], device=self.device)
        dist.recv(data, src=self.src, group=self.group)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return {"sample_index": sample, "data": data, "stop": False}

    def run(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        while self.running.is_set():
            msg = self.recv_msg()
            if msg is None:
                break
            self.message_queue.append(msg)
            self.queue_not_empty.set()

    def shutdown(self) -> None:
        self.running.clear()
        self.running_thread.join(timeout=3)  # a pending recv ends when the upstream node sends its close header


class OutputNodeConnection(NodeConnection):
    def __init__(self, config: Dict, next_node: Dict[str, Any], queue: deque, event_callback: threading.Event,
                 max_tries: int = 30, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.name = "output_queue"
        self.message_queue, self.queue_not_empty = queue, event_callback
        if len(self.message_queue):
            self.queue_not_empty.set()
        else:
            self.queue_not_empty.clear()
        self.dst = (self.rank + 1) % self.world
        self.group = _edge_groups()[self.rank]

    def send_msg(self, data: Any) -> None:
        t = data["data"]
        if data["stop"] or not isinstance(t, torch.Tensor):
            hdr = torch.tensor([data["sample_index"], 1, 0, 0, 0], dtype=torch.int64, device=self.device)
            dist.send(hdr, dst=self.dst, group=self.group)
            return
        t = t.to(self.device).contiguous()
        hdr = torch.tensor([data["sample_index"], 0, t.shape[-2], t.shape[-1], _DTYPES.index(t.dtype)], dtype=torch.int64,
                           device=self.device)
        dist.send(hdr, dst=self.dst, group=self.group)
        dist.send(t.view(1, t.shape[-2], t.shape[-1]), dst=self.dst, group=self.group)

    def run(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        while self.running.is_set():
            if self.queue_not_empty.wait(timeout=2):
                tx_msg = self.message_queue.popleft()
                if len(self.message_queue) < 1:
                    self.queue_not_empty.clear()
                self.send_msg(tx_msg)

    def shutdown(self) -> None:
        self.running.clear()
        self.running_thread.join(timeout=5)
        try:  # close header: lets the downstream RX thread leave its blocking recv
            if self.device.type == "cuda":
                torch.cuda.set_device(self.device)
            dist.send(torch.tensor([-1, 0, 0, 0, 0], dtype=torch.int64, device=self.device), dst=self.dst, group=self.group)
            if self.device.type == "cuda":
                torch.cuda.current_stream().synchronize()
        except Exception:  # noqa: BLE001
            pass
