"""Loop-back transport: what a node sends is what it receives next.

Standalone (single-node) operation of the reference aliases the out-queue to the in-queue
(``gptserver.py:276-278``); this is the same thing as a :class:`Transport`.  Also used by unit
tests to chain several stage runners inside one process (``pair()``).
"""
from __future__ import annotations

from typing import List, Optional

from .base import ChaosPolicy, Message, MessageQueue, Transport

__all__ = ["LoopbackTransport", "QueueTransport", "ring"]


class QueueTransport(Transport):
    """Endpoints joined by in-memory queues: ``send`` puts into ``out_q``, ``recv`` pops ``in_q``."""

    name = "inproc"

    def __init__(self, in_q: MessageQueue, out_q: MessageQueue, chaos: Optional[ChaosPolicy] = None) -> None:
        super().__init__(chaos)
        self.in_q, self.out_q = in_q, out_q

    def send(self, msg: Message) -> None:
        if self.chaos is not None and not self.chaos.admit(msg):
            return
        self.stats["sent"] += 1
        self.out_q.put(msg)

    def recv(self, timeout: Optional[float] = None) -> Optional[Message]:
        msg = self.in_q.get(timeout)
        if msg is not None:
            self.stats["received"] += 1
        return msg

    def shutdown(self) -> None:
        super().shutdown()
        self.in_q.close()


class LoopbackTransport(QueueTransport):
    name = "loopback"

    def __init__(self, chaos: Optional[ChaosPolicy] = None) -> None:
        q = MessageQueue()
        super().__init__(q, q, chaos)


def ring(n: int) -> List[QueueTransport]:
    """``n`` endpoints wired as a closed ring: endpoint ``i`` sends to ``(i+1) % n``."""
    queues = [MessageQueue() for _ in range(n)]  # queues[i] = input of node i
    return [QueueTransport(queues[i], queues[(i + 1) % n]) for i in range(n)]
