"""TCP + pickle data plane, wire-compatible with the reference.

Wire format (reference ``src/sub/connections.py:325-342``, ``config.py:100``): a 16-character
left-aligned ASCII decimal length, then ``pickle.dumps(message)``.  A node listens on its
``inference.port_in`` for the previous node and connects from ``inference.port_out`` to the
next node's ``port_in`` (``connections.py:120-153,288-320``); RX and TX run on their own threads
feeding :class:`MessageQueue` objects.

Differences from the reference, all behavioural fixes: frames are DECODED with the restricted unpickler
(``utils/safe_pickle.py`` — the reference's ``pickle.loads`` on a TCP port is remote code execution for whoever can
reach it); the ``running`` flag is per connection
(theirs is a class attribute shared by both directions, connections.py:23); ``accept`` has a
time-out; tensors are moved to CPU before pickling so the receiver never needs the sender's
device; queue waits do not poll.
"""
from __future__ import annotations

import pickle
import socket
import threading
import time
from typing import Any, Dict, Optional

import torch

from ...config import HEADERLENGTH
from ...utils.safe_pickle import safe_loads
from .base import ChaosPolicy, Message, MessageQueue, Transport

__all__ = ["InputNodeConnection", "OutputNodeConnection", "SocketTransport", "encode_frame", "read_exact"]


def encode_frame(obj: Any) -> bytes:
    payload = pickle.dumps(obj)
    return f"{len(payload):<{HEADERLENGTH}}".encode("utf-8") + payload


def read_exact(sock: socket.socket, size: int, running: threading.Event) -> bytes:
    buf = bytearray()
    while running.is_set() and len(buf) < size:
        try:
            chunk = sock.recv(size - len(buf))
        except socket.timeout:
            continue
        except OSError:
            running.clear()
            break
        if not chunk:  # peer closed
            running.clear()
            break
        buf += chunk
    return bytes(buf)


def _to_wire(msg: Message) -> Message:
    data = msg.get("data")
    if isinstance(data, torch.Tensor) and data.device.type != "cpu":
        msg = dict(msg)
        msg["data"] = data.detach().cpu()
    return msg


class NodeConnection:
    """Common interface of the two socket endpoints (reference connections.py:15-54): a ``running`` event, a
    ``launch()`` that starts the worker thread and a ``shutdown()`` that stops it."""

    msg_format = {"sample_index": 0, "data": None, "stop": False}
    name = "connection"

    def launch(self) -> None:  # pragma: no cover - overridden
        raise NotImplementedError

    def shutdown(self) -> None:  # pragma: no cover - overridden
        raise NotImplementedError


class InputNodeConnection(NodeConnection):
    """Server side: accept the previous node, then RX thread -> queue."""

    name = "input_queue"

    def __init__(self, config: Dict[str, Any], prev_node: Optional[Dict[str, Any]], queue: MessageQueue,
                 max_tries: int = 30, accept_timeout: float = 120.0, verb: bool = False, **_: Any) -> None:
        if "addr" not in config:
            raise ValueError("Missing IP address in configuration")
        if "inference" not in config or "port_in" not in config["inference"]:
            raise ValueError("Missing input port in configuration")
        self.verb = verb
        self.queue = queue
        self.running = threading.Event()
        self.thread: Optional[threading.Thread] = None
        self.n_received = 0
        self.bytes_received = 0
        self.listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        addr = (config["addr"], int(config["inference"]["port_in"]))
        for attempt in range(max_tries):
            try:
                self.listener.bind(addr)
                break
            except OSError:
                if attempt == max_tries - 1:
                    raise ConnectionError(f"Unable to bind input socket {addr}")
                time.sleep(1)
        self.listener.listen(1)
        self.listener.settimeout(accept_timeout)
        expected = prev_node.get("addr") if prev_node else None
        deadline = time.time() + accept_timeout
        self.conn: Optional[socket.socket] = None
        while self.conn is None:
            try:
                conn, peer = self.listener.accept()
            except socket.timeout:
                raise ConnectionError("Timed out waiting for the previous node to connect") from None
            if (expected is None or peer[0] == expected or expected == "0.0.0.0"
                    or (expected in ("localhost", "::1") and (peer[0].startswith("127.") or peer[0] == "::1"))):
                self.conn = conn  # ("localhost" means a loopback peer, not "anybody")
            else:
                conn.close()
                if time.time() > deadline:
                    raise ConnectionError("Unable to connect to previous node!")
        self.conn.settimeout(0.5)
        self.conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def launch(self) -> None:
        self.running.set()
        self.thread = threading.Thread(target=self.run, name=self.name, daemon=True)
        self.thread.start()

    def run(self) -> None:
        assert self.conn is not None
        while self.running.is_set():
            header = read_exact(self.conn, HEADERLENGTH, self.running)
            if len(header) < HEADERLENGTH:
                break
            try:
                size = int(header)
            except ValueError:
                self.running.clear()
                break
            payload = read_exact(self.conn, size, self.running)
            if len(payload) < size:
                break
            try:
                msg = safe_loads(payload)  # containers, scalars, strings and tensors only: a frame cannot run code here
            except (EOFError, pickle.UnpicklingError, ValueError, TypeError, AttributeError, IndexError, ImportError):
                continue
            self.n_received += 1
            self.bytes_received += HEADERLENGTH + size
            self.queue.put(msg)
        self.queue.close()

    def shutdown(self) -> None:
        self.running.clear()
        if self.thread is not None:
            self.thread.join(timeout=3)
        for s in (self.conn, self.listener):
            try:
                if s is not None:
                    s.close()
            except OSError:
                pass


class OutputNodeConnection(NodeConnection):
    """Client side: connect to the next node, then queue -> TX thread."""

    name = "output_queue"

    def __init__(self, config: Dict[str, Any], next_node: Dict[str, Any], queue: MessageQueue,
                 max_tries: int = 30, verb: bool = False, chaos: Optional[ChaosPolicy] = None, **_: Any) -> None:
        self.verb = verb
        self.queue = queue
        self.chaos = chaos
        self.running = threading.Event()
        self.thread: Optional[threading.Thread] = None
        self.n_sent = 0
        self.bytes_sent = 0
        target = (next_node["addr"], int(next_node["inference"]["port_in"]))
        local = (config["addr"], int(config["inference"]["port_out"]))
        self.sock: Optional[socket.socket] = None
        last: Optional[BaseException] = None
        for _ in range(max_tries):
            s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            try:
                s.bind(local)
                s.connect(target)
                s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                self.sock = s
                break
            except OSError as e:
                last = e
                s.close()
                time.sleep(1)
        if self.sock is None:
            raise ConnectionError(f"Unable to connect to next node! ({last})")

    def send_msg(self, data: Any) -> None:
        assert self.sock is not None
        frame = encode_frame(_to_wire(data) if isinstance(data, dict) else data)
        self.sock.sendall(frame)
        self.n_sent += 1
        self.bytes_sent += len(frame)

    def launch(self) -> None:
        self.running.set()
        self.thread = threading.Thread(target=self.run, name=self.name, daemon=True)
        self.thread.start()

    def run(self) -> None:
        while self.running.is_set():
            msg = self.queue.get(timeout=0.5)
            if msg is None:
                continue
            if self.chaos is not None and not self.chaos.admit(msg):
                continue
            try:
                self.send_msg(msg)
            except OSError:
                self.running.clear()

    def shutdown(self) -> None:
        # drain what is queued (e.g. the last stop marker), then stop
        t0 = time.time()
        while len(self.queue) and self.running.is_set() and time.time() - t0 < 2.0:
            time.sleep(0.01)
        self.running.clear()
        if self.thread is not None:
            self.thread.join(timeout=3)
        try:
            if self.sock is not None:
                self.sock.close()
        except OSError:
            pass


class SocketTransport(Transport):
    """Both connections of a node behind the :class:`Transport` interface.

    Connection order avoids the ring deadlock exactly like the reference
    (``gptserver.py:540-583``): the starter dials OUT first and then accepts; every other node
    accepts first and then dials out.
    """

    name = "socket"

    def __init__(self, own: Dict[str, Any], prev_node: Dict[str, Any], next_node: Dict[str, Any],
                 is_starter: bool, chaos: Optional[ChaosPolicy] = None, verb: bool = False,
                 max_tries: int = 30) -> None:
        super().__init__(chaos)
        self.in_q, self.out_q = MessageQueue(), MessageQueue()
        self.conn_in: Optional[InputNodeConnection] = None
        self.conn_out: Optional[OutputNodeConnection] = None
        if is_starter:
            self.conn_out = OutputNodeConnection(own, next_node, self.out_q, max_tries=max_tries, verb=verb, chaos=chaos)
            self.conn_in = InputNodeConnection(own, prev_node, self.in_q, max_tries=max_tries, verb=verb)
        else:
            self.conn_in = InputNodeConnection(own, prev_node, self.in_q, max_tries=max_tries, verb=verb)
            self.conn_out = OutputNodeConnection(own, next_node, self.out_q, max_tries=max_tries, verb=verb, chaos=chaos)

    def launch(self) -> None:
        super().launch()
        assert self.conn_in and self.conn_out
        self.conn_in.launch()
        self.conn_out.launch()

    def send(self, msg: Message) -> None:
        self.stats["sent"] += 1
        self.out_q.put(msg)

    def recv(self, timeout: Optional[float] = None) -> Optional[Message]:
        msg = self.in_q.get(timeout)
        if msg is not None:
            self.stats["received"] += 1
        return msg

    def shutdown(self) -> None:
        super().shutdown()
        if self.conn_out is not None:
            self.conn_out.shutdown()
            self.stats["bytes_sent"] = self.conn_out.bytes_sent
        if self.conn_in is not None:
            self.conn_in.shutdown()
            self.stats["bytes_received"] = self.conn_in.bytes_received
