import pickle
import socket
import threading
import time

import pytest
import torch

from mdi_llm_b200.config import HEADERLENGTH
from mdi_llm_b200.parallel.control import ControlServer, HTTPError, http_get_json, request_to_node
from mdi_llm_b200.parallel.transport import (InputNodeConnection, LoopbackTransport, MessageQueue,
                                              OutputNodeConnection, build_msg)
from mdi_llm_b200.parallel.transport.socket_transport import encode_frame
from conftest import free_ports


def test_frame_is_reference_wire_format():
    """16-char left-aligned ASCII decimal length + pickle (connections.py:338-342)."""
    msg = build_msg(torch.arange(6).view(1, 2, 3), 4)
    frame = encode_frame(msg)
    payload = pickle.dumps(msg)
    assert frame[:HEADERLENGTH] == f"{len(payload):<16}".encode()
    assert int(frame[:HEADERLENGTH]) == len(frame) - HEADERLENGTH
    back = pickle.loads(frame[HEADERLENGTH:])
    assert back["sample_index"] == 4 and back["stop"] is False and torch.equal(back["data"], msg["data"])


def test_socket_pair_roundtrip_and_raw_reference_sender():
    p_in, p_out, p_raw = free_ports(3)
    rx_cfg = {"addr": "127.0.0.1", "inference": {"port_in": p_in, "port_out": 0}}
    tx_cfg = {"addr": "127.0.0.1", "inference": {"port_in": 0, "port_out": p_out}}
    q_in, q_out = MessageQueue(), MessageQueue()
    holder = {}
    t = threading.Thread(target=lambda: holder.setdefault("rx", InputNodeConnection(rx_cfg, {"addr": "127.0.0.1"}, q_in)))
    t.start()
    tx = OutputNodeConnection(tx_cfg, rx_cfg, q_out)
    t.join(timeout=5)
    rx = holder["rx"]
    rx.launch()
    tx.launch()
    for i in range(5):
        q_out.put(build_msg(torch.full((1, 1, 8), float(i)), i))
    q_out.put(build_msg("", 9, stop=True))
    got = [q_in.get(timeout=2) for _ in range(6)]
    assert [g["sample_index"] for g in got] == [0, 1, 2, 3, 4, 9]
    assert got[-1]["stop"] and got[-1]["data"] == ""
    assert float(got[3]["data"].sum()) == 24.0
    tx.shutdown()
    rx.shutdown()
    assert tx.n_sent == 6 and rx.n_received == 6 and tx.bytes_sent == rx.bytes_received


def test_loopback_transport_is_fifo():
    t = LoopbackTransport()
    for i in range(3):
        t.send(build_msg(i, i))
    assert [t.recv(0.1)["data"] for _ in range(3)] == [0, 1, 2]
    assert t.recv(0.01) is None


class _App:
    def __init__(self):
        self.got = []

    def GET(self, path, body):
        if not path:
            return '{"hello": "node"}'
        raise HTTPError(404, "Not found")

    def POST(self, path, body):
        if path == ("init",):
            self.got.append(pickle.loads(body))
            return None
        raise HTTPError(404, "Not found")

    def PUT(self, path, body):
        raise HTTPError(501, "PUT not implemented!")


def test_control_server_verbs_and_retrying_client():
    (port,) = free_ports(1)
    app = _App()
    srv = ControlServer(app, "127.0.0.1", port)
    srv.start()
    try:
        assert http_get_json(f"http://127.0.0.1:{port}/") == {"hello": "node"}
        big = {"role": "secondary:0", "params": {"w": torch.randn(256, 256)}}
        assert request_to_node("post", f"http://127.0.0.1:{port}/init", big, max_n_requests=2, retry_wait=0.01) == 1
        assert torch.equal(app.got[0]["params"]["w"], big["params"]["w"])
        assert request_to_node("put", f"http://127.0.0.1:{port}/stop", "", max_n_requests=2, retry_wait=0.01) == 0
        assert request_to_node("post", f"http://127.0.0.1:{port}/nope", {}, max_n_requests=1, retry_wait=0.01) == 0
        with pytest.raises(ValueError):
            request_to_node("delete", "http://x", {})
    finally:
        srv.stop()
    (dead,) = free_ports(1)
    assert request_to_node("post", f"http://127.0.0.1:{dead}/init", {}, max_n_requests=2, retry_wait=0.01) == 0


@pytest.mark.parametrize("world", [2, 3])
def test_torch_distributed_transport_ring_gloo(world):
    """The NCCL/gloo p2p transport under the host-driven scheduler: token-exact vs single device."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (port,) = free_ports(1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "_mp_transport_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=root)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("TR_RESULT ")]
    assert lines, p.stderr[-2000:]
    res = json.loads(lines[-1][len("TR_RESULT "):])
    assert res["ok"] and res["bytes_sent"] > 0 and p.returncode == 0


def test_hostile_frame_on_the_data_port_runs_nothing(tmp_path):
    """A frame whose pickle would execute code (the reference's `pickle.loads` on a TCP port) is dropped by the restricted
    unpickler; the stream goes on with the next well-formed message.  A peer that is not loopback is refused when the
    topology says the previous node is `localhost`."""
    import os
    import pickle
    import socket

    (p_in,) = free_ports(1)
    rx_cfg = {"addr": "127.0.0.1", "inference": {"port_in": p_in, "port_out": 0}}
    q_in = MessageQueue()
    holder = {}
    t = threading.Thread(target=lambda: holder.setdefault("rx", InputNodeConnection(rx_cfg, {"addr": "localhost"}, q_in)))
    t.start()
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return os.system, (f"touch {marker}",)

    s = None
    for _ in range(50):
        try:
            s = socket.create_connection(("127.0.0.1", p_in), timeout=5)
            break
        except OSError:
            time.sleep(0.1)
    assert s is not None
    t.join(timeout=5)
    rx = holder["rx"]
    rx.launch()
    bad = pickle.dumps({"sample_index": 0, "data": Evil(), "stop": False})
    s.sendall(f"{len(bad):<16}".encode() + bad)
    s.sendall(encode_frame(build_msg(torch.ones(1, 1, 4), 3)))
    got = q_in.get(timeout=3)
    assert got["sample_index"] == 3 and float(got["data"].sum()) == 4.0
    assert not marker.exists() and rx.n_received == 1
    s.close()
    rx.shutdown()
