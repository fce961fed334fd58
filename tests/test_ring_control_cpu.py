"""CPU tests of the pieces around the device ring that do not need a GPU: the restricted unpickler, the
control plane's token / loopback policy and response bodies, the ring protocol (with a recording stand-in for
the CUDA backend), stage-shape inference of half-layer chunks, partition-independent random weights, and the
one-box launcher."""
import json
import os
import pickle
import subprocess
import sys
import threading
import time

import pytest
import torch

from conftest import free_ports
from mdi_llm_b200.models.config import Config
from mdi_llm_b200.parallel.control import TOKEN_HEADER as TOKEN_HEADER_NAME
from mdi_llm_b200.parallel.control import ControlServer, HTTPError, call_node, request_to_node
from mdi_llm_b200.utils.safe_pickle import UnsafePayload, safe_loads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Evil:
    def __reduce__(self):
        return (os.system, ("echo pwned > /tmp/mdi_pwned",))


def test_safe_loads_accepts_messages_and_rejects_code():
    msg = {"sample_index": 3, "data": torch.randn(1, 2, 8).to(torch.bfloat16), "stop": False,
           "params": {"w": torch.nn.Parameter(torch.ones(3))}, "cfg": {"a": [1, 2.5, None, "x", (1, 2)]}}
    back = safe_loads(pickle.dumps(msg))
    assert torch.equal(back["data"], msg["data"]) and back["cfg"] == msg["cfg"] and back["sample_index"] == 3
    assert torch.equal(back["params"]["w"], msg["params"]["w"])
    with pytest.raises(UnsafePayload):
        safe_loads(pickle.dumps({"x": _Evil()}))
    assert not os.path.exists("/tmp/mdi_pwned")


class _App:
    def __init__(self):
        self.n = 0

    def POST(self, path, body):
        self.n += 1
        if path == ("init",):
            return {"handles": {"hidden": b"\x01" * 64, "flag_off": 256}, "echo": safe_loads(body)["role"]}
        if path == ("boom",):
            raise HTTPError(500, "failed on purpose")
        raise HTTPError(404, "Not found")


def test_control_plane_token_bodies_and_no_blind_retries(monkeypatch):
    (port,) = free_ports(1)
    app = _App()
    srv = ControlServer(app, "127.0.0.1", port, token="s3cret")
    srv.start()
    try:
        url = f"http://127.0.0.1:{port}"
        status, body = call_node("post", url + "/init", {"role": "secondary:0"}, max_n_requests=1, token="s3cret")
        assert status == 200 and body["echo"] == "secondary:0" and body["handles"]["hidden"] == b"\x01" * 64
        status, body = call_node("post", url + "/init", {"role": "x"}, max_n_requests=1)  # no token
        assert status == 401 and app.n == 1
        status, body = call_node("post", url + "/boom", {}, max_n_requests=50, retry_wait=5.0, token="s3cret")
        assert status == 500 and "on purpose" in body and app.n == 2  # a node that answered is not asked again
        monkeypatch.setenv("MDI_CLUSTER_TOKEN", "s3cret")
        assert request_to_node("post", url + "/init", {"role": "y"}, max_n_requests=1) == 1
    finally:
        srv.stop()
    monkeypatch.delenv("MDI_CLUSTER_TOKEN")
    with pytest.raises(PermissionError):
        ControlServer(app, "10.1.2.3", port)  # non-loopback bind without a shared secret is refused up front


def test_stage_shape_and_partition_independent_random_weights(tiny_llama_cfg):
    from mdi_llm_b200.models.partition import (split_parameters_half, stage_shape_from_state_dict, stage_specs)
    from mdi_llm_b200.models.stage import build_stage
    from mdi_llm_b200.utils.checkpoint import random_init_stage_, random_state_dict

    cfg = tiny_llama_cfg
    sd = random_state_dict(cfg, dtype=torch.float32)
    chunks = split_parameters_half(dict(sd), [3, 4, 3])
    shapes = [stage_shape_from_state_dict(c) for c in [chunks["starter"]] + chunks["secondary"]]
    assert [(s["n_blocks"], s["first_parts"], s["last_parts"]) for s in shapes] == [(2, "both", "attn"), (3, "mlp", "attn"), (2, "mlp", "both")]
    # third-layer units: cuts between attention | gate/up | down; chunk files describe their own shape
    from mdi_llm_b200.models.partition import split_parameters_units, third_stages

    tspecs = third_stages([5, 5, 5])  # 15 units = 5 layers: [L0, L1.attn, L1.gu] [L1.down, L2, L3.attn] [L3.gu, L3.down, L4]
    tchunks = split_parameters_units(dict(sd), tspecs)
    tshapes = [stage_shape_from_state_dict(c) for c in [tchunks["starter"]] + tchunks["secondary"]]
    assert [(s["n_blocks"], s["first_parts"], s["last_parts"]) for s in tshapes] == [(2, "both", "attn_gu"), (3, "down", "attn"), (2, "mlp", "both")]
    assert [(s["n_blocks"], s["first_parts"], s["last_parts"]) for s in tspecs] == [(2, "both", "attn_gu"), (3, "down", "attn"), (2, "mlp", "both")]
    assert "transformer.h.0.mlp.proj.weight" in tchunks["secondary"][0] and "transformer.h.1.mlp.fc_1.weight" in tchunks["starter"]
    # the same seed gives the same model whatever the partition
    full = build_stage(cfg, "starter", cfg.n_layer, meta=True)
    random_init_stage_(full, "cpu", torch.float32, seed=11)
    ref = full.state_dict()
    for policy in ("auto", "half", "third"):
        specs = stage_specs(3, cfg, policy)
        assert abs(sum(s["layers"] for s in specs) - cfg.n_layer) < 0.02
        for i, sp in enumerate(specs):
            st = build_stage(cfg, "starter" if i == 0 else f"secondary:{i - 1}", sp["n_blocks"], meta=True,
                             first_parts=sp["first_parts"], last_parts=sp["last_parts"])
            random_init_stage_(st, "cpu", torch.float32, seed=11, layer_offset=sp["layer_offset"])
            for k, v in st.state_dict().items():
                if k.startswith("transformer.h."):
                    _, _, li, tail = k.split(".", 3)
                    k = f"transformer.h.{int(li) + sp['layer_offset']}.{tail}"
                assert torch.equal(v, ref[k]), (policy, i, k)


class _FakeBackend:
    """Stands in for RingBackend on a machine without CUDA: records the protocol."""

    def __init__(self, rank, log):
        self.rank, self.log, self.n_samples, self.world = rank, log, 2, 3
        self.prepared = False

    def handle(self, msg):
        from mdi_llm_b200.parallel.ring import RingBackend

        return RingBackend.handle(self, msg)

    def connect(self, nxt):
        self.log.append((self.rank, "connect", nxt["rank"]))

    def prepare(self, lens, max_new, prompts=None):
        self.prepared = True
        self.log.append((self.rank, "prepare", tuple(lens), max_new, prompts is not None))

    def run(self, prefill, rounds, start_at=None, mode="device", on_token=None):
        assert self.prepared
        self.log.append((self.rank, "run", prefill, rounds, mode if self.rank == 0 else "device"))
        return {"rank": self.rank, "prefill_ms": 1.0 * prefill, "decode_ms": 0.5 * rounds + self.rank, "steps": rounds * 2,
                "kernel_launches": rounds * 14, "wait_cycles": 0, "status": [0, 0], "h2d": 0, "d2h": 0, "round": rounds}

    def abort(self):
        self.log.append((self.rank, "abort"))

    def tokens(self):
        return {0: torch.zeros(1, 4, dtype=torch.int64)}


class _Node:
    def __init__(self, backend):
        self.ring = backend

    def POST(self, path, body):
        assert path == ("ring",)
        return self.ring.handle(safe_loads(body))


def test_ring_session_protocol_over_http():
    """prepare everywhere before anything runs; prefill only in the first segment; per-node device times come
    back; a node reporting an abort turns into RingError on the starter."""
    from mdi_llm_b200.parallel.ring import RingError, RingSession

    log = []
    ports = free_ports(2)
    nodes = [{"addr": "127.0.0.1", "communication": {"port": p}} for p in ports]
    backends = [_FakeBackend(i + 1, log) for i in range(2)]
    servers = [ControlServer(_Node(b), "127.0.0.1", p) for b, p in zip(backends, ports)]
    for s in servers:
        s.start()
    try:
        local = _FakeBackend(0, log)
        prompts = [torch.tensor([1, 2, 3]), torch.tensor([4, 5])]
        sess = RingSession(local, nodes, prompts, 6, mode="host")
        assert sorted(e for e in log if e[1] == "prepare") == [(0, "prepare", (3, 2), 6, True), (1, "prepare", (3, 2), 6, False),
                                                               (2, "prepare", (3, 2), 6, False)]
        r1 = sess.run(2)
        r2 = sess.run()
        assert r1["tokens"] == 4 and r1["decode_ms"] == 3.0 and r1["prefill_ms"] == 1.0 and len(r1["per_node"]) == 3
        assert r2["rounds"] == 4 and r2["prefill_ms"] == 0.0
        runs = [e for e in log if e[1] == "run"]
        assert sorted(runs[:3]) == [(0, "run", True, 2, "host"), (1, "run", True, 2, "device"), (2, "run", True, 2, "device")]
        assert all(e[2] is False and e[3] == 4 for e in runs[3:])
        backends[1].run = lambda *a, **k: {"rank": 2, "prefill_ms": 0.0, "decode_ms": 0.0, "steps": 0, "kernel_launches": 0,
                                           "wait_cycles": 0, "status": [1, 1], "h2d": 0, "d2h": 0, "round": 0}
        sess.max_new = 99
        with pytest.raises(RingError, match="node"):
            sess.run(1)
        sess.abort()
        assert {e[0] for e in log if e[1] == "abort"} == {0, 1, 2}
        sess.close()
    finally:
        for s in servers:
            s.stop()


def test_one_box_launcher_cpu(tmp_path, tiny_llama_cfg):
    """`python -m mdi_llm_b200.cli.launch`: 3 node processes on loopback (CPU, socket transport), two runs,
    run statistics appended, no stragglers (parity: old/nanoGPT/test_mdi_local.sh)."""
    from mdi_llm_b200.utils.checkpoint import write_random_checkpoint

    ck = write_random_checkpoint(tmp_path / "custom" / "NanoLlama", tiny_llama_cfg, dtype=torch.float32)
    stats = tmp_path / "runs.csv"
    env = dict(os.environ, MDI_LOGS_DIR=str(tmp_path / "logs"), MDI_IMG_DIR=str(tmp_path / "img"))
    cmd = [sys.executable, "-m", "mdi_llm_b200.cli.launch", "--ckpt", str(ck), "--n-nodes", "3", "--device", "cpu", "--dtype", "float32",
           "--runs", "2", "--", "--n-samples", "3", "--n-tokens", "4", "--prompt", "Hi", "--greedy", "--time-run", str(stats)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert p.stdout.count("Sample 3:") == 2 and "=== run 2/2" in p.stdout
    assert len(stats.read_text().strip().splitlines()) == 3  # header + one row per run


def test_abort_is_served_while_a_run_is_in_flight():
    """The abort path depends on the control server answering `POST /ring abort` while the same node's `run` request
    is still blocked on its GPU: a node whose run only returns once it has been aborted must not dead-lock the session."""
    import threading

    from mdi_llm_b200.parallel.ring import RingError, RingSession

    log = []
    (port,) = free_ports(1)
    released = threading.Event()

    class Blocking(_FakeBackend):
        def run(self, prefill, rounds, start_at=None, mode="device", on_token=None):
            ok = released.wait(timeout=20)  # "spinning on a flag that never comes"
            out = super().run(prefill, rounds, start_at, mode, on_token)
            out["status"] = [0, 1] if ok else [1, 1]
            return out

        def abort(self):
            super().abort()
            released.set()

    remote = Blocking(1, log)
    server = ControlServer(_Node(remote), "127.0.0.1", port)
    server.start()
    try:
        local = _FakeBackend(0, log)
        sess = RingSession(local, [{"addr": "127.0.0.1", "communication": {"port": port}}], [torch.tensor([1, 2])], 4)
        threading.Timer(0.3, sess.abort).start()
        t0 = time.time()
        with pytest.raises(RingError, match="aborted"):
            sess.run(2)
        assert time.time() - t0 < 10  # released by the abort, not by the 20 s stand-in for the watchdog
        assert (1, "abort") in log and (0, "abort") in log
        sess.close()
    finally:
        server.stop()


def test_failed_ring_generation_poisons_before_raising():
    """`GPTServer._starter_ring`: whatever makes the generation fail (here: a node reporting an abort), the session is
    told to abort — so nothing keeps spinning on the other GPUs — and closed, and the error still reaches the caller."""
    import types

    from mdi_llm_b200.parallel.ring import RingError
    from mdi_llm_b200.parallel.server import GPTServer

    calls = []

    class Sess:
        mode, prompt_lens, t0_host = "device", [2], 0.0

        def run(self, on_token=None):
            calls.append("run")
            raise RingError("pipeline aborted: node 1")

        def abort(self):
            calls.append("abort")

        def close(self):
            calls.append("close")

    fake = types.SimpleNamespace(open_ring_session=lambda n, p, m: Sess(), on_token=None, ring=None)
    with pytest.raises(RingError, match="node 1"):
        GPTServer._starter_ring(fake, 1, "x", 4)
    assert calls == ["run", "abort", "close"]


def test_ring_generation_success_path_collects_tokens_and_timeline():
    import types

    from mdi_llm_b200.parallel.server import GPTServer

    calls = []

    class Sess:
        mode, prompt_lens, t0_host = "device", [2, 2], 0.0

        def run(self, on_token=None):
            calls.append("run")
            return {"rounds": 3, "tokens": 6, "decode_ms": 1.0, "prefill_ms": 0.5, "per_node": []}

        def tokens(self):
            return {0: torch.tensor([[5, 6, 7, 8, 9]]), 1: torch.tensor([[5, 6, 1, 2, 3]])}

        def abort(self):
            calls.append("abort")

        def close(self):
            calls.append("close")

    fake = types.SimpleNamespace(open_ring_session=lambda n, p, m: Sess(), on_token=None,
                                 ring=types.SimpleNamespace(token_times=lambda: [0.01 * i for i in range(1, 7)]),
                                 stop_tokens=(), tok=types.SimpleNamespace(decode=lambda t: " ".join(str(int(x)) for x in t.reshape(-1))))
    texts, tok_time = GPTServer._starter_ring(fake, 2, "x", 3)
    assert calls == ["run", "close"] and texts == ["5 6 7 8 9", "5 6 1 2 3"]
    assert tok_time[0] == (0, 0.0) and tok_time[-1] == (6, pytest.approx(0.06)) and fake.last_result.n_tokens == 6


def test_ring_session_validates_the_sample_count():
    """Same contract as the socket path (gptserver.py:816-821): fewer samples than nodes only warns, zero raises."""
    import types
    import warnings

    from mdi_llm_b200.parallel.server import GPTServer

    fake = types.SimpleNamespace(ring=object(), model=object(), n_nodes=3)
    with pytest.raises(ValueError, match="less than 1 sample"):
        GPTServer.open_ring_session(fake, 0, "x", 4)
    fake.ring = types.SimpleNamespace(pipe=types.SimpleNamespace(set_sampling=lambda s: (_ for _ in ()).throw(StopIteration())))
    fake.sampling = None
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(StopIteration):  # the fake stops right after the checks
            GPTServer.open_ring_session(fake, 2, "x", 4)
    assert any("will not be efficient" in str(x.message) for x in w)


def test_control_server_survives_malformed_and_unauthenticated_requests():
    """Hardening of the per-node endpoint: the token is checked before a body is read (a stranger cannot make a node buffer
    gigabytes), garbage on the wire never takes the server down, and a well-formed request is served afterwards."""
    import http.client
    import socket

    class App:
        def __init__(self):
            self.bodies = []

        def GET(self, path, body):
            return json.dumps({"ok": True})

        def POST(self, path, body):
            self.bodies.append(len(body))
            return {"n": len(body)}

    (port,) = free_ports(1)
    app = App()
    server = ControlServer(app, "127.0.0.1", port, token="s3cret")
    server.start()
    try:
        # 1. no token, 1 GB announced, nothing sent: answered at once with 401, the body is never awaited
        with socket.create_connection(("127.0.0.1", port), timeout=5) as s:
            s.sendall(b"POST /init HTTP/1.1\r\nHost: x\r\nContent-Length: 1073741824\r\n\r\n")
            s.settimeout(5)
            reply = s.recv(4096)
        assert reply.startswith(b"HTTP/1.1 401") and b"Connection: close" in reply and app.bodies == []
        # 2. malformed lengths, unknown verbs, binary noise
        for raw in (b"POST /init HTTP/1.1\r\nHost: x\r\nX-MDI-Token: s3cret\r\nContent-Length: minus-one\r\n\r\n",
                    b"POST /init HTTP/1.1\r\nHost: x\r\nX-MDI-Token: s3cret\r\nContent-Length: -5\r\n\r\n",
                    b"BREW /coffee HTTP/1.1\r\nHost: x\r\n\r\n", b"\x00\xff\xfe garbage \r\n\r\n", b"GET\r\n\r\n"):
            with socket.create_connection(("127.0.0.1", port), timeout=5) as s:
                s.sendall(raw)
                s.settimeout(5)
                try:
                    reply = s.recv(4096)
                except (socket.timeout, ConnectionError):
                    reply = b""
            assert not reply.startswith(b"HTTP/1.1 200"), raw
        # 3. wrong token of the right length, then a pickle that is not a plain message
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
        conn.request("POST", "/init", body=b"x" * 10, headers={TOKEN_HEADER_NAME: "s3creT"})
        assert conn.getresponse().status == 401
        conn.close()
        # 4. the server is still healthy
        status, body = call_node("post", f"http://127.0.0.1:{port}/init", {"hello": 1}, token="s3cret", max_n_requests=1)
        assert status == 200 and body["n"] > 0 and len(app.bodies) == 1
        conn = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
        conn.request("GET", "/", headers={TOKEN_HEADER_NAME: "s3cret"})
        resp = conn.getresponse()
        assert resp.status == 200 and json.loads(resp.read()) == {"ok": True}
        conn.close()
    finally:
        server.stop()
