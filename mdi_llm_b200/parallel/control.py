"""Control plane: the per-node HTTP endpoint and its client.

Parity: the reference mounts every ``GPTServer`` on a CherryPy ``MethodDispatcher``
(``src/sub/gptserver.py:328-350``) exposing ``GET /`` (node config as JSON), ``POST /init``
(pickled init message, optionally carrying the model chunk), ``PUT /stop``, ``DELETE`` (501)
(``:1114-1226``), and talks to it with ``requests`` + retries (``model_dist.py:499-573``).

CherryPy is not available on the target image, so the same verbs/paths/bodies are served by
the standard library (``http.server.ThreadingHTTPServer``), with no body-size limit (model
chunks are GBs, gptserver.py:345).  Bodies stay pickled dicts for protocol compatibility, but they are
*decoded* with :func:`~mdi_llm_b200.utils.safe_pickle.safe_loads` (containers + tensors only), a
shared secret (``MDI_CLUSTER_TOKEN`` / ``token=``) is checked on every request when configured, and a
node refuses to listen on a non-loopback address without one unless ``insecure=True`` /
``MDI_ALLOW_INSECURE=1`` (the reference's trust model: anybody who reaches the port owns the node).

Requests may return a body: handlers that return a ``dict`` answer with its pickle — this is how a
secondary hands the CUDA-IPC handles of its hop buffers back to the starter at ``POST /init``.
"""
from __future__ import annotations

import hmac
import json
import os
import pickle
import threading
import time
import urllib.error
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Callable, Dict, Optional, Tuple

__all__ = ["ControlServer", "HTTPError", "request_to_node", "http_get_json", "call_node", "is_loopback", "TOKEN_HEADER"]

TOKEN_HEADER = "X-MDI-Token"


def is_loopback(addr: str) -> bool:
    return addr in ("localhost", "::1") or addr.startswith("127.")


def cluster_token(explicit: Optional[str] = None) -> Optional[str]:
    return explicit if explicit else (os.environ.get("MDI_CLUSTER_TOKEN") or None)


class HTTPError(Exception):
    def __init__(self, status: int, message: str = "") -> None:
        super().__init__(f"{status} {message}")
        self.status, self.message = status, message


Handler = Callable[[Tuple[str, ...], bytes], Any]


class ControlServer:
    """Threaded HTTP server dispatching on the verb to ``app.GET/POST/PUT/DELETE(path, body)``.

    A handler returns ``None`` (=> 200, empty), ``bytes``/``str`` (=> 200 with that body) or
    raises :class:`HTTPError`.
    """

    def __init__(self, app: Any, host: str, port: int, token: Optional[str] = None, insecure: Optional[bool] = None) -> None:
        self.app = app
        self.token = cluster_token(token)
        if insecure is None:
            insecure = os.environ.get("MDI_ALLOW_INSECURE", "") not in ("", "0")
        if not is_loopback(str(host)) and self.token is None and not insecure:
            raise PermissionError(
                f"refusing to serve the control plane on {host}:{port} without a shared secret: set MDI_CLUSTER_TOKEN "
                "on every node (or pass insecure=True / MDI_ALLOW_INSECURE=1 to accept unauthenticated peers)")
        outer = self

        class _Req(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, fmt: str, *args: Any) -> None:  # silence default stderr log
                pass

            def _dispatch(self, verb: str) -> None:
                path = tuple(p for p in self.path.split("?")[0].split("/") if p)
                status, payload, ctype = 200, b"", "text/plain"
                unread = False  # the request body was not consumed: this connection cannot be reused
                try:
                    # authenticate BEFORE reading the body (an init message may be GBs: a stranger must not make the node
                    # buffer one), in constant time
                    if outer.token is not None and not hmac.compare_digest((self.headers.get(TOKEN_HEADER) or "").encode("utf-8"),
                                                                           outer.token.encode("utf-8")):
                        unread = True
                        raise HTTPError(401, "bad or missing cluster token")
                    try:
                        length = int(self.headers.get("Content-Length") or 0)
                    except ValueError:
                        length = -1
                    if length < 0:
                        unread = True
                        raise HTTPError(400, "bad Content-Length")
                    body = self.rfile.read(length) if length else b""
                    fn = getattr(outer.app, verb, None)
                    if fn is None:
                        raise HTTPError(501, f"{verb} not implemented!")
                    out = fn(path, body)
                    if isinstance(out, str):
                        payload, ctype = out.encode("utf-8"), "application/json"
                    elif isinstance(out, bytes):
                        payload, ctype = out, "application/octet-stream"
                    elif isinstance(out, dict):
                        payload, ctype = pickle.dumps(out), "application/x-pickle"
                except HTTPError as e:
                    status, payload = e.status, e.message.encode("utf-8")
                except Exception as e:  # noqa: BLE001
                    status, payload = 500, repr(e).encode("utf-8")
                self.send_response(status)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(payload)))
                if unread:
                    self.send_header("Connection", "close")
                    self.close_connection = True
                self.end_headers()
                if payload:
                    self.wfile.write(payload)

            def do_GET(self) -> None:  # noqa: N802
                self._dispatch("GET")

            def do_POST(self) -> None:  # noqa: N802
                self._dispatch("POST")

            def do_PUT(self) -> None:  # noqa: N802
                self._dispatch("PUT")

            def do_DELETE(self) -> None:  # noqa: N802
                self._dispatch("DELETE")

        ThreadingHTTPServer.allow_reuse_address = True
        self.httpd = ThreadingHTTPServer((host, int(port)), _Req)
        self.httpd.daemon_threads = True
        self.thread = threading.Thread(target=self.httpd.serve_forever, name="control-http", daemon=True)
        self._stopped = threading.Event()

    def start(self) -> None:
        self.thread.start()

    def stop(self) -> None:
        if self._stopped.is_set():
            return
        self._stopped.set()
        self.httpd.shutdown()
        self.httpd.server_close()

    def block(self) -> None:
        """Park the calling thread until :meth:`stop` (CherryPy's ``engine.block()``)."""
        try:
            while not self._stopped.wait(0.5):
                pass
        except KeyboardInterrupt:
            raise


def _http(method: str, addr: str, data: Optional[bytes], timeout: float, token: Optional[str] = None) -> Tuple[int, bytes]:
    req = urllib.request.Request(addr, data=data, method=method.upper())
    if data is not None:
        req.add_header("Content-Type", "application/octet-stream")
    tok = cluster_token(token)
    if tok is not None:
        req.add_header(TOKEN_HEADER, tok)
    try:
        with urllib.request.urlopen(req, timeout=timeout) as resp:
            return resp.status, resp.read()
    except urllib.error.HTTPError as e:
        return e.code, e.read()


def call_node(req_type: str, addr: str, content: Any, max_n_requests: int = 100, retry_wait: float = 2.0,
              timeout: float = 100.0, verb: bool = False, token: Optional[str] = None) -> Tuple[int, Any]:
    """POST/PUT ``pickle.dumps(content)``; returns ``(status, decoded body | text)``.

    Retries only while the node cannot be *reached* (not up yet — the reference's start-up race,
    model_dist.py:548-569).  Once a node answers, its status is final: a 4xx/5xx is surfaced with its
    body instead of re-posting a multi-GB chunk a hundred times."""
    method = req_type.lower()
    if method not in ("post", "put"):
        raise ValueError(f"Unsupported request type '{req_type}'")
    payload = pickle.dumps(content)
    status, body = 0, b""
    for attempt in range(max(1, max_n_requests)):
        try:
            status, body = _http(method, addr, payload, timeout, token)
            break
        except (urllib.error.URLError, ConnectionError, TimeoutError, OSError) as e:
            status, body = 0, repr(e).encode()
        if verb:
            print(f"Unable to reach node ({addr}) - retrying in {retry_wait}s ({attempt + 1}/{max_n_requests})")
        if attempt + 1 < max(1, max_n_requests):
            time.sleep(retry_wait)
    if status == 413:
        raise ConnectionError(f"Max payload for {req_type} was exceeded!")
    if status == 200 and body[:1] == b"\x80":
        from ..utils.safe_pickle import safe_loads

        return status, safe_loads(body)
    return status, body.decode("utf-8", "replace") if isinstance(body, bytes) else body


def request_to_node(req_type: str, addr: str, content: Any, max_n_requests: int = 100,
                    retry_wait: float = 2.0, timeout: float = 100.0, verb: bool = False) -> int:
    """POST/PUT ``pickle.dumps(content)`` to ``addr``; 1 if the node answered 200 else 0."""
    status, body = call_node(req_type, addr, content, max_n_requests=max_n_requests, retry_wait=retry_wait,
                             timeout=timeout, verb=verb)
    if status != 200 and verb:
        print(f"Node {addr} answered {status}: {body}")
    return 1 if status == 200 else 0


def http_get_json(addr: str, timeout: float = 10.0) -> Dict[str, Any]:
    req = urllib.request.Request(addr)
    tok = cluster_token()
    if tok is not None:
        req.add_header(TOKEN_HEADER, tok)
    with urllib.request.urlopen(req, timeout=timeout) as resp:
        return json.loads(resp.read().decode("utf-8"))
