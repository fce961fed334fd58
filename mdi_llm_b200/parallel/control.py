"""Control plane: the per-node HTTP endpoint and its client.

Parity: the reference mounts every ``GPTServer`` on a CherryPy ``MethodDispatcher``
(``src/sub/gptserver.py:328-350``) exposing ``GET /`` (node config as JSON), ``POST /init``
(pickled init message, optionally carrying the model chunk), ``PUT /stop``, ``DELETE`` (501)
(``:1114-1226``), and talks to it with ``requests`` + retries (``model_dist.py:499-573``).

CherryPy is not available on the target image, so the same verbs/paths/bodies are served by
the standard library (``http.server.ThreadingHTTPServer``), with no body-size limit (model
chunks are GBs, gptserver.py:345).  Bodies stay pickled dicts for protocol compatibility — the
control plane is meant for a trusted cluster network, exactly like the reference.
"""
from __future__ import annotations

import json
import pickle
import threading
import time
import urllib.error
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Callable, Dict, Optional, Tuple

__all__ = ["ControlServer", "HTTPError", "request_to_node", "http_get_json"]


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

    def __init__(self, app: Any, host: str, port: int) -> None:
        self.app = app
        outer = self

        class _Req(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, fmt: str, *args: Any) -> None:  # silence default stderr log
                pass

            def _dispatch(self, verb: str) -> None:
                length = int(self.headers.get("Content-Length") or 0)
                body = self.rfile.read(length) if length else b""
                path = tuple(p for p in self.path.split("?")[0].split("/") if p)
                status, payload, ctype = 200, b"", "text/plain"
                try:
                    fn = getattr(outer.app, verb, None)
                    if fn is None:
                        raise HTTPError(501, f"{verb} not implemented!")
                    out = fn(path, body)
                    if isinstance(out, str):
                        payload, ctype = out.encode("utf-8"), "application/json"
                    elif isinstance(out, bytes):
                        payload, ctype = out, "application/octet-stream"
                except HTTPError as e:
                    status, payload = e.status, e.message.encode("utf-8")
                except Exception as e:  # noqa: BLE001
                    status, payload = 500, repr(e).encode("utf-8")
                self.send_response(status)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(payload)))
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


def _http(method: str, addr: str, data: Optional[bytes], timeout: float) -> int:
    req = urllib.request.Request(addr, data=data, method=method.upper())
    if data is not None:
        req.add_header("Content-Type", "application/octet-stream")
    try:
        with urllib.request.urlopen(req, timeout=timeout) as resp:
            resp.read()
            return resp.status
    except urllib.error.HTTPError as e:
        return e.code


def request_to_node(req_type: str, addr: str, content: Any, max_n_requests: int = 100,
                    retry_wait: float = 2.0, timeout: float = 100.0, verb: bool = False) -> int:
    """POST/PUT ``pickle.dumps(content)`` to ``addr`` until it answers 200; 1 on success else 0."""
    method = req_type.lower()
    if method not in ("post", "put"):
        raise ValueError(f"Unsupported request type '{req_type}'")
    payload = pickle.dumps(content)
    for attempt in range(max(1, max_n_requests)):
        try:
            status = _http(method, addr, payload, timeout)
            if status == 413:
                raise ConnectionError(f"Max payload for {req_type} was exceeded!")
            if status == 200:
                return 1
        except (urllib.error.URLError, ConnectionError, TimeoutError, OSError):
            status = None
        if verb:
            print(f"Unable to reach node ({addr}) - retrying in {retry_wait}s ({attempt + 1}/{max_n_requests})")
        time.sleep(retry_wait)
    return 0


def http_get_json(addr: str, timeout: float = 10.0) -> Dict[str, Any]:
    with urllib.request.urlopen(addr, timeout=timeout) as resp:
        return json.loads(resp.read().decode("utf-8"))
