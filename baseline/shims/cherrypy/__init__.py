"""Minimal stand-in for the parts of CherryPy the reference uses (not installable offline).

Covers exactly: ``cp.dispatch.MethodDispatcher()``, ``cp.tree.mount(app, "/", conf)``,
``cp.config.update({...socket_host/socket_port...})``, ``cp.engine.start/stop/exit/block``,
``cp.request.body.read()``, ``cp.response.status`` and ``cp.HTTPError``.  Requests are dispatched
to ``app.GET/POST/PUT/DELETE(*path_segments)`` like CherryPy's MethodDispatcher does.  This is
environment plumbing for running the UNMODIFIED reference; no reference logic lives here.
"""
from __future__ import annotations

import io
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer


class HTTPError(Exception):
    def __init__(self, status=500, message=""):
        super().__init__(f"{status} {message}")
        self.status, self.message = int(status), str(message)


class _Local(threading.local):
    pass


class _Request(_Local):
    body = io.BytesIO(b"")


class _Response(_Local):
    status = 200


request = _Request()
response = _Response()


class _Dispatch:
    class MethodDispatcher:  # marker only
        pass


dispatch = _Dispatch()


class _Config(dict):
    def update(self, other=None, **kw):  # noqa: A003
        super().update(other or {}, **kw)


config = _Config()


class _Tree:
    def __init__(self):
        self.app = None

    def mount(self, app, script_name="/", config=None):  # noqa: A002
        self.app = app


tree = _Tree()


class _Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"

    def log_message(self, *a):
        pass

    def _go(self, verb):
        n = int(self.headers.get("Content-Length") or 0)
        request.body = io.BytesIO(self.rfile.read(n) if n else b"")
        response.status = 200
        path = [p for p in self.path.split("?")[0].split("/") if p]
        status, payload = 200, b""
        try:
            fn = getattr(tree.app, verb, None)
            if fn is None:
                raise HTTPError(405, "method not allowed")
            out = fn(*path)
            status = int(response.status or 200)
            if isinstance(out, str):
                payload = out.encode()
            elif isinstance(out, bytes):
                payload = out
        except HTTPError as e:
            status, payload = e.status, e.message.encode()
        except Exception as e:  # noqa: BLE001
            status, payload = 500, repr(e).encode()
        self.send_response(status)
        self.send_header("Content-Length", str(len(payload)))
        self.end_headers()
        if payload:
            self.wfile.write(payload)

    def do_GET(self):  # noqa: N802
        self._go("GET")

    def do_POST(self):  # noqa: N802
        self._go("POST")

    def do_PUT(self):  # noqa: N802
        self._go("PUT")

    def do_DELETE(self):  # noqa: N802
        self._go("DELETE")


class _Engine:
    def __init__(self):
        self.httpd = None
        self.thread = None
        self._stop = threading.Event()

    def start(self):
        host = config.get("server.socket_host", "127.0.0.1")
        port = int(config.get("server.socket_port", 8080))
        ThreadingHTTPServer.allow_reuse_address = True
        self.httpd = ThreadingHTTPServer((host, port), _Handler)
        self.httpd.daemon_threads = True
        self._stop.clear()
        self.thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self.thread.start()

    def stop(self):
        if self.httpd is not None:
            httpd, self.httpd = self.httpd, None
            threading.Thread(target=lambda: (httpd.shutdown(), httpd.server_close()), daemon=True).start()

    def exit(self):  # noqa: A003
        self._stop.set()

    def block(self):
        while not self._stop.wait(0.5):
            pass


engine = _Engine()
