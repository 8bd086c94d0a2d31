"""Engine daemon + client shim (hard part D).

A corpus shard in HBM belongs to ONE process per host, but the reference reaches its vector store
from 2 gunicorn workers x 4 threads, 4 Celery children and the chatbot process
(docker-compose.yaml:191, :283-285) -- over gRPC to the Weaviate container.  Here the owner is this
daemon and the other processes talk to it over a Unix-domain socket with the SAME module API:

    server:  python -m aurora_b200.daemon --socket /run/aurora_b200.sock     (after configure())
    client:  from aurora_b200.daemon import Client; kb = Client("/run/aurora_b200.sock")
             kb.search_knowledge_base(user_id, query, limit=5)              # weaviate_client.py:215

Wire format: 4-byte big-endian length + JSON ``{"fn", "args", "kwargs"}`` -> ``{"ok", "result" | "error"}``.
The client keeps the reference's error conventions when the daemon is unreachable (search -> [],
deletes -> -1, counts -> 0, insert re-raises so the Celery task retries; weaviate_client.py:210-212,
:283-285, :317-319, :369-371).  ``health()`` replaces the Weaviate readiness probe of
routes/health_routes.py:76-91.
"""

from __future__ import annotations

import json
import logging
import os
import socket
import socketserver
import struct
import threading
from typing import Any, Optional

logger = logging.getLogger(__name__)

API = ("insert_chunks", "search_knowledge_base", "delete_document_chunks", "delete_user_chunks",
       "get_document_chunk_count", "delete_discovery_chunks")


def _send(sock: socket.socket, obj: Any) -> None:
    raw = json.dumps(obj).encode("utf-8")
    sock.sendall(struct.pack(">I", len(raw)) + raw)


def _recv(sock: socket.socket) -> Optional[Any]:
    def read(n: int) -> Optional[bytes]:
        buf = b""
        while len(buf) < n:
            chunk = sock.recv(n - len(buf))
            if not chunk:
                return None
            buf += chunk
        return buf

    head = read(4)
    if head is None:
        return None
    body = read(struct.unpack(">I", head)[0])
    return None if body is None else json.loads(body.decode("utf-8"))


class _Handler(socketserver.BaseRequestHandler):
    def handle(self) -> None:
        module = self.server.module          # type: ignore[attr-defined]
        while True:
            try:
                req = _recv(self.request)
            except (OSError, ValueError):
                return
            if req is None:
                return
            fn = req.get("fn")
            try:
                if fn == "health":
                    res = {"ready": True, "pid": os.getpid()}
                elif fn in API:
                    res = getattr(module, fn)(*req.get("args", []), **req.get("kwargs", {}))
                else:
                    raise ValueError(f"unknown function {fn!r}")
                _send(self.request, {"ok": True, "result": res})
            except Exception as e:           # insert_chunks re-raises by design: report it to the caller
                _send(self.request, {"ok": False, "error": f"{type(e).__name__}: {e}"})


class _Server(socketserver.ThreadingMixIn, socketserver.UnixStreamServer):
    daemon_threads = True
    allow_reuse_address = True
    request_queue_size = 256      # gunicorn threads + Celery children connect at once


def serve(socket_path: str, module=None, background: bool = False):
    """Serve ``module`` (default: aurora_b200.retriever, already ``configure()``d) on ``socket_path``."""
    if module is None:
        from . import retriever as module
    if os.path.exists(socket_path):
        os.unlink(socket_path)
    srv = _Server(socket_path, _Handler)
    srv.module = module                     # type: ignore[attr-defined]
    if background:
        t = threading.Thread(target=srv.serve_forever, name="aurora-b200-daemon", daemon=True)
        t.start()
        return srv
    try:
        srv.serve_forever()
    finally:
        srv.server_close()
    return srv


class DaemonUnavailable(RuntimeError):
    pass


class Client:
    """Same names / signatures as routes.knowledge_base.weaviate_client; one connection per thread."""

    def __init__(self, socket_path: str, timeout: float = 30.0):
        self._path, self._timeout = socket_path, timeout
        self._local = threading.local()

    def _call(self, fn: str, *args, **kwargs):
        for attempt in (0, 1):               # one reconnect: the daemon may have restarted
            sock = getattr(self._local, "sock", None)
            try:
                if sock is None:
                    sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                    sock.settimeout(self._timeout)
                    sock.connect(self._path)
                    self._local.sock = sock
                _send(sock, {"fn": fn, "args": list(args), "kwargs": kwargs})
                resp = _recv(sock)
                if resp is None:
                    raise ConnectionError("daemon closed the connection")
                if not resp["ok"]:
                    raise RuntimeError(resp["error"])
                return resp["result"]
            except (OSError, ConnectionError) as e:
                if sock is not None:
                    try:
                        sock.close()
                    except OSError:
                        pass
                self._local.sock = None
                if attempt == 1:
                    raise DaemonUnavailable(str(e)) from e

    def health(self) -> dict:
        try:
            return self._call("health")
        except Exception as e:
            return {"ready": False, "error": str(e)}

    # ---- the reference's module API, with its error conventions on transport failure
    def insert_chunks(self, user_id, document_id, source_filename, chunks, org_id=None) -> int:
        if not chunks:
            return 0
        return self._call("insert_chunks", user_id, document_id, source_filename, chunks, org_id=org_id)   # raises -> retry

    def search_knowledge_base(self, user_id, query, limit=5, alpha=0.5, min_score=0.0, org_id=None):
        if not query.strip():
            return []
        try:
            return self._call("search_knowledge_base", user_id, query, limit=limit, alpha=alpha, min_score=min_score, org_id=org_id)
        except Exception as e:
            logger.error(f"[KB B200 client] Error searching: {e}")
            return []

    def delete_document_chunks(self, user_id, document_id) -> int:
        try:
            return self._call("delete_document_chunks", user_id, document_id)
        except Exception as e:
            logger.error(f"[KB B200 client] Error deleting chunks: {e}")
            return -1

    def delete_user_chunks(self, user_id) -> int:
        try:
            return self._call("delete_user_chunks", user_id)
        except Exception as e:
            logger.error(f"[KB B200 client] Error deleting user chunks: {e}")
            return -1

    def get_document_chunk_count(self, user_id, document_id) -> int:
        try:
            return self._call("get_document_chunk_count", user_id, document_id)
        except Exception:
            return 0

    def delete_discovery_chunks(self, org_id, before=None) -> int:
        try:
            return self._call("delete_discovery_chunks", org_id, before=before)
        except Exception:
            return 0


if __name__ == "__main__":       # pragma: no cover
    import argparse

    ap = argparse.ArgumentParser(description="aurora_b200 engine daemon (configure the retriever first via AURORA_B200_BOOT)")
    ap.add_argument("--socket", default=os.getenv("AURORA_B200_SOCKET", "/tmp/aurora_b200.sock"))
    ap.add_argument("--boot", default=os.getenv("AURORA_B200_BOOT"), help="module:function that calls retriever.configure(...)")
    a = ap.parse_args()
    if a.boot:
        mod, fn = a.boot.split(":")
        getattr(__import__(mod, fromlist=[fn]), fn)()
    serve(a.socket)
