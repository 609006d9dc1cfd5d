"""REST host: `GET /` (or POST) with the reference's JSON payload -> the reference's JSON response.

Mirrors `main.py:30-66` of the reference: `check_input`, solver chosen by name from
`neptune_mip_b200.core.solvers`, `processing_time` around `load_data + solve`, response keys
`cpu_routing_rules, cpu_allocations, gpu_routing_rules, gpu_allocations, score, processing_time`.
Differences, on purpose: the solver class is looked up in the solver package instead of `eval`;
the server is single-process and single-threaded (a CUDA context does not survive the reference's
`processes=10` pre-fork, `main.py:69`); `with_db` must be false.  Standard library only.
"""
from __future__ import annotations

import json
import time
from http.server import BaseHTTPRequestHandler, HTTPServer

from .core import check_input, data_to_solver_input
from .core import solvers as _solvers


def solve_payload(payload: dict) -> dict:
    check_input(payload)
    cfg = payload.get("solver", {"type": "NeptuneMinDelayAndUtilization"})
    cls = getattr(_solvers, cfg.get("type"), None)
    if cls is None or not isinstance(cls, type):
        raise KeyError(f"unknown solver type {cfg.get('type')!r}")
    solver = cls(**cfg.get("args", {}))
    with_db = payload.get("with_db", True)
    t0 = time.time()
    solver.load_data(data_to_solver_input(payload, with_db=with_db, workload_coeff=payload.get("workload_coeff", 1)))
    solver.solve()
    dt = time.time() - t0
    x, c = solver.results()
    return {"cpu_routing_rules": x, "cpu_allocations": c, "gpu_routing_rules": {}, "gpu_allocations": {},
            "score": solver.score(), "processing_time": dt}


class Handler(BaseHTTPRequestHandler):
    def _serve(self):
        try:
            n = int(self.headers.get("Content-Length", "0"))
            payload = json.loads(self.rfile.read(n) or b"{}")
            body = json.dumps(solve_payload(payload)).encode()
            self.send_response(200)
        except Exception as e:          # the reference answers 500 with the Flask debugger page
            body = json.dumps({"error": f"{type(e).__name__}: {e}"}).encode()
            self.send_response(500)
        self.send_header("Content-Type", "application/json")
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        self.wfile.write(body)

    do_GET = _serve
    do_POST = _serve

    def log_message(self, fmt, *args):
        print("[neptune_mip_b200] " + fmt % args)


def main(host="0.0.0.0", port=5000):
    HTTPServer((host, port), Handler).serve_forever()


if __name__ == "__main__":
    main()
