"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference from /root/reference.

The reference cannot be imported as shipped: `ortools`, `sqlalchemy`, `hurry.filesize`
and `flask` are absent offline.  This module registers stub
modules for the first three (flask is only needed by `main.py`, which we never import
because it starts a server at import time, `main.py:69`) and then imports `core`.

`serve(payload)` restates `main.py:31-66` without Flask so the reference's response
dict can be produced for golden fixtures.  /root/reference does not exist on the GPU
box: every caller must be guarded by `reference_available()`.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import time
import types

REFERENCE_ROOT = os.environ.get("NEPTUNE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "solvers"))


_loaded = {}


def load_reference():
    """Return the reference's `core` package (imported once)."""
    if "core" in _loaded:
        return _loaded["core"]
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    from . import pywraplp_highs

    sqlalchemy = types.ModuleType("sqlalchemy")
    sqlalchemy.create_engine = lambda *a, **k: (_ for _ in ()).throw(
        RuntimeError("with_db=True needs the in-cluster Postgres; out of scope"))
    hurry = types.ModuleType("hurry")
    hurry_fs = types.ModuleType("hurry.filesize")
    hurry_fs.size = lambda b: f"{b}B"
    hurry.filesize = hurry_fs
    ortools = types.ModuleType("ortools")
    ls = types.ModuleType("ortools.linear_solver")
    ortools.linear_solver = ls
    ls.pywraplp = pywraplp_highs
    sys.modules.setdefault("sqlalchemy", sqlalchemy)
    sys.modules.setdefault("hurry", hurry)
    sys.modules.setdefault("hurry.filesize", hurry_fs)
    sys.modules["ortools"] = ortools
    sys.modules["ortools.linear_solver"] = ls
    sys.modules["ortools.linear_solver.pywraplp"] = pywraplp_highs

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        core = importlib.import_module("core")
        importlib.import_module("core.solvers")
    _loaded["core"] = core
    return core


class _DiscardSet(set):
    """`set` whose remove() does not raise -- see `enable_efttc_discard_patch`."""

    def remove(self, item):
        self.discard(item)


def enable_efttc_discard_patch(enable: bool = True):
    """Rebind the name `set` inside the reference's efttc_step1 module.

    The reference raises KeyError at `efttc_step1.py:118` whenever a function appears
    twice in an accepted cycle.  The oracle policy is: run the
    reference unmodified; where it raises, rerun with `remaining_functions.remove`
    behaving like `discard`.  The reference file is untouched.
    """
    load_reference()
    mod = sys.modules["core.solvers.efttc.efttc_step1"]
    if enable:
        mod.set = _DiscardSet
    elif "set" in mod.__dict__:
        del mod.__dict__["set"]


def serve(payload: dict, quiet: bool = True) -> dict:
    """Restatement of `main.py:31-66` (no Flask): payload dict -> response dict."""
    core = load_reference()
    solvers = sys.modules["core.solvers"]
    sink = io.StringIO() if quiet else sys.stdout
    with contextlib.redirect_stdout(sink):
        core.check_input(payload)
        solver_cfg = payload.get("solver", {"type": "NeptuneMinDelayAndUtilization"})
        solver = getattr(solvers, solver_cfg.get("type"))(**solver_cfg.get("args", {}))
        t0 = time.time()
        data = core.data_to_solver_input(payload, with_db=payload.get("with_db", True),
                                         workload_coeff=payload.get("workload_coeff", 1))
        solver.load_data(data)
        solved = solver.solve()
        dt = time.time() - t0
        x, c = solver.results()
        score = solver.score()
    return {"cpu_routing_rules": x, "cpu_allocations": c, "gpu_routing_rules": {},
            "gpu_allocations": {}, "score": score, "processing_time": dt,
            "_solved": bool(solved), "_solver": solver, "_data": data}
