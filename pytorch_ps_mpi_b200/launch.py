"""SPMD launcher: the ``mpirun -n N`` stand-in (``/root/reference/Makefile:2``).

``python -m pytorch_ps_mpi_b200.launch -n 2 script.py args…`` starts N copies of a script with
``RANK`` / ``WORLD_SIZE`` / ``LOCAL_RANK`` / ``MASTER_ADDR=127.0.0.1`` set (torchrun-compatible),
one per GPU when GPUs are visible.  :func:`spawn` does the same for a Python callable and is
what the multi-process tests use.
"""
from __future__ import annotations

import argparse
import os
import socket
import subprocess
import sys
import traceback
from typing import Callable, Optional, Sequence

import torch.multiprocessing as mp


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _env_for(rank: int, size: int, port: int) -> dict:
    return {"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(size),
            "LOCAL_WORLD_SIZE": str(size), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)}


def _entry(rank: int, fn: Callable, size: int, port: int, args: tuple, env: dict, errq):
    os.environ.update(_env_for(rank, size, port))
    os.environ.update(env)
    try:
        import torch
        torch.set_num_threads(int(os.environ.get("PSB200_SPAWN_THREADS", "1")))   # N ranks share the cores
        fn(rank, size, *args)
    except BaseException:
        errq.put((rank, traceback.format_exc()))
        raise
    finally:
        try:
            from . import runtime
            runtime.shutdown()
        except Exception:
            pass


def spawn(fn: Callable, nprocs: int, args: Sequence = (), env: Optional[dict] = None,
          timeout: float = 300.0) -> None:
    """Run ``fn(rank, size, *args)`` in ``nprocs`` fresh processes; raise if any rank fails."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = free_port()
    procs = []
    for r in range(nprocs):
        p = ctx.Process(target=_entry, args=(r, fn, nprocs, port, tuple(args), dict(env or {}), errq))
        p.start()
        procs.append(p)
    # ONE deadline for the whole group, and fail fast: as soon as any rank exits non-zero the others (which would sit in a
    # barrier / device spin until their own time-outs) are terminated — a single failing rank used to cost nprocs x timeout
    import time as _time
    failed = []
    deadline = _time.time() + timeout
    while True:
        alive = [p for p in procs if p.is_alive()]
        if not alive:
            break
        if any(p.exitcode not in (None, 0) for p in procs):
            _time.sleep(2.0)                      # let the failing rank's traceback reach the queue
            break
        if _time.time() > deadline:
            failed.append(f"{len(alive)} rank(s) still running after {timeout}s")
            break
        _time.sleep(0.05)
    for p in procs:
        if p.is_alive():
            p.terminate()          # exact processes we started
    for p in procs:
        p.join(10)
        if p.is_alive():
            p.kill()
            p.join(5)
    msgs = []
    while not errq.empty():
        r, tb = errq.get()
        msgs.append(f"--- rank {r} ---\n{tb}")
    bad = [p.exitcode for p in procs if p.exitcode not in (0,)]
    if failed or bad or msgs:
        raise RuntimeError("spawned ranks failed: " + "; ".join(failed) + f" exitcodes={[p.exitcode for p in procs]}\n"
                           + "\n".join(msgs))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="SPMD launcher (mpirun -n N stand-in)")
    ap.add_argument("-n", "--nproc", type=int, default=1)
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    port = a.port or free_port()
    procs = []
    for r in range(a.nproc):
        env = dict(os.environ)
        env.update(_env_for(r, a.nproc, port))
        procs.append(subprocess.Popen([sys.executable, a.script, *a.script_args], env=env))
    rc = 0
    for p in procs:
        rc = p.wait() or rc
    return rc


if __name__ == "__main__":
    sys.exit(main())
