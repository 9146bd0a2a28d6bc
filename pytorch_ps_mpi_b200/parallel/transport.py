"""L1 host transports: non-blocking byte messaging between ranks for the generic-object path.

The reference's L1 is ``mpi4py`` non-blocking collectives on host byte buffers
(``/root/reference/mpi_comms.py:88,132,153,162``) plus, in the async spec, point-to-point
``recv(MPI.ANY_SOURCE)`` / ``send`` (``/root/reference/README.md:65-76``).  Here one small
interface — ``isend`` / ``irecv`` (with ``ANY_SOURCE``) / ``barrier`` returning
:class:`Request` objects with the MPI-style ``Wait()`` / ``Test()`` — has three backends:

* :class:`ShmTransport`  — the native single-node transport (``csrc/runtime/host_ext.cpp``),
* :class:`GlooTransport` — ``torch.distributed`` p2p (multi-node capable fallback),
* :class:`LocalTransport` — world size 1 loop-back.

Collectives (gather / bcast / all-gather) are composed from these in
:mod:`pytorch_ps_mpi_b200.mpi_comms`; the GPU hot paths do not use any of this.
"""
from __future__ import annotations

import collections
import os
import time
from typing import Deque, Dict, List, Optional, Tuple

import torch

from .. import runtime

ANY_SOURCE = -1
__all__ = ["ANY_SOURCE", "Request", "Transport", "LocalTransport", "ShmTransport",
           "GlooTransport", "get_transport", "reset_transport"]


class Request:
    """MPI-request look-alike.  ``Wait()`` returns the received buffer for receives."""

    def Wait(self, timeout: Optional[float] = None):   # noqa: N802 (MPI naming)
        raise NotImplementedError

    def Test(self) -> bool:   # noqa: N802
        raise NotImplementedError

    wait = Wait
    test = Test


class _Done(Request):
    def __init__(self, value=None, src: int = -1):
        self.value, self.source = value, src

    def Wait(self, timeout=None):
        return self.value

    def Test(self):
        return True


class Transport:
    rank: int = 0
    size: int = 1
    @property
    def default_timeout(self) -> float:
        """Seconds a blocking wait may take before it raises (``PSB200_COMM_TIMEOUT``, read at use)."""
        return float(os.environ.get("PSB200_COMM_TIMEOUT", "300"))

    def isend(self, dst: int, data, tag: int = 0) -> Request:
        raise NotImplementedError

    def irecv(self, src: int = ANY_SOURCE, tag: int = 0) -> Request:
        """``Wait()`` → buffer-protocol object; the request gains ``.source`` once complete."""
        raise NotImplementedError

    def barrier(self) -> None:
        raise NotImplementedError

    def close(self) -> None:
        pass


# ---------------------------------------------------------------------------------------
class LocalTransport(Transport):
    """World of one: sends to self are queued and matched by tag."""

    def __init__(self):
        self.rank, self.size = 0, 1
        self._q: Dict[int, Deque[bytearray]] = collections.defaultdict(collections.deque)

    def isend(self, dst, data, tag=0):
        assert dst == 0
        self._q[tag].append(bytearray(data))
        return _Done()

    def irecv(self, src=ANY_SOURCE, tag=0):
        q = self._q[tag]
        if not q:
            raise RuntimeError("LocalTransport.irecv: nothing was sent (would deadlock)")
        return _Done(q.popleft(), 0)

    def barrier(self):
        pass


# ---------------------------------------------------------------------------------------
class _ShmRequest(Request):
    def __init__(self, comm, op, timeout):
        self._comm, self._op, self._timeout = comm, op, timeout
        self.source = -1

    def Wait(self, timeout=None):
        self._comm.wait(self._op, self._timeout if timeout is None else timeout)
        return self._finish()

    def Test(self):
        return bool(self._comm.test(self._op))

    def _finish(self):
        if self._op.is_send:
            return None
        msg = self._op.message
        self.source = msg.src
        return msg


class ShmTransport(Transport):
    """Native POSIX-shm transport: one SPSC ring per ordered rank pair, MPI-style progress."""

    def __init__(self, rank: int, size: int, job_id: str, ring_bytes: Optional[int] = None):
        from ..ops import ext
        h = ext.host()
        ring = int(ring_bytes or os.environ.get("PSB200_SHM_RING_BYTES", 8 << 20))
        self.rank, self.size = rank, size
        self._comm = h.ShmComm(f"/psb200_{job_id}", rank, size, ring, 120.0)

    def isend(self, dst, data, tag=0):
        if not isinstance(data, (bytes, bytearray, memoryview)):
            data = memoryview(data)
        return _ShmRequest(self._comm, self._comm.post_send(dst, tag, data), self.default_timeout)

    def irecv(self, src=ANY_SOURCE, tag=0):
        return _ShmRequest(self._comm, self._comm.post_recv(src, tag), self.default_timeout)

    def barrier(self):
        self._comm.barrier(self.default_timeout)

    def dead_peers(self) -> List[int]:
        return list(self._comm.dead_peers())

    def abort(self):
        self._comm.abort()


# ---------------------------------------------------------------------------------------
class _GlooSend(Request):
    def __init__(self, works, keep):
        self._works, self._keep = works, keep

    def Wait(self, timeout=None):
        for w in self._works:
            w.wait()
        self._keep = None

    def Test(self):
        return all(w.is_completed() for w in self._works)


class _GlooRecv(Request):
    """Two-phase receive: 8-byte length (possibly from ANY_SOURCE) then the payload."""

    def __init__(self, tr: "GlooTransport", src: int, tag: int):
        import torch.distributed as dist
        self._tr, self._tag = tr, tag
        self._len = torch.zeros(1, dtype=torch.int64)
        self._w = dist.irecv(self._len, src=None if src == ANY_SOURCE else src, group=tr.group, tag=tag)
        self._src = src
        self.source = -1
        self._payload = None
        self._pw = None

    def _phase2(self):
        import torch.distributed as dist
        if self._pw is None and self._payload is None:
            self._w.wait()
            self.source = self._src if self._src != ANY_SOURCE else self._w._source_rank()
            n = int(self._len.item())
            self._payload = torch.empty(max(n, 1), dtype=torch.uint8)
            self._n = n
            if n:
                self._pw = dist.irecv(self._payload, src=self.source, group=self._tr.group, tag=self._tag)

    def Wait(self, timeout=None):
        self._phase2()
        if self._pw is not None:
            self._pw.wait()
            self._pw = None
        return memoryview(self._payload.numpy())[: self._n]

    def Test(self):
        if not self._w.is_completed():
            return False
        self._phase2()
        return self._pw is None or self._pw.is_completed()


class GlooTransport(Transport):
    """``torch.distributed`` point-to-point over the host (gloo) group."""

    def __init__(self, world: runtime.World):
        self.rank, self.size = world.rank, world.size
        self.group = world.cpu_group
        self._world = world

    def isend(self, dst, data, tag=0):
        import torch.distributed as dist
        mv = memoryview(data).cast("B")
        n = len(mv)
        ln = torch.tensor([n], dtype=torch.int64)
        works = [dist.isend(ln, dst=dst, group=self.group, tag=tag)]
        keep = [ln]
        if n:
            if mv.readonly:
                payload = torch.frombuffer(bytearray(mv), dtype=torch.uint8)
            else:
                payload = torch.frombuffer(mv, dtype=torch.uint8)
            works.append(dist.isend(payload, dst=dst, group=self.group, tag=tag))
            keep.append(payload)
        return _GlooSend(works, keep)

    def irecv(self, src=ANY_SOURCE, tag=0):
        return _GlooRecv(self, src, tag)

    def barrier(self):
        self._world.barrier()


# ---------------------------------------------------------------------------------------
_TRANSPORT: Optional[Transport] = None


def get_transport(kind: Optional[str] = None) -> Transport:
    """The process-wide host transport.

    ``kind`` / ``PSB200_TRANSPORT``: ``shm`` (default on one node), ``gloo``, ``local``.
    """
    global _TRANSPORT
    if _TRANSPORT is not None:
        return _TRANSPORT
    w = runtime.world()
    kind = (kind or os.environ.get("PSB200_TRANSPORT") or "").lower()
    runtime.tune_host_allocator()          # message-sized host buffers recycle instead of page-faulting every step
    if w.size == 1:
        _TRANSPORT = LocalTransport()
        return _TRANSPORT
    if not kind:
        single_node = int(os.environ.get("LOCAL_WORLD_SIZE", w.size)) == w.size
        kind = "shm" if single_node else "gloo"
    if kind == "shm":
        try:
            _TRANSPORT = ShmTransport(w.rank, w.size, w.job_id)
            ok = True
        except Exception as e:   # pragma: no cover - compiler missing etc.
            ok, err = False, e
        oks = w.all_gather_object(ok)
        if not all(oks):
            _TRANSPORT = None
            if os.environ.get("PSB200_TRANSPORT", "").lower() == "shm":
                raise RuntimeError(f"shm transport requested but unavailable on some rank: {oks}")
            kind = "gloo"
    if kind == "gloo":
        _TRANSPORT = GlooTransport(w)
    if _TRANSPORT is None:
        raise ValueError(f"unknown transport {kind!r}")
    return _TRANSPORT


def reset_transport() -> None:
    global _TRANSPORT
    t, _TRANSPORT = _TRANSPORT, None
    if t is not None:
        t.close()
