"""L0 runtime: SPMD world discovery (rank / size / local device) and the comm object.

The reference captures ``MPI.COMM_WORLD`` rank/size at import time
(``/root/reference/mpi_comms.py:11-13``) and again per optimizer
(``/root/reference/ps.py:71-73``) and is launched with ``mpirun -n N``
(``/root/reference/Makefile:2``).  Here the world is one process per GPU started by
``torchrun`` / :mod:`pytorch_ps_mpi_b200.launch`; rank and size come from the standard
``RANK`` / ``WORLD_SIZE`` / ``LOCAL_RANK`` environment (or an explicit :func:`init`), and
``torch.distributed`` (NCCL on GPUs, gloo on CPU) is used for *bootstrap plumbing only* —
handle exchange, barriers, object slow path.  The hot paths never touch it.
"""
from __future__ import annotations

import datetime
import os
import threading
import uuid
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["World", "init", "world", "rank", "size", "is_initialized", "shutdown", "COMM_WORLD", "bind_to_gpu_numa_node"]


@dataclass
class World:
    """The process's view of the SPMD job (the ``MPI.COMM_WORLD`` analogue)."""

    rank: int = 0
    size: int = 1
    local_rank: int = 0
    device: torch.device = field(default_factory=lambda: torch.device("cpu"))
    backend: Optional[str] = None          # 'nccl' | 'gloo' | None (single process)
    job_id: str = ""
    owns_pg: bool = False                  # we created the default process group
    _cpu_group: Optional[object] = None    # gloo group for host objects when the default is NCCL

    # mpi4py-flavoured accessors so reference-style code keeps working
    def Get_rank(self) -> int:
        return self.rank

    def Get_size(self) -> int:
        return self.size

    # -- bootstrap helpers (slow path; never on a hot path) -----------------------------
    @property
    def cpu_group(self):
        """A gloo group for host-side object exchange (handles, sizes, pickled objects)."""
        if self.size == 1:
            return None
        if self.backend == "gloo":
            return dist.group.WORLD
        if self._cpu_group is None:
            self._cpu_group = dist.new_group(backend="gloo")
        return self._cpu_group

    def barrier(self) -> None:
        if self.size > 1:
            dist.barrier(group=self.cpu_group)

    def all_gather_object(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.cpu_group)
        return out

    def broadcast_object(self, obj, src: int = 0):
        if self.size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.cpu_group)
        return box[0]


_WORLD: Optional[World] = None
_LOCK = threading.Lock()


def _env_int(name: str, default: int) -> int:
    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


def init(backend: Optional[str] = None, device: Optional[torch.device] = None,
         timeout_s: float = 600.0) -> World:
    """Discover (and if needed create) the world.  Idempotent.

    * If ``torch.distributed`` is already initialised, adopt it.
    * Else if ``WORLD_SIZE`` > 1 is in the environment, create the default process group
      (``nccl`` when CUDA is available, else ``gloo``) using the env rendezvous.
    * Else run single-process (size 1) with no process group at all.
    """
    global _WORLD
    with _LOCK:
        if _WORLD is not None:
            return _WORLD
        env_size = _env_int("WORLD_SIZE", 1)
        local_rank = _env_int("LOCAL_RANK", _env_int("RANK", 0))
        use_cuda = torch.cuda.is_available() and (device is None or device.type == "cuda")
        if device is None:
            if use_cuda:
                device = torch.device("cuda", local_rank % max(torch.cuda.device_count(), 1))
            else:
                device = torch.device("cpu")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        owns = False
        if dist.is_available() and dist.is_initialized():
            be = dist.get_backend()
            w = World(dist.get_rank(), dist.get_world_size(), local_rank, device, be)
        elif env_size > 1:
            be = backend or os.environ.get("PSB200_PG_BACKEND") or ("nccl" if device.type == "cuda" else "gloo")
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            kw = {}
            if be == "nccl":
                kw["device_id"] = device
            dist.init_process_group(backend=be, timeout=datetime.timedelta(seconds=timeout_s), **kw)
            owns = True
            w = World(dist.get_rank(), dist.get_world_size(), local_rank, device, be)
        else:
            w = World(0, 1, 0, device, None)
        w.owns_pg = owns
        # a job-unique id every rank agrees on (names shm segments / unix sockets)
        jid = uuid.uuid4().hex[:12] if w.rank == 0 else None
        w.job_id = w.broadcast_object(jid, src=0) if w.size > 1 else jid
        _WORLD = w
        return w


_ALLOC_TUNED = False


def tune_host_allocator() -> bool:
    """Keep large host buffers mapped between messages (host engine / host transports only).

    glibc hands every block above its mmap threshold (at most 32 MB) back to the kernel on ``free`` and maps fresh,
    zero pages for the next one, so each multi-megabyte message pays a page fault per 4 KB on the sender (frame buffer)
    and on the receiver — 277 ms instead of 8 ms for a 64 MB copy in this container.  Raising ``M_MMAP_THRESHOLD`` /
    ``M_TRIM_THRESHOLD`` makes those buffers recycle through the heap instead.  ``PSB200_MALLOC_TUNE=0`` disables it."""
    global _ALLOC_TUNED
    if _ALLOC_TUNED:
        return True
    if os.environ.get("PSB200_MALLOC_TUNE", "1") == "0":
        return False
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_MMAP_THRESHOLD = -1, -3
        ok = libc.mallopt(M_MMAP_THRESHOLD, 1 << 30) == 1 and libc.mallopt(M_TRIM_THRESHOLD, 2 ** 31 - 1) == 1
    except Exception:          # not glibc
        ok = False
    _ALLOC_TUNED = ok
    return ok


def bind_to_gpu_numa_node(device=None, min_cpus: int = 8, nvml=None) -> dict:
    """Pin every thread of this process to the CPUs NVML reports as local to ``device``'s PCIe root (the socket / NUMA node
    the GPU hangs off), so the launch path rings a local doorbell and pinned staging buffers allocated afterwards are
    first-touched on the local node — what ``mpirun --bind-to numa`` does for the reference's ranks (``Makefile:2`` leaves
    it to the MPI launcher).  One process per GPU on a two-socket box otherwise runs wherever the scheduler put it and
    half of the ranks pull their input batches across the socket interconnect.

    Never raises; returns ``{"bound": bool, "cpus": n, "why": ...}``.  Skipped when the allowed set would shrink below
    ``min_cpus`` or a quarter of what the process may use now, or with ``PSB200_NUMA_BIND=0``."""
    out = {"bound": False, "cpus": 0, "why": ""}
    if os.environ.get("PSB200_NUMA_BIND", "1") == "0":
        out["why"] = "disabled"
        return out
    try:
        allowed = os.sched_getaffinity(0)
        if nvml is None:
            import pynvml as nvml
            nvml.nvmlInit()
        if device is None:
            device = world().device
        idx = device if isinstance(device, int) else (device.index or 0)
        handle = None
        if not isinstance(device, int) and torch.cuda.is_available():
            pr = torch.cuda.get_device_properties(idx)
            if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                try:                          # CUDA_VISIBLE_DEVICES renumbers CUDA devices, not NVML's
                    handle = nvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
                except Exception:             # noqa: BLE001
                    handle = None
        if handle is None:
            handle = nvml.nvmlDeviceGetHandleByIndex(idx)
        words = nvml.nvmlDeviceGetCpuAffinity(handle, (max(allowed | {os.cpu_count() or 1}) + 64) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if int(word) >> b & 1}
        want = allowed & local
        out["cpus"] = len(want)
        # never trade locality for contention: the local share of the allowed CPUs must be a real socket's worth (a cpuset
        # that leaves only a handful of local CPUs would stack every rank of that socket onto them)
        if len(want) < max(min_cpus, len(allowed) // 4):
            out["why"] = f"only {len(want)} of the {len(allowed)} allowed CPUs are local to the GPU"
            return out
        if want == allowed:
            out.update(bound=True, why="already local")
            return out
        tids = [int(t) for t in os.listdir("/proc/self/task")] if os.path.isdir("/proc/self/task") else [0]
        for tid in tids:                      # threads that already exist (CUDA, gloo, NVML sampler) keep their own mask
            try:
                os.sched_setaffinity(tid, want)
            except OSError:
                pass
        os.sched_setaffinity(0, want)         # and threads created from here on inherit this one
        out.update(bound=True, why="ok")
    except Exception as exc:      # noqa: BLE001 - an optimisation only: never take the job down
        out["why"] = f"{type(exc).__name__}: {exc}"[:120]
    return out


def world() -> World:
    return _WORLD if _WORLD is not None else init()


def is_initialized() -> bool:
    return _WORLD is not None


def rank() -> int:
    return world().rank


def size() -> int:
    return world().size


def shutdown() -> None:
    """Tear down the world (destroys the process group only if :func:`init` created it)."""
    global _WORLD
    with _LOCK:
        w, _WORLD = _WORLD, None
    if w is not None and w.owns_pg and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass


class _CommWorldProxy:
    """``COMM_WORLD`` stand-in: resolves lazily so importing the package never blocks."""

    def Get_rank(self) -> int:
        return world().rank

    def Get_size(self) -> int:
        return world().size

    def Barrier(self) -> None:
        world().barrier()

    def __getattr__(self, item):
        return getattr(world(), item)


COMM_WORLD = _CommWorldProxy()
