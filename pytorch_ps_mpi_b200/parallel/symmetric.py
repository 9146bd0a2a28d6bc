"""Symmetric + multicast device memory for the PS kernels.

One :class:`SymmetricArena` = one physical block per GPU, the same size on every rank, every
rank's block mapped into every process (peer pointers over NVLink) plus — when the fabric has
NVLS — one multicast alias through which a single store lands on all GPUs and a single load
returns the sum over all GPUs.

Providers
---------
``native`` (default)  our own C++ runtime on the CUDA VMM driver API
    (``csrc/runtime/symm_mem.cpp``: ``cuMemCreate`` → POSIX fd over ``SCM_RIGHTS`` →
    ``cuMemImportFromShareableHandle`` → ``cuMemMap``; ``cuMulticastCreate/AddDevice/BindMem``).
``torch``  bootstrap fallback through ``torch.distributed._symmetric_memory`` (same layout, same
    kernels) if the native exchange cannot be set up on some rank.

``torch.distributed`` is only used for barriers / agreeing on success — never for data.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from .. import runtime
from ..ops import ext

__all__ = ["SymmetricArena"]

_SERIAL = 0


class SymmetricArena:
    """A symmetric block of ``nbytes`` (rounded up to the VMM granularity) on every rank."""

    def __init__(self, nbytes: int, device: torch.device, world: Optional[runtime.World] = None,
                 provider: Optional[str] = None, multicast: bool = True):
        global _SERIAL
        self.world = world or runtime.world()
        self.rank, self.size = self.world.rank, self.world.size
        self.device = device
        self.requested = int(nbytes)
        self.mc_ptr = 0
        self.provider = None
        self._native = None
        self._torch_buf = None
        self._torch_hdl = None
        _SERIAL += 1
        serial = _SERIAL
        want = (provider or os.environ.get("PSB200_SYMM_PROVIDER") or "native").lower()
        err = None
        if want == "native":
            try:
                self._init_native(serial, multicast)
                ok = True
            except Exception as e:    # noqa: BLE001 - any failure → collective fallback decision
                ok, err = False, e
            oks = self.world.all_gather_object(ok)
            if all(oks):
                self.provider = "native"
            else:
                self._native = None
                if os.environ.get("PSB200_SYMM_PROVIDER", "").lower() == "native":
                    raise RuntimeError(f"native symmetric memory failed on ranks "
                                       f"{[i for i, o in enumerate(oks) if not o]}: {err}")
                if self.rank == 0:
                    print(f"[psb200] native symmetric memory unavailable ({err}); "
                          f"falling back to torch symmetric memory", flush=True)
        if self.provider is None:
            self._init_torch()
            self.provider = "torch"

    # ------------------------------------------------------------------------------------
    def _init_native(self, serial: int, multicast: bool) -> None:
        m = ext.cuda()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        prefix = f"psb200-{self.world.job_id}-{serial}"
        blk = m.SymmBlock(self.rank, self.size, dev_index, self.requested, prefix)
        self._native = blk
        self.world.barrier()                       # every rank is listening
        if self.size > 1:
            blk.map_peers()
        self.ptrs: List[int] = list(blk.ptrs)
        self.nbytes = int(blk.size)
        # ---- multicast (NVLS) ----
        mc_ok = bool(multicast and self.size > 1 and blk.mc_supported())
        mc_ok = all(self.world.all_gather_object(mc_ok))
        if mc_ok:
            step = blk.mc_create() if self.rank == 0 else True
            step = all(self.world.all_gather_object(bool(step)))
            if step:
                step = blk.mc_import() if self.rank != 0 else True
                step = all(self.world.all_gather_object(bool(step)))
            if step:
                step = all(self.world.all_gather_object(bool(blk.mc_add_device())))
            if step:
                step = all(self.world.all_gather_object(bool(blk.mc_bind_and_map())))
            if step:
                self.mc_ptr = int(blk.mc_ptr)
            elif self.rank == 0:
                print(f"[psb200] multicast unavailable: {blk.last_error}", flush=True)
        self.world.barrier()
        blk.stop_server()

    def _init_torch(self) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if self.size == 1:
            self._torch_buf = torch.zeros(self.requested, dtype=torch.uint8, device=self.device)
            self.ptrs = [self._torch_buf.data_ptr()]
            self.nbytes = self.requested
            return
        buf = symm.empty(self.requested, dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(buf, group=dist.group.WORLD)
        buf.zero_()
        self._torch_buf, self._torch_hdl = buf, hdl
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.nbytes = self.requested
        try:
            self.mc_ptr = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", True) else 0
        except Exception:
            self.mc_ptr = 0
        torch.cuda.synchronize()
        self.world.barrier()

    # ------------------------------------------------------------------------------------
    @property
    def local_ptr(self) -> int:
        return self.ptrs[self.rank]

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0

    def tensor(self, offset: int, nbytes: int, dtype: torch.dtype, rank: Optional[int] = None) -> torch.Tensor:
        """A tensor view of ``[offset, offset+nbytes)`` of rank ``rank``'s block (default: own)."""
        r = self.rank if rank is None else rank
        if offset % 16 or offset + nbytes > self.nbytes:
            raise ValueError("arena view out of range / misaligned")
        if self._native is not None:
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            t = ext.cuda().blob_tensor(self.ptrs[r] + offset, nbytes, dev_index, self._native)
        else:
            if r == self.rank:
                t = self._torch_buf[offset: offset + nbytes]
            else:
                t = self._torch_hdl.get_buffer(r, (nbytes,), torch.uint8, offset)
        return t.view(dtype)

    def close(self) -> None:
        self._native = None
        self._torch_buf = None
        self._torch_hdl = None
