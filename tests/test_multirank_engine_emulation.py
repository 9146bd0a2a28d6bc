"""N ranks of the REAL device engine driving the REAL kernel source — on the CPU, in one process.

Every rank is a Python thread with its own model, ``ps.SGD`` / ``ps.Adam`` and :class:`DeviceEngine`; the "symmetric arenas" are
plain host buffers every thread can address (one address space = NVLink peer mappings), the extension module is replaced by a
shim that executes the repository's ``ps_kernels.cu`` through the CPU emulator (``tests/_cuda_emu.py``, real launchers
included), and a rank's launches run synchronously in program order — i.e. one in-order stream per rank.  A launch that would
spin on a flag (``psb_wait_kernel``, ``psb_select_kernel``) first polls the flag words from Python so that the other ranks' threads
keep running, then executes the real kernel.

What this pins, in every CPU round: the flag protocol as the engine really drives it (progress values per chunk and step, slots,
masks, PARAMS_READY / CONSUMED / ACK / DONE), the chunk pipeline across ranks, inactive parameters, the async server loop with
device-side selection, staleness and the consistent-read sequence lock — against a single-process fp32 oracle for the synchronous
modes and against the protocol's invariants for AsySG-InCon.  The same scenarios run on hardware in ``test_gpu_engine.py``."""
import contextlib
import ctypes
import threading
import time
from functools import partial

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200 import runtime
from pytorch_ps_mpi_b200.parallel import device_engine as de
from tests import _cuda_emu
from tests.test_device_engine_control_flow import FakeEvent, FakeStream

DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
DONE = 1 << 62
_tls = threading.local()
_EXT = None          # set by the fixture: None = Python shim over ctypes, else the real bindings module


@pytest.fixture(autouse=True)
def _single_threaded_torch():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _p(x):
    return ctypes.c_void_p(int(x))


def _words(ptr, n=512):
    return (ctypes.c_uint64 * n).from_address(int(ptr))


class Cluster:
    def __init__(self, lib, n):
        self.lib, self.n = lib, n
        self.lock = threading.Lock()                  # one emulated kernel at a time (the emulator is single-threaded)
        self.bar = threading.Barrier(n)
        self.box = [None] * n
        self.failed = None
        self.multicast = False
        self.jitter_s = 0.0                            # > 0: every launch is preceded by a random pause (schedule fuzzing)

    def jitter(self):
        if self.jitter_s:
            import random
            time.sleep(random.random() * self.jitter_s)

    def fail(self, exc):
        if self.failed is None:
            self.failed = exc
        self.bar.abort()

    def poll(self, cond, what, timeout=60.0):
        """Block THIS rank (not the others) until ``cond()``; a stuck protocol fails the test instead of hanging it."""
        t0 = time.time()
        while not cond():
            if self.failed is not None:
                raise RuntimeError("another rank failed")
            if time.time() - t0 > timeout:
                raise TimeoutError(f"stuck waiting for {what}")
            time.sleep(0.0005)


class World:
    backend = "emulated"

    def __init__(self, cluster, rank):
        self.c, self.rank, self.size, self.local_rank = cluster, rank, cluster.n, rank
        self.device = torch.device("cpu")
        self.job_id = "emu"

    def barrier(self):
        if self.size > 1:
            self.c.bar.wait(timeout=120)

    def all_gather_object(self, obj):
        if self.size == 1:
            return [obj]
        self.c.box[self.rank] = obj
        self.barrier()
        out = list(self.c.box)
        self.barrier()
        return out

    def broadcast_object(self, obj, src=0):
        return self.all_gather_object(obj)[src]


class SharedArena:
    """``SymmetricArena`` over host memory: every rank's block is addressable by every thread."""

    def __init__(self, nbytes, device, world):
        self.buf = torch.zeros(nbytes + 64, dtype=torch.uint8)
        self.nbytes, self.rank, self.size = nbytes, world.rank, world.size
        self.bufs = world.all_gather_object(self.buf)
        self.ptrs = [b.data_ptr() for b in self.bufs]
        self.mc_ptr, self.provider = 0, "host-emulated"
        c = world.c
        if c.multicast and world.size > 1:
            # NVLS: a multicast window bound to every rank's block (symm_mem.cpp::mc_create / mc_bind_and_map); the window's
            # address range is a dummy allocation that is never dereferenced
            self.mc_buf = world.broadcast_object(torch.zeros(nbytes + 64, dtype=torch.uint8) if world.rank == 0 else None, src=0)
            if world.rank == 0:
                with c.lock:
                    assert c.lib.emu_mc_register(_p(self.mc_buf.data_ptr()), ctypes.c_size_t(nbytes), world.size,
                                                 (ctypes.c_void_p * world.size)(*self.ptrs)) >= 0
            world.barrier()
            self.mc_ptr = self.mc_buf.data_ptr()

    local_ptr = property(lambda self: self.ptrs[self.rank])
    has_multicast = property(lambda self: self.mc_ptr != 0)

    def tensor(self, offset, nbytes, dtype, rank=None):
        b = self.bufs[self.rank if rank is None else rank]
        return b[offset: offset + nbytes].view(dtype)

    def close(self):
        pass


class Plan:
    def __init__(self, m):
        self.m = m
        self.kind = self.wire = self.opt = self.grid = 0
        self.window_bytes = 128 << 20
        self.rank_ptrs = {}

    def set_rank_ptrs(self, r, wire_p, scales_p, param_p, signal_p):
        self.rank_ptrs[r] = (wire_p, scales_p, param_p, signal_p)

    def configure(self, world, rank, ntiles, bpt, cap, param_dt, bcast, reduce, param_mc, wire_mc, param_local, master, buf0, buf1,
                  buf2, tiles, signal_local, done_counter, stats):
        self.mc = (param_mc, wire_mc, reduce)
        self.c = dict(world=world, rank=rank, ntiles=ntiles, bpt=bpt, cap=cap, param_dt=param_dt, bcast=bcast,
                      param_local=param_local, master=master, buf0=buf0, buf1=buf1, buf2=buf2, tiles=tiles,
                      signal_local=signal_local, done_counter=done_counter, stats=stats)

    def launch(self, epoch, groups, contrib_mask, inv_count, wait_grads, signal_mode, ack_mask=0, version=0, select_out=0,
               average_dynamic=0, active_ptr=0, timeout_s=30.0, wait_mask=0xffffffff, stream=0, tile_begin=0, tile_end=-1,
               wait_value=0, param_hyper=0):
        self.m.cluster.jitter()
        c, lib = self.c, self.m.lib
        assert wait_grads == 0          # the engine always waits with the one-warp kernel
        if tile_end < 0:
            tile_begin, tile_end, wait_value = 0, c["ntiles"], epoch
        n = c["world"]
        arr = lambda xs: (ctypes.c_void_p * len(xs))(*xs)      # noqa: E731
        flat = [float(x) for g in groups for x in g]
        w, s, p, sig = (arr([self.rank_ptrs[r][i] for r in range(n)]) for i in range(4))
        with self.m.cluster.lock:
            lib.emu_update_extra(ctypes.c_uint64(version), _p(select_out), average_dynamic, ctypes.c_double(2.0))
            lib.emu_update_mc(_p(self.mc[0]), _p(self.mc[1]), self.mc[2])
            rc = lib.emu_update(self.kind, self.wire, self.opt, n, c["rank"], w, s, p, _p(c["param_local"]), _p(c["master"]),
                                _p(c["buf0"]), _p(c["buf1"]), _p(c["buf2"]), _p(c["tiles"]), _p(active_ptr), _p(param_hyper),
                                _p(c["signal_local"]), sig, _p(c["done_counter"]), _p(c["stats"]),
                                (ctypes.c_float * len(flat))(*flat), len(groups), c["ntiles"], c["bpt"], c["cap"], c["param_dt"],
                                c["bcast"], ctypes.c_uint32(contrib_mask), ctypes.c_uint32(wait_mask), ctypes.c_float(inv_count),
                                ctypes.c_uint64(epoch), ctypes.c_uint64(wait_value), tile_begin, tile_end, 0, signal_mode,
                                ctypes.c_uint32(ack_mask), min(3, tile_end - tile_begin))
        assert rc == 0
        self.m.log.append(("update", tile_begin, tile_end, signal_mode))


class M:
    """The extension module as the engine sees it, executing on the emulator."""
    TILE, SIGNAL_SLOTS, MAX_RANKS, MAX_GROUPS = 2048, 512, 16, 16
    SIG_GRAD_READY, SIG_PARAMS_READY, SIG_CONSUMED, SIG_ERROR, SIG_VERSION = 0, 64, 128, 200, 201
    SIG_ACK, SIG_GRAD_VERSION, SIG_STAGE_BEGIN, SIG_SEEN_VERSION = 256, 320, 202, 203

    def __init__(self, cluster):
        self.cluster, self.lib, self.log = cluster, cluster.lib, []

    def UpdatePlan(self):
        return Plan(self)

    def update_max_grid(self, *a):
        return 444

    def launch_count(self):
        return 0

    def encode(self, kind, wire, grads, first_tile, ntiles, param_idx, tiles_ptr, wire_ptr, scales_ptr, amax_ptr, residual_ptr,
               bpt, cap, ratio, sig_targets, sig_slot, sig_value, sig_counter, stream):
        self.cluster.jitter()
        assert len(grads) > 0
        tg = (ctypes.c_void_p * max(len(sig_targets), 1))(*sig_targets)
        for base in range(0, len(grads), 64):          # PSB_ENCODE_MAX tensors per launch; only the last launch raises the flag
            sl = slice(base, base + 64)
            n = len(grads[sl])
            last = base + 64 >= len(grads)
            ia = lambda xs: (ctypes.c_int * n)(*xs[sl])      # noqa: E731
            with self.cluster.lock:
                rc = self.lib.emu_encode(kind, wire, n, (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads[sl]]), ia(first_tile),
                                         ia(ntiles), ia(param_idx), _p(tiles_ptr), _p(wire_ptr), _p(scales_ptr), _p(amax_ptr),
                                         _p(residual_ptr), bpt, cap, ctypes.c_double(ratio), DT[grads[0].dtype], tg,
                                         len(sig_targets) if last else 0, sig_slot, ctypes.c_uint64(sig_value), _p(sig_counter))
            assert rc == 0
        assert rc == 0
        self.log.append(("encode", list(first_tile), sig_value if sig_targets else None))

    def signal(self, targets, slot, value, extra_slot=-1, extra_value=0, stream=0, version_local=0, version_slot=0):
        self.cluster.jitter()
        tg = (ctypes.c_void_p * len(targets))(*targets)
        with self.cluster.lock:
            self.lib.emu_signal(tg, len(targets), slot, ctypes.c_uint64(value), _p(1 if extra_slot >= 0 else 0),
                                max(extra_slot, 0), ctypes.c_uint64(extra_value), _p(version_local), version_slot)
        self.log.append(("signal", slot, value))

    def wait_flags(self, signal_local, slot0, mask, want, timeout_s, stream=0):
        self.cluster.jitter()
        sig = _words(signal_local)
        try:
            self.cluster.poll(lambda: sig[self.SIG_ERROR] != 0 or all(sig[slot0 + r] >= want for r in range(32) if mask >> r & 1),
                              f"slot {slot0} mask {mask:#x} >= {want}", timeout=min(timeout_s, 60.0))
        except TimeoutError:
            if timeout_s >= 60.0:
                raise                       # a stuck protocol, not a scenario with a deliberately short device time-out
        with self.cluster.lock:             # the real kernel: returns at once, or times out and poisons SIG_ERROR
            self.lib.emu_wait(_p(signal_local), slot0, ctypes.c_uint32(mask), ctypes.c_uint64(want), ctypes.c_double(0.01))
        assert sig[self.SIG_ERROR] == 0 or timeout_s < 60.0, "device wait timed out"
        self.log.append(("wait", slot0, mask, want))

    def select_ready(self, signal_local, consumed, cand_mask, quota, out, timeout_s, version=0, begin_targets=(), stream=0):
        self.cluster.jitter()
        sig, cons = _words(signal_local), _words(consumed, 64)

        def ready():
            fin = [r for r in range(32) if cand_mask >> r & 1 and sig[r] >= DONE]
            rdy = [r for r in range(32) if cand_mask >> r & 1 and r not in fin and sig[r] > cons[r]]
            need = min(quota, bin(cand_mask).count("1") - len(fin))
            return need == 0 or len(rdy) >= need
        self.cluster.poll(ready, "quota gradients")
        bt = (ctypes.c_void_p * max(len(begin_targets), 1))(*begin_targets)
        with self.cluster.lock:
            self.lib.emu_select(_p(signal_local), _p(consumed), ctypes.c_uint32(cand_mask), quota, _p(out), ctypes.c_uint64(version),
                                bt, len(begin_targets), ctypes.c_double(1.0))
        o = _words(out, 64)
        self.log.append(("select", int(o[0]), int(o[1]), [int(o[2 + r]) for r in range(16)]))

    def snapshot(self, signal_local, stage, shadow, params, nbytes, scratch, attempts=2, stream=0):
        with self.cluster.lock:
            self.lib.emu_snapshot(_p(signal_local), _p(stage), _p(shadow), _p(params), ctypes.c_size_t(nbytes), _p(scratch), attempts)


class PlanProxy:
    """The REAL ``UpdatePlan`` of ``csrc/bindings.cpp`` (argument marshalling, window loop), with a launch log."""

    def __init__(self, m):
        object.__setattr__(self, "_m", m)
        object.__setattr__(self, "_p", m.x.UpdatePlan())
        # 8 KB windows instead of 128 MB: every multi-tile launch walks the window loop (only the first window waits, only the
        # last one raises PARAMS_READY / CONSUMED / ACK) — on hardware that loop only runs for >= 128 MB chunks
        self._p.window_bytes = 8192

    def __setattr__(self, k, v):
        setattr(self._p, k, v)

    def __getattr__(self, k):
        return getattr(self._p, k)

    def launch(self, *a, **k):
        self._m.cluster.jitter()
        self._p.launch(*a, **k)
        self._m.log.append(("update", k.get("tile_begin", 0), k.get("tile_end", -1), a[5]))


class RealM:
    """``_psb200_emu``: the repository's real pybind11 bindings linked against the emulated kernels
    (``_cuda_emu.build_extension``).  Calls hold the GIL, which serialises the (single-threaded) emulator; launches that would
    spin on a flag are preceded by the same Python-side poll as in :class:`M`."""

    def __init__(self, cluster, ext):
        self.cluster, self.x, self.log = cluster, ext, []

    def __getattr__(self, k):
        return getattr(self.x, k)

    def UpdatePlan(self):
        return PlanProxy(self)

    def encode(self, *a, **k):
        self.cluster.jitter()
        self.x.encode(*a, **k)
        self.log.append(("encode", list(a[3]), a[16] if a[14] else None))

    def signal(self, *a, **k):
        self.cluster.jitter()
        self.x.signal(*a, **k)
        self.log.append(("signal", a[1], a[2]))

    def wait_flags(self, signal_local, slot0, mask, want, timeout_s, stream=0):
        self.cluster.jitter()
        sig = _words(signal_local)
        try:
            self.cluster.poll(lambda: sig[M.SIG_ERROR] != 0 or all(sig[slot0 + r] >= want for r in range(32) if mask >> r & 1),
                              f"slot {slot0} mask {mask:#x} >= {want}", timeout=min(timeout_s, 60.0))
        except TimeoutError:
            if timeout_s >= 60.0:
                raise
        self.x.wait_flags(signal_local, slot0, mask, want, 0.01, stream)
        assert sig[M.SIG_ERROR] == 0 or timeout_s < 60.0, "device wait timed out"
        self.log.append(("wait", slot0, mask, want))

    def select_ready(self, signal_local, consumed, cand_mask, quota, out, timeout_s, version=0, begin_targets=(), stream=0):
        self.cluster.jitter()
        sig, cons = _words(signal_local), _words(consumed, 64)

        def ready():
            fin = [r for r in range(32) if cand_mask >> r & 1 and sig[r] >= DONE]
            rdy = [r for r in range(32) if cand_mask >> r & 1 and r not in fin and sig[r] > cons[r]]
            need = min(quota, bin(cand_mask).count("1") - len(fin))
            return need == 0 or len(rdy) >= need
        self.cluster.poll(ready, "quota gradients")
        self.x.select_ready(signal_local, consumed, cand_mask, quota, out, 1.0, version, list(begin_targets), stream)
        o = _words(out, 64)
        self.log.append(("select", int(o[0]), int(o[1]), [int(o[2 + r]) for r in range(16)]))


@pytest.fixture(params=["shim", "bindings"])
def emu(monkeypatch, request):
    global _EXT
    if request.param == "bindings":
        # the extension module as built from csrc/bindings.cpp; its library also carries the emulator's own entry points
        _EXT = _cuda_emu.build_extension()
        lib = None if _EXT is None else _EXT.emu
        if lib is not None:
            lib.emu_set_sm_count(1)         # update grid = 3 CTAs: several tiles per CTA (grid-stride loops), as on big arenas
    else:
        _EXT, lib = None, _cuda_emu.build()
    if lib is None:
        pytest.skip("no g++")
    _tls.world, _tls.m = World(Cluster(lib, 1), 0), None      # the test's own thread: a single-process world (oracles)
    monkeypatch.setattr(runtime, "world", lambda: _tls.world)
    monkeypatch.setattr(de.ext, "cuda", lambda: _tls.m)
    monkeypatch.setattr(de, "SymmetricArena", SharedArena)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setenv("PSB200_CHUNK_BYTES", str(2048 * 4))              # one tile per chunk: a real pipeline on a tiny model
    return lib


def run_ranks(lib, n, fn, multicast=False, jitter_s=0.0):
    """Run ``fn(rank, world)`` on n threads; returns their results, re-raises the first failure."""
    cluster = Cluster(lib, n)
    cluster.multicast, cluster.jitter_s = multicast, jitter_s
    lib.emu_mc_clear()
    out, errs = [None] * n, []

    def main(r):
        _tls.world, _tls.m = World(cluster, r), (M(cluster) if _EXT is None else RealM(cluster, _EXT))
        try:
            out[r] = fn(r, _tls.world)
        except BaseException as exc:       # noqa: BLE001
            errs.append((r, exc))
            cluster.fail(exc)

    ts = [threading.Thread(target=main, args=(r,), daemon=True) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "a rank thread is stuck"
    real = [e for e in errs if not isinstance(e[1], (threading.BrokenBarrierError, RuntimeError)) or "another rank" not in str(e[1])]
    if errs:
        raise (real or errs)[0][1]
    return out


_rng_lock = threading.Lock()


def _model(dtype=torch.float32):
    with _rng_lock:                       # the global RNG is shared by the rank threads: seed + init must not interleave
        torch.manual_seed(0)
        return _build(dtype)


def _build(dtype):
    return torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Tanh(), torch.nn.Linear(30, 24), torch.nn.Tanh(),
                               torch.nn.Linear(24, 10)).to(dtype)


def _data(rank, step, dtype=torch.float32):
    g = torch.Generator().manual_seed(1000 * rank + step)
    return torch.randn(8, 20, generator=g).to(dtype), torch.randint(0, 10, (8,), generator=g)


def _loss(model, x, y, skip_head):
    h = model[:-1](x)
    out = h[:, :10] if skip_head else model[-1](h)
    return torch.nn.functional.cross_entropy(out.float(), y)


def _attach(opt, **kw):
    for h in opt._hooks:
        h.remove()
    opt._engine = de.DeviceEngine(opt, **kw)
    opt._hooks = [p.register_hook(partial(opt._engine.on_grad, name=n, param=p)) for n, p in opt._named.items()]


def _oracle(n, steps, optim, hyper, average, skip_until):
    """Single process: every rank's gradient on the SAME parameters, summed in rank order, then one ``torch.optim.SGD`` step or
    the reference's Adam rule (``ps.Adam.optim_step`` = ``/root/reference/ps.py:217-261``; parameters without a gradient are
    skipped and keep their own step count, ``ps.py:178-179``)."""
    model = _model()
    if optim == "sgd":
        opt = torch.optim.SGD(model.parameters(), **hyper)
    else:
        opt = ps.Adam(model.named_parameters(), model.parameters(), engine="host", **hyper)
        for h in opt._hooks:
            h.remove()
    for s in range(steps):
        tot = None
        for r in range(n):
            model.zero_grad(set_to_none=True)
            _loss(model, *_data(r, s), skip_head=s < skip_until).backward()
            gs = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
            tot = gs if tot is None else [a if b is None else a + b for a, b in zip(tot, gs)]
        tot = [None if g is None else (g / n if average else g) for g in tot]
        if optim == "sgd":
            for p, g in zip(model.parameters(), tot):
                p.grad = g
            opt.step()
        else:
            with torch.no_grad():
                for p, g in zip(model.parameters(), tot):
                    if g is not None:
                        opt.optim_step(p, g, betas=(0.9, 0.999), eps=1e-8, lr=hyper["lr"], weight_decay=hyper["weight_decay"])
    if optim != "sgd":
        opt.close()
    return [p.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("n,mode,optim,average,skip_until", [
    (2, "ps", "sgd", False, 0), (3, "ps", "sgd", True, 0), (3, "allgather", "sgd", False, 0), (2, "ps", "adam", True, 0),
    (2, "ps", "sgd", False, 2), (2, "allgather", "adam", False, 2)])
def test_sync_modes_match_single_process_oracle(emu, n, mode, optim, average, skip_until):
    """PS / all-gather at 2-3 ranks, 4 steps, per-chunk pipeline (one tile per chunk), optionally with a parameter (the head) that
    gets no gradient on ANY rank for the first steps: every rank ends on the oracle's parameters."""
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-3, dampening=0.2) if optim == "sgd" else dict(lr=1e-2, weight_decay=1e-2)
    steps = 4

    def rank_main(rank, w):
        model = _model()
        cls = ps.SGD if optim == "sgd" else ps.Adam
        opt = cls(model.named_parameters(), model.parameters(), engine="host", mode=mode, average=average, **hyper)
        assert opt.rank == rank and opt.size == n
        _attach(opt)
        eng = opt._engine
        assert eng.nchunks >= 3 and eng.pipeline and eng.is_server == (mode == "allgather" or rank == 0)
        for s in range(steps):
            if mode == "ps" and rank != 0 and s > 0:
                assert _words(eng.arena.local_ptr)[M.SIG_PARAMS_READY] == s         # the forward below reads published weights
            opt.zero_grad(set_to_none=True)
            _loss(model, *_data(rank, s), skip_head=s < skip_until).backward()
            opt.step()
        eng.check()
        w.barrier()
        mine = [p.detach().clone() for p in model.parameters()]
        sig = list(_words(eng.arena.local_ptr))
        log = list(_tls.m.log)
        opt.close()
        return mine, sig, log, eng.nchunks

    res = run_ranks(emu, n, rank_main)
    want = _oracle(n, steps, optim, hyper, average, skip_until)
    for mine, _, _, _ in res:
        for a, b in zip(mine, want):
            assert torch.allclose(a, b, rtol=3e-5, atol=3e-6), float((a - b).abs().max())
    for r, (mine, sig, log, nchunks) in enumerate(res):
        for a, b in zip(mine, res[0][0]):
            assert torch.equal(a, b)                                   # ranks bit-identical
        assert sig[M.SIG_ERROR] == 0
        if mode == "ps":
            assert sig[M.SIG_PARAMS_READY] == steps
            if r == 0:
                assert all(sig[M.SIG_GRAD_READY + q] == steps * nchunks for q in range(1, n))       # monotone progress values
        else:
            assert all(sig[M.SIG_CONSUMED + q] == steps for q in range(n) if q != r)
            assert all(sig[M.SIG_GRAD_READY + q] == steps * nchunks for q in range(n) if q != r)
        ups = [e for e in log if e[0] == "update"]
        if mode == "allgather" or r == 0:
            assert len(ups) == steps * nchunks and [u[3] for u in ups].count(0) == steps * (nchunks - 1)
        else:
            assert not ups                                             # workers never run the update kernel in PS mode


def test_async_server_with_device_selection(emu):
    """AsySG-InCon at 3 ranks (server + 2 workers), quota 1: every gradient is applied exactly once, acknowledged, its staleness
    recorded; after the drain every rank holds the server's final parameters."""
    nsteps, n = 3, 3

    def rank_main(rank, w):
        model = _model()
        opt = ps.SGD(model.named_parameters(), model.parameters(), engine="host", mode="async", quota=1, lr=0.05, average=True)
        _attach(opt)
        eng = opt._engine
        seen = []
        if rank == 0:
            applied = opt.serve()
            assert applied == nsteps * (n - 1), applied
            seen = dict(eng._async_last)
        else:
            for s in range(nsteps):
                opt.zero_grad(set_to_none=True)
                _loss(model, *_data(rank, s), skip_head=False).backward()
                _, data = opt.step()
                time.sleep(0.002 * rank)
        opt.close()                                # workers: wait for the last ACK, then post DONE
        sig = list(_words(eng.arena.local_ptr))
        return [p.detach().clone() for p in model.parameters()], sig, seen

    res = run_ranks(emu, n, rank_main)
    before = [p.detach() for p in _model().parameters()]
    assert not torch.equal(res[0][0][0], before[0])
    assert all(torch.isfinite(p).all() for p in res[0][0])
    srv = res[0][1]
    for r in (1, 2):
        assert res[r][1][M.SIG_ACK] == nsteps                      # every gradient of every worker was acknowledged
        assert srv[M.SIG_GRAD_READY + r] == DONE                   # ... and the worker said goodbye
    last = res[0][2]
    assert last["updates_applied"] == nsteps * (n - 1) and last["param_version"] == nsteps * (n - 1)
    assert all(v >= 0 for v in last["staleness"].values()) and set(last["staleness"]) == set(last["contributors"])


def test_async_consistent_reads(emu):
    """``consistent=True``: workers adopt whole versions through the device-side sequence lock; versions never go backwards and
    every rank leaves with the server's final parameters."""
    nsteps, n = 3, 2

    def rank_main(rank, w):
        model = _model()
        opt = ps.SGD(model.named_parameters(), model.parameters(), engine="host", mode="async", quota=1, lr=0.05, consistent=True)
        _attach(opt)
        eng = opt._engine
        assert eng.consistent
        seen = []
        if rank == 0:
            assert opt.serve() == nsteps * (n - 1)
        else:
            for s in range(nsteps):
                opt.zero_grad(set_to_none=True)
                _loss(model, *_data(rank, s), skip_head=False).backward()
                _, data = opt.step()
                seen.append(data["param_version"])
        opt.close()
        return [p.detach().clone() for p in model.parameters()], seen, eng._snap_version

    res = run_ranks(emu, n, rank_main)
    assert res[1][1] == sorted(res[1][1])
    assert res[1][2] == nsteps * (n - 1) and res[0][2] == nsteps * (n - 1)          # the final snapshot is the last version
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,mode,optim,coding,dtype", [
    (2, "ps", "sgd", "topk", torch.float32), (2, "allgather", "sgd", "scale", torch.float32), (3, "ps", "adam", "cast", torch.float32),
    (2, "ps", "sgd", "identity", torch.bfloat16), (2, "allgather", "adam", "topk", torch.bfloat16)])
def test_codings_and_bf16_masters_multirank(emu, n, mode, optim, coding, dtype):
    """Coded wires (block-wise top-k, abs-max int8, bf16 cast) and bf16 parameters with fp32 masters across ranks: each step all
    ranks' ACTUAL gradients are gathered, ``decode(encode(.))`` applied per rank, summed in fp32 in rank order and fed to the
    reference optimizer math on fp32 shadows (the oracle of ``tests/_mp.py::gpu_train``); the engine's masters must match, the
    published parameters must be the rounded masters, and all ranks must be bit-identical."""
    factory = {"identity": ps.Identity, "cast": lambda: ps.Cast("bf16"), "scale": lambda: ps.Scale("int8"),
               "topk": lambda: ps.TopK(ratio=0.25)}[coding]
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-4) if optim == "sgd" else dict(lr=1e-2, eps=1e-8)
    steps = 3

    def rank_main(rank, w):
        model = _model(dtype)
        shadow = [torch.nn.Parameter(p.detach().float().clone()) for p in model.parameters()]
        cls = ps.SGD if optim == "sgd" else ps.Adam
        oracle = cls([(f"p{i}", q) for i, q in enumerate(shadow)], shadow, engine="host", use_mpi=False, **hyper)
        for h in oracle._hooks:
            h.remove()
        groups = oracle._group_of()
        opt = cls(model.named_parameters(), model.parameters(), engine="host", mode=mode, code=factory(), **hyper)
        _attach(opt)
        eng = opt._engine
        assert (eng.master is not None) == (dtype != torch.float32 and eng.is_server)
        for s in range(steps):
            opt.zero_grad(set_to_none=True)
            x, y = _data(rank, s, dtype)
            _loss(model, x, y, skip_head=False).backward()
            mine = [p.grad.detach().clone() for p in model.parameters()]
            opt.step()
            allg = w.all_gather_object(mine)
            with torch.no_grad():
                for i, q in enumerate(shadow):
                    total = torch.zeros_like(q)
                    for r in range(n):
                        code = factory()
                        total += code.decode(code.encode(allg[r][i], name=f"p{i}")).reshape(q.shape).float()
                    oracle.optim_step(q, total, **oracle._hyper(groups[id(q)]))
        eng.check()
        w.barrier()
        got = [(opt.state[p]["master_param"] if eng.master is not None else p).detach().float().clone() for p in model.parameters()]
        pub = [p.detach().clone() for p in model.parameters()]
        opt.close()
        oracle.close()
        return got, pub, [q.detach().clone() for q in shadow], eng.is_server

    res = run_ranks(emu, n, rank_main)
    for got, pub, shadow, is_server in res:
        for a, b in zip(pub, res[0][1]):
            assert torch.equal(a, b)                                   # ranks bit-identical
        if is_server:
            for g, q, p in zip(got, shadow, pub):
                assert torch.allclose(g, q, rtol=2e-4, atol=2e-5), (coding, float((g - q).abs().max()))
                if dtype != torch.float32:
                    assert torch.equal(g.to(dtype), p)                 # published parameter == the rounded master


@pytest.mark.parametrize("mode", ["ps", "allgather"])
def test_stalled_peer_is_detected_then_recovered(emu, monkeypatch, mode):
    """Failure detection and recovery: a rank that sits a step out makes its peers' bounded device waits time out —
    the error slot is poisoned, the update kernels of that step return without touching parameters or raising their flags,
    ``check()`` raises; ``recover()`` (collective) realigns clocks, flags, parameters (and replicated state), after which
    training continues and ends exactly where a run WITHOUT the failed step ends."""
    monkeypatch.setenv("PSB200_DEVICE_TIMEOUT", "0.3")
    hyper = dict(lr=0.05, momentum=0.9)
    order = [0, 1, 2, 3]                                                # data steps; rank 1 stalls at step 1

    def rank_main(rank, w):
        model = _model()
        opt = ps.SGD(model.named_parameters(), model.parameters(), engine="host", mode=mode, **hyper)
        _attach(opt)
        eng = opt._engine
        raised = None
        for s in order:
            if s == 1:
                before = [p.detach().clone() for p in model.parameters()]
                if rank == 0:
                    opt.zero_grad(set_to_none=True)
                    _loss(model, *_data(rank, s), skip_head=False).backward()
                    opt.step()                                         # rank 1 never posts: every wait of this step times out
                    assert all(torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))     # nothing applied
                    if mode == "ps":
                        assert _words(eng.arena.local_ptr)[M.SIG_PARAMS_READY] == 1
                    try:
                        eng.check()
                        raised = False
                    except RuntimeError as exc:
                        raised = "timed out" in str(exc)
                w.barrier()
                eng.recover()
                assert _words(eng.arena.local_ptr)[M.SIG_ERROR] == 0 and eng._epoch == 0
                continue
            opt.zero_grad(set_to_none=True)
            _loss(model, *_data(rank, s), skip_head=False).backward()
            opt.step()
        eng.check()
        w.barrier()
        mine = [p.detach().clone() for p in model.parameters()]
        opt.close()
        return raised, mine

    res = run_ranks(emu, 2, rank_main)
    assert res[0][0] is True and res[1][0] is None
    # oracle: steps 0, 2, 3 of both ranks (the failed step contributed nothing)
    model = _model()
    o = torch.optim.SGD(model.parameters(), **hyper)
    for s in (0, 2, 3):
        tot = None
        for r in range(2):
            model.zero_grad(set_to_none=True)
            _loss(model, *_data(r, s), skip_head=False).backward()
            gs = [p.grad.clone() for p in model.parameters()]
            tot = gs if tot is None else [a + b for a, b in zip(tot, gs)]
        for p, g in zip(model.parameters(), tot):
            p.grad = g
        o.step()
    for _, mine in res:
        for a, b in zip(mine, model.parameters()):
            assert torch.allclose(a, b.detach(), rtol=3e-5, atol=3e-6), float((a - b.detach()).abs().max())


class _DirectLinear(torch.autograd.Function):
    """A producer we own: writes dW / db straight into the views handed out by ``param.ps_grad_out()`` (what the fused BN backward,
    the stem wgrad and ``BcastLinear`` do on the GPU) and returns those views as the gradients."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.params = (w, b)
        return torch.addmm(b, x, w.t())

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        wp, bp = ctx.params
        gw = getattr(wp, "ps_grad_out", lambda: None)()
        gb = getattr(bp, "ps_grad_out", lambda: None)()
        gx = gy @ w
        gw = torch.mm(gy.t(), x, out=gw) if gw is not None else gy.t() @ x
        if gb is not None:
            torch.sum(gy, 0, out=gb)
        else:
            gb = gy.sum(0)
        return gx, gw, gb


class _DirectNet(torch.nn.Module):
    def __init__(self, base, direct_layers):
        super().__init__()
        self.base, self.direct_layers = base, direct_layers

    def forward(self, x):
        for i, m in enumerate(self.base):
            x = _DirectLinear.apply(x, m.weight, m.bias) if i in self.direct_layers else m(x)
        return x


@pytest.mark.parametrize("option", ["unpipelined", "gate", "direct", "direct+inactive"])
def test_engine_options_multirank(emu, monkeypatch, option):
    """The same 3-rank PS-SGD run through the engine's other paths: one fused launch per step (``pipeline=False``), a registered
    gate (workers queue no wait kernel; the consumer acquires ``gate()``'s flag itself), and direct gradient placement (producers
    write into the wire arena, mixed with encoded gradients inside a chunk; with a frozen-for-a-while parameter on top)."""
    n, steps = 3, 4
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-3)
    skip_until = 2 if option == "direct+inactive" else 0

    def rank_main(rank, w):
        base = _model()
        model = _DirectNet(base, {0, 4} if option.startswith("direct") else set())
        opt = ps.SGD(base.named_parameters(), base.parameters(), engine="host", mode="ps", pipeline=option != "unpipelined", **hyper)
        _attach(opt)
        eng = opt._engine
        assert eng.nchunks == (1 if option == "unpipelined" else 6)          # six parameters, one tile each
        if option == "gate":
            eng.register_gate(object())
        for s in range(steps):
            if option == "gate":
                flag_ptr, epoch = eng.gate()              # what the first forward GEMM's TMA producer acquires on the GPU
                assert (flag_ptr == 0) == (rank == 0 or s == 0) and epoch == (0 if rank == 0 else s)
                if flag_ptr:
                    _tls.m.cluster.poll(lambda: _words(flag_ptr, 1)[0] >= epoch, "gated PARAMS_READY")
            opt.zero_grad(set_to_none=True)
            x, y = _data(rank, s)
            h = model(x) if s >= skip_until else _DirectNet(base[:-1], {0})(x)[:, :10]
            torch.nn.functional.cross_entropy(h, y).backward()
            opt.step()
        if option == "gate":
            eng.ensure_params()                          # nothing consumed the last broadcast through a gate: plain wait
        eng.check()
        w.barrier()
        mine = [p.detach().clone() for p in base.parameters()]
        waits = [e for e in _tls.m.log if e[0] == "wait" and e[1] == M.SIG_PARAMS_READY]
        direct = eng.direct_grads
        opt.close()
        return mine, len(waits), direct

    res = run_ranks(emu, n, rank_main)
    want = _oracle(n, steps, "sgd", hyper, False, skip_until)
    for r, (mine, waits, direct) in enumerate(res):
        for a, b in zip(mine, want):
            assert torch.allclose(a, b, rtol=3e-5, atol=3e-6), (option, r, float((a - b).abs().max()))
        if option == "gate":
            assert waits == (0 if r == 0 else 1)                       # only the final ensure_params()
        elif r > 0:
            assert waits == steps
        if option == "direct":
            assert direct == 4 * steps                                 # two layers x (weight, bias) per step, no encode pass
        if option == "direct+inactive":
            assert direct == 2 * skip_until + 4 * (steps - skip_until)


@pytest.mark.parametrize("optim", ["sgd", "adam"])
def test_checkpoint_resume_multirank(emu, optim):
    """``state_dict()`` / ``load_state_dict()`` through the device engine at 2 ranks, bf16 parameters with fp32 masters, a
    parameter whose first gradient arrives late (its own step count must survive the round trip): 2 steps + save + load into
    fresh objects + 2 steps == 4 straight steps, bit for bit."""
    import copy
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-4, dampening=0.1) if optim == "sgd" else dict(lr=1e-2, weight_decay=1e-2)

    def rank_main(rank, w):
        def make():
            m = _model(torch.bfloat16)
            cls = ps.SGD if optim == "sgd" else ps.Adam
            o = cls(m.named_parameters(), m.parameters(), engine="host", mode="ps", **hyper)
            _attach(o)
            return m, o

        def run(m, o, steps, start=0):
            for s in range(start, start + steps):
                o.zero_grad(set_to_none=True)
                x, y = _data(rank, s, torch.bfloat16)
                _loss(m, x, y, skip_head=s < 1).backward()          # the head sits out step 0
                o.step()

        m1, o1 = make()
        run(m1, o1, 4)
        want = [p.detach().clone() for p in m1.parameters()]
        m2, o2 = make()
        run(m2, o2, 2)
        sd_model = {k: v.clone() for k, v in m2.state_dict().items()}
        sd_opt = copy.deepcopy(o2.state_dict())
        if rank == 0:
            st = list(sd_opt["state"].values())
            assert all("master_param" in s for s in st) and sorted(int(s["step"]) for s in st) == [1, 1, 2, 2, 2, 2]
        m3, o3 = make()
        with torch.no_grad():
            for k, v in m3.state_dict().items():
                v.copy_(sd_model[k])
        o3.load_state_dict(sd_opt)
        run(m3, o3, 2, start=2)
        got = [p.detach().clone() for p in m3.parameters()]
        w.barrier()
        for o in (o1, o2, o3):
            o.close()
        return got, want

    for got, want in run_ranks(emu, 2, rank_main):
        for a, b in zip(got, want):
            assert torch.equal(a, b), float((a.float() - b.float()).abs().max())


@pytest.mark.parametrize("n,optim,dtype,reduce", [(4, "sgd", torch.bfloat16, "auto"), (4, "adam", torch.float32, "auto"),
                                                  (2, "sgd", torch.bfloat16, "nvls"), (3, "sgd", torch.float32, "auto")])
def test_switch_reduction_and_multicast_publish_multirank(emu, n, optim, dtype, reduce):
    """With multicast memory the engine publishes through ``multimem.st`` and — at N >= 4 (or when forced) — reduces through
    ``multimem.ld_reduce``: same grad-gather oracle as the GPU suite, with its tolerance for the switch's single rounding of the
    16-bit sums (``tests/_mp.py::gpu_train``)."""
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-4) if optim == "sgd" else dict(lr=1e-2, eps=1e-8)
    steps = 3

    def rank_main(rank, w):
        model = _model(dtype)
        shadow = [torch.nn.Parameter(p.detach().float().clone()) for p in model.parameters()]
        cls = ps.SGD if optim == "sgd" else ps.Adam
        oracle = cls([(f"p{i}", q) for i, q in enumerate(shadow)], shadow, engine="host", use_mpi=False, **hyper)
        for h in oracle._hooks:
            h.remove()
        groups = oracle._group_of()
        opt = cls(model.named_parameters(), model.parameters(), engine="host", mode="ps", **hyper)
        _attach(opt, reduce=reduce)
        eng = opt._engine
        assert eng.arena.has_multicast and eng.bcast == 2
        assert eng.reduce == (1 if (n >= 4 or reduce == "nvls") else 0)
        sum_mag = 0.0
        for s in range(steps):
            opt.zero_grad(set_to_none=True)
            x, y = _data(rank, s, dtype)
            _loss(model, x, y, skip_head=False).backward()
            mine = [p.grad.detach().clone() for p in model.parameters()]
            opt.step()
            allg = w.all_gather_object(mine)
            with torch.no_grad():
                for i, q in enumerate(shadow):
                    total = sum(allg[r][i].float() for r in range(n))
                    sum_mag = max(sum_mag, float(total.abs().max()))
                    oracle.optim_step(q, total, **oracle._hyper(groups[id(q)]))
        eng.check()
        w.barrier()
        got = [(opt.state[p]["master_param"] if eng.master is not None else p).detach().float().clone() for p in model.parameters()]
        pub = [p.detach().clone() for p in model.parameters()]
        nvls = eng.reduce == 1
        opt.close()
        oracle.close()
        return got, pub, [q.detach().clone() for q in shadow], eng.is_server, sum_mag, nvls

    res = run_ranks(emu, n, rank_main, multicast=True)
    for got, pub, shadow, is_server, sum_mag, nvls in res:
        for a, b in zip(pub, res[0][1]):
            assert torch.equal(a, b)                                   # multimem.st: every rank bit-identical
        if is_server:
            lossy = nvls and dtype != torch.float32
            atol = 2e-5 + (steps * hyper["lr"] * sum_mag * 2.0 ** -8 if lossy else 0.0)
            for g, q in zip(got, shadow):
                assert torch.allclose(g, q, rtol=2e-4, atol=atol), float((g - q).abs().max())
    emu.emu_mc_clear()


@pytest.mark.parametrize("seed", range(6))
def test_randomised_schedules_keep_the_oracle(emu, seed):
    """Schedule fuzzing: random world size / mode / optimizer / averaging / inactive steps, and a random pause in front of EVERY
    launch of every rank (so chunk k of one rank meets chunk k+2 of another, servers run ahead of workers and vice versa).
    Whatever the interleaving, every rank must end on the single-process oracle's parameters."""
    import random
    rnd = random.Random(seed)
    n = rnd.choice([2, 3, 4])
    mode = rnd.choice(["ps", "ps", "allgather"])
    optim = rnd.choice(["sgd", "adam"])
    average = rnd.random() < 0.5
    skip_until = rnd.choice([0, 0, 1, 2])
    multicast = rnd.random() < 0.5
    steps = 4
    hyper = dict(lr=0.05, momentum=0.9, weight_decay=1e-3) if optim == "sgd" else dict(lr=1e-2, weight_decay=1e-2)

    def rank_main(rank, w):
        model = _model()
        cls = ps.SGD if optim == "sgd" else ps.Adam
        opt = cls(model.named_parameters(), model.parameters(), engine="host", mode=mode, average=average, **hyper)
        _attach(opt)
        for s in range(steps):
            opt.zero_grad(set_to_none=True)
            _loss(model, *_data(rank, s), skip_head=s < skip_until).backward()
            opt.step()
            time.sleep(random.random() * 0.003)
        opt._engine.check()
        w.barrier()
        mine = [p.detach().clone() for p in model.parameters()]
        opt.close()
        return mine

    res = run_ranks(emu, n, rank_main, multicast=multicast, jitter_s=0.002)
    want = _oracle(n, steps, optim, hyper, average, skip_until)
    for mine in res:
        for a, b in zip(mine, want):
            assert torch.allclose(a, b, rtol=5e-5, atol=5e-6), (seed, n, mode, optim, multicast, float((a - b).abs().max()))
    emu.emu_mc_clear()


@pytest.mark.parametrize("seed", range(5))
def test_randomised_async_applies_every_gradient_exactly_once(emu, seed):
    """AsySG-InCon under schedule fuzzing: random worker count, quota, consistent reads, per-worker step counts and pauses.
    Invariants of the protocol: every posted gradient (worker r, epoch e) is selected by exactly one update, in epoch order per
    worker; every worker is acknowledged up to its last gradient; after the drain all ranks hold the server's parameters."""
    import random
    rnd = random.Random(100 + seed)
    n = rnd.choice([2, 3, 4])
    quota = rnd.randint(1, n - 1)
    consistent = rnd.random() < 0.5
    nsteps = [0] + [rnd.randint(1, 4) for _ in range(n - 1)]

    def rank_main(rank, w):
        model = _model()
        opt = ps.SGD(model.named_parameters(), model.parameters(), engine="host", mode="async", quota=quota, lr=0.05,
                     average=True, consistent=consistent)
        _attach(opt)
        eng = opt._engine
        if rank == 0:
            opt.serve()
        else:
            for s in range(nsteps[rank]):
                opt.zero_grad(set_to_none=True)
                _loss(model, *_data(rank, s), skip_head=False).backward()
                opt.step()
                time.sleep(random.random() * 0.004)
        log = list(_tls.m.log)
        opt.close()
        sig = list(_words(eng.arena.local_ptr))
        return [p.detach().clone() for p in model.parameters()], sig, log

    res = run_ranks(emu, n, rank_main, jitter_s=0.002)
    taken = {r: [] for r in range(1, n)}
    for kind, mask, cnt, epochs in (e for e in res[0][2] if e[0] == "select"):
        assert bin(mask).count("1") == cnt <= quota
        for r in range(1, n):
            if mask >> r & 1:
                taken[r].append(epochs[r])
    for r in range(1, n):
        assert taken[r] == list(range(1, nsteps[r] + 1)), (seed, r, taken[r], nsteps[r])     # exactly once, in order
        assert res[r][1][M.SIG_ACK] == nsteps[r]
    for mine, _, _ in res[1:]:
        for a, b in zip(mine, res[0][0]):
            assert torch.equal(a, b)
    assert all(torch.isfinite(p).all() for p in res[0][0])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_many_odd_shaped_parameters_groups_and_encode_batches(emu, monkeypatch, dtype):
    """Layout / batching stress at 2 ranks: 150 parameters of awkward sizes (1 element … several tiles, a channels-last conv
    weight that keeps its physical order in the arena) in ONE big chunk → encode launches are split at 64 tensors per launch
    and only the last one raises the flag; three parameter groups with different hyper-parameters; a random third of the
    parameters sits out each step (the same ones on every rank).  Oracle: ``torch.optim.SGD`` per group on the summed gradients."""
    monkeypatch.setenv("PSB200_CHUNK_BYTES", str(64 << 20))               # everything in one chunk → > 64 tensors per encode call
    import random
    rnd = random.Random(5)
    sizes = [rnd.choice([1, 3, 7, 8, 17, 64, 300, 2048, 2049, 5000]) for _ in range(149)]
    steps, n = 3, 2
    groups_of = [i % 3 for i in range(150)]
    ghyper = [dict(lr=0.1, momentum=0.9, weight_decay=0.0), dict(lr=0.05, momentum=0.0, weight_decay=1e-2),
              dict(lr=0.02, momentum=0.5, weight_decay=1e-3, nesterov=True)]
    skip = [set(rnd.sample(range(150), 50)) for _ in range(steps)]

    def make_params():
        with _rng_lock:
            torch.manual_seed(3)
            ps_ = [torch.nn.Parameter(torch.randn(s).to(dtype)) for s in sizes]
            conv = torch.nn.Parameter(torch.randn(8, 4, 3, 3).to(dtype).contiguous(memory_format=torch.channels_last))
            return ps_ + [conv]

    def grads_for(rank, s, params):
        g = torch.Generator().manual_seed(1000 * s + rank)
        return [torch.randn(p.shape, generator=g).to(dtype) for p in params]

    def rank_main(rank, w):
        params = make_params()
        named = [(f"p{i}", p) for i, p in enumerate(params)]
        pgs = [dict(params=[p for i, p in enumerate(params) if groups_of[i] == g], **ghyper[g]) for g in range(3)]
        opt = ps.SGD(named, pgs, engine="host", mode="ps", lr=0.1)
        _attach(opt)
        eng = opt._engine
        assert eng.nchunks == 1 and eng.layout.nparams == 150
        for s in range(steps):
            opt.zero_grad(set_to_none=True)
            gs = grads_for(rank, s, params)
            loss = sum((p.float() * g.float()).sum() for i, (p, g) in enumerate(zip(params, gs)) if i not in skip[s])
            loss.backward()
            opt.step()
        eng.check()
        w.barrier()
        got = [(opt.state[p]["master_param"] if eng.master is not None else p).detach().float().clone() for p in params]
        pub = [p.detach().clone() for p in params]
        strides = params[-1].stride()
        enc = [e for e in _tls.m.log if e[0] == "encode"]
        opt.close()
        return got, pub, strides, enc, eng.is_server

    res = run_ranks(emu, n, rank_main)
    # oracle in fp32 on the dtype-rounded gradients the ranks produced
    ref = [torch.nn.Parameter(p.detach().float().clone()) for p in make_params()]
    o = torch.optim.SGD([dict(params=[p for i, p in enumerate(ref) if groups_of[i] == g], **ghyper[g]) for g in range(3)], lr=0.1)
    for s in range(steps):
        per_rank = [grads_for(r, s, ref) for r in range(n)]
        for i, p in enumerate(ref):
            p.grad = None if i in skip[s] else sum(pr[i].float() for pr in per_rank)
        o.step()
    for got, pub, strides, enc, is_server in res:
        assert strides == (36, 1, 12, 4)                                   # channels-last physical order kept inside the arena
        for a, b in zip(pub, res[0][1]):
            assert torch.equal(a, b)
        if is_server:
            for i, (g, q) in enumerate(zip(got, ref)):
                assert torch.allclose(g, q.detach(), rtol=1e-5, atol=1e-6), (i, sizes[i] if i < 149 else "conv", float((g - q.detach()).abs().max()))
        assert len(enc) == steps                                           # one encode CALL per step; the binding splits it


def test_lr_scheduler_retained_grads_and_two_optimizers_per_rank(emu):
    """Everyday usage around the engine, 2 ranks: (a) a ``torch.optim.lr_scheduler`` drives the learning rate (sampled at the
    first chunk of every step), (b) ``zero_grad(set_to_none=False)`` keeps ``.grad`` tensors alive between steps, (c) TWO
    optimizers (two engines, two arenas, two flag pads) live in the same process and step alternately.  Each model must end on its
    own single-process oracle."""
    steps, n = 4, 2
    hyper = dict(lr=0.08, momentum=0.9, weight_decay=1e-3)

    def rank_main(rank, w):
        models_, opts, scheds = [], [], []
        for k in range(2):
            m = _model()
            with torch.no_grad():
                for p in m.parameters():
                    p.mul_(1.0 + 0.5 * k)                               # two different models
            o = ps.SGD(m.named_parameters(), m.parameters(), engine="host", mode="ps", **hyper)
            _attach(o)
            models_.append(m), opts.append(o), scheds.append(torch.optim.lr_scheduler.StepLR(o, step_size=2, gamma=0.5))
        for s in range(steps):
            for k in (0, 1):
                opts[k].zero_grad(set_to_none=False)
                _loss(models_[k], *_data(rank + 10 * k, s), skip_head=False).backward()
                opts[k].step()
                scheds[k].step()
        for o in opts:
            o._engine.check()
        w.barrier()
        out = [[p.detach().clone() for p in m.parameters()] for m in models_]
        for o in opts:
            o.close()
        return out

    res = run_ranks(emu, n, rank_main)
    for k in range(2):
        ref = _model()
        with torch.no_grad():
            for p in ref.parameters():
                p.mul_(1.0 + 0.5 * k)
        o = torch.optim.SGD(ref.parameters(), **hyper)
        sch = torch.optim.lr_scheduler.StepLR(o, step_size=2, gamma=0.5)
        for s in range(steps):
            tot = None
            for r in range(n):
                ref.zero_grad(set_to_none=True)
                _loss(ref, *_data(r + 10 * k, s), skip_head=False).backward()
                gs = [p.grad.clone() for p in ref.parameters()]
                tot = gs if tot is None else [a + b for a, b in zip(tot, gs)]
            for p, g in zip(ref.parameters(), tot):
                p.grad = g
            o.step()
            sch.step()
        for rank_out in res:
            for a, b in zip(rank_out[k], ref.parameters()):
                assert torch.allclose(a, b.detach(), rtol=3e-5, atol=3e-6), (k, float((a - b.detach()).abs().max()))


def test_mismatched_models_fail_loudly_on_every_rank(emu):
    """The kernels address peers' arenas by tile number, so ranks that built different models (or codings / modes) must not get as
    far as allocating symmetric memory: every rank raises, naming the first difference."""
    def rank_main(rank, w):
        model = _model()
        if rank == 1:
            model[4] = torch.nn.Linear(24, 11)                            # a different head on rank 1
        opt = ps.SGD(model.named_parameters(), model.parameters(), engine="host", mode="ps", lr=0.1)
        try:
            _attach(opt)
        except ValueError as exc:
            return str(exc)
        return None

    res = run_ranks(emu, 2, rank_main)
    assert all(r is not None and "rank 1 and rank 0 disagree" in r and "4.bias" in r and "(11,)" in r for r in res), res
