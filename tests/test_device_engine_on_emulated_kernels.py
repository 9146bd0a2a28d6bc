"""The WHOLE device-engine path on the CPU, single rank: ``ps.SGD`` / ``ps.Adam`` → ``DeviceEngine`` (layout, chunk pipeline, hyper
tuples, fp32 masters, per-parameter table, state views) → the REAL kernel source of ``ps_kernels.cu`` executed by the CPU emulator
(``tests/_cuda_emu.py``) over a fake arena — compared with ``torch.optim`` on the same gradients.  The GPU suite checks the same
thing on hardware (``test_gpu_engine.py``); this one keeps the glue between Python and the kernels honest in every CPU round."""
import contextlib
import ctypes

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200.parallel import device_engine as de
from tests import _cuda_emu
from tests.test_device_engine_control_flow import FakeArena, FakeEvent, FakeStream


@pytest.fixture(autouse=True)
def _single_threaded_torch():
    """The emulated kernels run on the calling thread; keeping torch single-threaded makes the timings of these tests repeatable."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)

DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _p(x):
    return ctypes.c_void_p(int(x))


class EmuPlan:
    def __init__(self, lib):
        self.lib = lib
        self.kind = self.wire = self.opt = self.grid = 0
        self.window_bytes = 128 << 20
        self.rank_ptrs = {}

    def set_rank_ptrs(self, r, wire_p, scales_p, param_p, signal_p):
        self.rank_ptrs[r] = (wire_p, scales_p, param_p, signal_p)

    def configure(self, world, rank, ntiles, bpt, cap, param_dt, bcast, reduce, param_mc, wire_mc, param_local, master, buf0, buf1,
                  buf2, tiles, signal_local, done_counter, stats):
        self.c = dict(world=world, rank=rank, ntiles=ntiles, bpt=bpt, cap=cap, param_dt=param_dt, bcast=bcast, param_local=param_local,
                      master=master, buf0=buf0, buf1=buf1, buf2=buf2, tiles=tiles, signal_local=signal_local,
                      done_counter=done_counter, stats=stats)
        assert reduce == 0 and world == 1

    def launch(self, epoch, groups, contrib_mask, inv_count, wait_grads, signal_mode, ack_mask=0, version=0, select_out=0,
               average_dynamic=0, active_ptr=0, timeout_s=30.0, wait_mask=0xffffffff, stream=0, tile_begin=0, tile_end=-1,
               wait_value=0, param_hyper=0):
        c = self.c
        if tile_end < 0:
            tile_begin, tile_end, wait_value = 0, c["ntiles"], epoch
        flat = [float(x) for g in groups for x in g]
        arr = lambda xs: (ctypes.c_void_p * len(xs))(*xs)      # noqa: E731
        w, s, p, sig = self.rank_ptrs[0]
        rc = self.lib.emu_update(self.kind, self.wire, self.opt, 1, 0, arr([w]), arr([s]), arr([p]), _p(c["param_local"]),
                                 _p(c["master"]), _p(c["buf0"]), _p(c["buf1"]), _p(c["buf2"]), _p(c["tiles"]), _p(active_ptr),
                                 _p(param_hyper), _p(c["signal_local"]), arr([sig]), _p(c["done_counter"]), _p(c["stats"]),
                                 (ctypes.c_float * len(flat))(*flat), len(groups), c["ntiles"], c["bpt"], c["cap"], c["param_dt"],
                                 c["bcast"], ctypes.c_uint32(contrib_mask), ctypes.c_uint32(0), ctypes.c_float(inv_count),
                                 ctypes.c_uint64(epoch), ctypes.c_uint64(wait_value), tile_begin, tile_end, wait_grads, signal_mode,
                                 ctypes.c_uint32(ack_mask), min(4, tile_end - tile_begin))
        assert rc == 0


class EmuM:
    TILE, SIGNAL_SLOTS, MAX_RANKS, MAX_GROUPS = 2048, 512, 16, 16
    SIG_GRAD_READY, SIG_PARAMS_READY, SIG_CONSUMED, SIG_ERROR, SIG_VERSION = 0, 64, 128, 200, 201
    SIG_ACK, SIG_GRAD_VERSION, SIG_STAGE_BEGIN, SIG_SEEN_VERSION = 256, 320, 202, 203

    def __init__(self, lib):
        self.lib = lib

    def UpdatePlan(self):
        return EmuPlan(self.lib)

    def update_max_grid(self, *a):
        return 444

    def encode(self, kind, wire, grads, first_tile, ntiles, param_idx, tiles_ptr, wire_ptr, scales_ptr, amax_ptr, residual_ptr,
               bpt, cap, ratio, sig_targets, sig_slot, sig_value, sig_counter, stream):
        n = len(grads)
        ia = lambda xs: (ctypes.c_int * n)(*xs)      # noqa: E731
        rc = self.lib.emu_encode(kind, wire, n, (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads]), ia(first_tile), ia(ntiles),
                                 ia(param_idx), _p(tiles_ptr), _p(wire_ptr), _p(scales_ptr), _p(amax_ptr), _p(residual_ptr), bpt,
                                 cap, ctypes.c_double(ratio), DT[grads[0].dtype], (ctypes.c_void_p * 1)(0), 0, 0,
                                 ctypes.c_uint64(0), _p(sig_counter))
        assert rc == 0

    def signal(self, *a):
        raise AssertionError("single rank: nothing to signal")

    def wait_flags(self, *a):
        raise AssertionError("single rank: nothing to wait for")

    def launch_count(self):
        return 0


@pytest.fixture
def emu(monkeypatch):
    lib = _cuda_emu.build()
    if lib is None:
        pytest.skip("no g++")
    m = EmuM(lib)
    monkeypatch.setattr(de.ext, "cuda", lambda: m)
    monkeypatch.setattr(de, "SymmetricArena", FakeArena)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setenv("PSB200_CHUNK_BYTES", str(2048 * 4 * 2))          # several chunks even for a tiny model
    return m


def _model(dtype):
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(40, 70), torch.nn.Tanh(), torch.nn.Linear(70, 55), torch.nn.Tanh(),
                               torch.nn.Linear(55, 10)).to(dtype)


def _attach_engine(opt):
    """What ``MPI_PS.__init__`` does for CUDA parameters (the ctor refuses the device engine for CPU tensors)."""
    for h in opt._hooks:
        h.remove()
    opt._engine = de.DeviceEngine(opt)
    from functools import partial
    opt._hooks = [p.register_hook(partial(opt._engine.on_grad, name=n, param=p)) for n, p in opt._named.items()]


def _train(model, opt, steps, dtype, skip_last_layer_until=0):
    for s in range(steps):
        g = torch.Generator().manual_seed(100 + s)
        x, y = torch.randn(16, 40, generator=g).to(dtype), torch.randint(0, 10, (16,), generator=g)
        opt.zero_grad(set_to_none=True)
        h = model[:-1](x)
        out = model[-1](h) if s >= skip_last_layer_until else h[:, :10]     # early steps: the head gets no gradient at all
        torch.nn.functional.cross_entropy(out.float(), y).backward()
        opt.step()


@pytest.mark.parametrize("optim,dtype", [("sgd", torch.float32), ("sgd", torch.bfloat16), ("adam", torch.float32)])
def test_engine_with_emulated_kernels_matches_torch(emu, optim, dtype):
    a, b = _model(dtype), _model(dtype)
    b.load_state_dict(a.state_dict())
    kw = dict(lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=True) if optim == "sgd" else dict(lr=1e-2, weight_decay=1e-2)
    cls = ps.SGD if optim == "sgd" else ps.Adam
    opt = cls(a.named_parameters(), a.parameters(), engine="host", **kw)
    _attach_engine(opt)
    eng = opt._engine
    assert eng.nchunks >= 2 and eng.size == 1
    if dtype == torch.float32:
        # SGD: torch.optim.SGD; Adam: the reference's formula (sqrt(v) + eps, bias correction folded into the step size,
        # ps.py:218-261) lives in the host engine — torch.optim.Adam places eps differently
        ref = torch.optim.SGD(b.parameters(), **kw) if optim == "sgd" else \
            ps.Adam(b.named_parameters(), b.parameters(), engine="host", **kw)
        _train(a, opt, 3, dtype)
        _train(b, ref, 3, dtype)
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p.detach() - q.detach()).abs().max())
    else:
        _train(a, opt, 3, dtype)
        assert eng.master is not None
        for s in eng.layout.slots:                       # published bf16 parameter == round(master); and it moved
            m = eng.master[s.offset:s.offset + s.numel]
            assert torch.equal(m.to(torch.bfloat16), s.param.data.reshape(-1))
        assert not torch.equal(a[0].weight, b[0].weight)
    sd = opt.state_dict()
    assert all(int(st["step"]) == 3 for st in sd["state"].values())
    opt.close()


def test_late_parameter_through_the_emulated_engine(emu):
    """A parameter whose first gradient arrives on step 3 (``ps.py:178-179,203-205``): skipped through the active mask while it has
    none, then started with buf = d_p via the per-parameter table — equal to the host engine, which keeps state per parameter."""
    a, b = _model(torch.float32), _model(torch.float32)
    b.load_state_dict(a.state_dict())
    kw = dict(lr=0.05, momentum=0.9, dampening=0.3)
    o1 = ps.SGD(a.named_parameters(), a.parameters(), engine="host", **kw)
    _attach_engine(o1)
    o2 = ps.SGD(b.named_parameters(), b.parameters(), engine="host", **kw)
    _train(a, o1, 5, torch.float32, skip_last_layer_until=2)
    _train(b, o2, 5, torch.float32, skip_last_layer_until=2)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p.detach() - q.detach()).abs().max())
    steps = sorted(int(st["step"]) for st in o1.state_dict()["state"].values())
    assert steps == [3, 3, 5, 5, 5, 5]
    o1.close(), o2.close()
