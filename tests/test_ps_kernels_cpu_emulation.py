"""The parameter-server CUDA kernels, compiled from the repository's ``.cu`` text and executed on the CPU (``tests/_cuda_emu.py``),
against the same fp32 PyTorch oracles as the GPU suite (``tests/test_gpu_kernels.py``): encode (cast / abs-max scale / radix-select
block-wise top-k with error feedback), the fused gather → decode → rank-ordered sum → SGD / Adam → publish kernel over N virtual ranks
(peer arenas are plain buffers here), chunked tile ranges, the active mask, per-parameter hyper table, completion flags, and the
async ``select`` kernel.  ``multimem`` paths are not emulated."""
import ctypes
import math
import os

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200.codings import KIND_SCALED, TILE
from pytorch_ps_mpi_b200.parallel.layout import FlatLayout
from tests import _cuda_emu


@pytest.fixture(autouse=True)
def _single_threaded_torch():
    """The emulated kernels run on the calling thread; keeping torch single-threaded makes the timings of these tests repeatable."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)

DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
SIG_PARAMS_READY, SIG_GRAD_READY, SIG_ERROR = 64, 0, 200
SIG_CONSUMED, SIG_VERSION, SIG_STAGE_BEGIN, SIG_SEEN_VERSION, SIG_ACK, SIG_GRAD_VERSION = 128, 201, 202, 203, 256, 320
c_void_pp = ctypes.POINTER(ctypes.c_void_p)


@pytest.fixture(scope="module")
def lib():
    lib = _cuda_emu.build()
    if lib is None:
        pytest.skip("no g++")
    return lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _arr(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class VirtualCPU:
    """N virtual ranks in one address space driving the emulated kernels exactly as the GPU harness drives the real ones."""

    def __init__(self, lib, shapes, dtype, code, nranks, optim="sgd", groups=None, master=True):
        self.lib, self.n, self.dtype, self.code, self.optim = lib, nranks, dtype, code, optim
        self.params = [torch.nn.Parameter(torch.randn(s).to(dtype)) for s in shapes]
        groups = groups or [list(range(len(shapes)))]
        self.pg = [{"params": [self.params[i] for i in g]} for g in groups]
        self.L = L = FlatLayout(self.pg, {id(p): f"p{i}" for i, p in enumerate(self.params)})
        spec = code.device_spec()
        self.spec, self.kind, self.wire = spec, spec.kind, spec.resolved_wire(dtype)
        self.bpt, self.cap = spec.bytes_per_tile(dtype), spec.tile_capacity()
        nt, npad = L.ntiles, L.numel_padded
        self.tiles = L.tile_table_fast().contiguous()
        z = lambda n, dt: torch.zeros(n, dtype=dt)        # noqa: E731
        self.wires = [z(nt * self.bpt + 64, torch.uint8) for _ in range(nranks)]
        self.scales = [z(L.nparams, torch.float32) for _ in range(nranks)]
        self.param_arenas = [z(npad, dtype) for _ in range(nranks)]
        self.signals = [z(512, torch.int64) for _ in range(nranks)]
        self.amax = z(L.nparams, torch.int32)
        self.residuals = [z(npad, torch.float32) for _ in range(nranks)] if spec.error_feedback else None
        for s in L.slots:
            for a in self.param_arenas:
                a[s.offset:s.offset + s.numel] = s.param.data.reshape(-1)
        self.master = self.param_arenas[0].float() if (dtype != torch.float32 and master) else None
        self.buf0, self.buf1, self.buf2 = z(npad, torch.float32), z(npad, torch.float32), z(npad, torch.float32)
        self.counters = z(8, torch.int32)

    def encode(self, r, grads, signal=None):
        L = self.L
        if self.kind == KIND_SCALED:
            self.amax.zero_()
        order = [(L.by_id[id(p)], g.contiguous()) for p, g in zip(self.params, grads)]
        keep = [g for _, g in order]
        n = len(order)
        ia = lambda xs: (ctypes.c_int * n)(*xs)      # noqa: E731
        sigp = _arr([signal[0].data_ptr()]) if signal else _arr([0])
        rc = self.lib.emu_encode(self.kind, self.wire, n, _arr([g.data_ptr() for g in keep]), ia([s.first_tile for s, _ in order]),
                                 ia([s.ntiles for s, _ in order]), ia([s.index for s, _ in order]), _ptr(self.tiles),
                                 _ptr(self.wires[r]), _ptr(self.scales[r]), _ptr(self.amax),
                                 _ptr(self.residuals[r] if self.residuals else None), self.bpt, self.cap,
                                 ctypes.c_double(float(self.spec.ratio)), DT[keep[0].dtype], sigp, 1 if signal else 0,
                                 signal[1] if signal else 0, ctypes.c_uint64(signal[2] if signal else 0),
                                 ctypes.c_void_p(self.counters.data_ptr() + 8))
        assert rc == 0

    def bind_multicast(self):
        """NVLS: one multicast window over the ranks' wire arenas, one over their parameter arenas (``symm_mem.cpp::mc_*``)."""
        self.lib.emu_mc_clear()
        self.mc_wire, self.mc_param = torch.zeros_like(self.wires[0]), torch.zeros_like(self.param_arenas[0])   # address space only
        for mc, bufs in ((self.mc_wire, self.wires), (self.mc_param, self.param_arenas)):
            assert self.lib.emu_mc_register(_ptr(mc), ctypes.c_size_t(mc.numel() * mc.element_size()), self.n,
                                            _arr([b.data_ptr() for b in bufs])) >= 0

    def update(self, epoch, hypers, wait=False, mask=None, inv=1.0, lo=0, hi=None, signal_mode=1, active=None, param_hyper=None,
               wait_value=None, grid=None, nvls=False):
        nt = self.L.ntiles
        hi = nt if hi is None else hi
        if nvls:          # bcast = multimem.st, reduce = multimem.ld_reduce
            self.lib.emu_update_mc(_ptr(self.mc_param), _ptr(self.mc_wire), 1)
        flat = [float(x) for h in hypers for x in h]
        hy = (ctypes.c_float * len(flat))(*flat)
        grid = grid or min(hi - lo, 6)                      # few CTAs: every CTA walks several tiles (grid-stride loop)
        rc = self.lib.emu_update(self.kind, self.wire, 0 if self.optim == "sgd" else 1, self.n, 0,
                                 _arr([w.data_ptr() for w in self.wires]), _arr([s.data_ptr() for s in self.scales]),
                                 _arr([a.data_ptr() for a in self.param_arenas]), _ptr(self.param_arenas[0]), _ptr(self.master),
                                 _ptr(self.buf0), _ptr(self.buf1), _ptr(self.buf2), _ptr(self.tiles), _ptr(active),
                                 _ptr(param_hyper), _ptr(self.signals[0]), _arr([s.data_ptr() for s in self.signals]),
                                 _ptr(self.counters), ctypes.c_void_p(self.counters.data_ptr() + 4), hy, len(hypers), nt, self.bpt,
                                 self.cap, DT[self.dtype], 2 if nvls else 1, ctypes.c_uint32((1 << self.n) - 1 if mask is None else mask),
                                 ctypes.c_uint32(((1 << self.n) - 1) & ~1), ctypes.c_float(inv), ctypes.c_uint64(epoch),
                                 ctypes.c_uint64(epoch if wait_value is None else wait_value), lo, hi, 1 if wait else 0,
                                 signal_mode, ctypes.c_uint32(0), grid)
        assert rc == 0

    def param_values(self, r=0):
        return [self.param_arenas[r][s.offset:s.offset + s.numel].view(s.param.shape)
                for s in (self.L.by_id[id(p)] for p in self.params)]


def sgd_h(lr=0.1, wd=0.0, mom=0.0, damp=0.0, nesterov=False, first=True):
    return [lr, wd, mom, damp, 0, 0, 0, 0, float(nesterov), 0, float(first)]


def adam_h(lr, b1, b2, eps, wd, t, amsgrad=False):
    return [lr, wd, 0, 0, b1, b2, eps, lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t), 0, float(amsgrad), float(t == 1)]


# the emulator runs a CTA's threads as fibers on one OS thread (~10 ms per 256-thread CTA): small arenas keep the whole
# coding x dtype x ranks matrix at a few seconds (PSB200_EMU_FULL=0 trims it to one case per coding)
SHAPES = [(60, 41), (2100,), (64,), (3, 3, 16, 16), (TILE + 9,)]
FULL = os.environ.get("PSB200_EMU_FULL", "1") == "1"
CODES = {
    "identity": lambda: ps.Identity(), "cast_bf16": lambda: ps.Cast("bf16"), "cast_fp16": lambda: ps.Cast("fp16"),
    "cast_e4m3": lambda: ps.Cast("fp8_e4m3"), "cast_e5m2": lambda: ps.Cast("fp8_e5m2"),
    "scale_i8": lambda: ps.Scale("int8"), "scale_e4m3": lambda: ps.Scale("fp8_e4m3"), "scale_f16": lambda: ps.Scale("fp16"),
    "topk_f32": lambda: ps.TopK(ratio=0.05), "topk_bf16": lambda: ps.TopK(ratio=0.3, values="bf16"),
}


def _oracle_sum(code_factory, grads_per_rank, shape_of):
    tot = None
    for gs in grads_per_rank:
        code = code_factory()
        dec = [code.decode(code.encode(g, name=str(i))).reshape(shape_of[i]).float() for i, g in enumerate(gs)]
        tot = dec if tot is None else [a + b for a, b in zip(tot, dec)]
    return tot


_MATRIX = ([(c, d, n) for c in CODES for d in (torch.float32, torch.bfloat16) for n in (1, 3)] if FULL else
           [("identity", torch.float32, 3), ("cast_bf16", torch.float32, 1), ("cast_e4m3", torch.bfloat16, 2),
            ("cast_fp16", torch.float32, 1), ("cast_e5m2", torch.float32, 1), ("scale_i8", torch.float32, 2),
            ("scale_e4m3", torch.bfloat16, 1), ("scale_f16", torch.float32, 1), ("topk_f32", torch.float32, 2),
            ("topk_bf16", torch.bfloat16, 1)])


@pytest.mark.parametrize("cname,dtype,nranks", _MATRIX)
def test_encode_gather_sgd(lib, cname, dtype, nranks):
    torch.manual_seed(0)
    V = VirtualCPU(lib, SHAPES, dtype, CODES[cname](), nranks)
    w0 = [p.data.float().clone() for p in V.params]
    grads = [[(torch.randn(s) * (1 + r)).to(dtype) for s in SHAPES] for r in range(nranks)]
    for r in range(nranks):
        V.encode(r, grads[r])
    V.update(1, [sgd_h(lr=0.5)])
    want_sum = _oracle_sum(CODES[cname], grads, SHAPES)
    for r in range(nranks):                               # unicast publication reached every rank
        for got, w, g in zip(V.param_values(r), w0, want_sum):
            want = (w - 0.5 * g).to(dtype)
            tol = 1e-5 if dtype == torch.float32 else 1e-2
            assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol), (cname, (got.float() - want.float()).abs().max())
    assert int(V.counters[0]) == 0 and int(V.counters[1]) == 1      # completion counter reset, one signal
    assert int(V.signals[nranks - 1][SIG_PARAMS_READY]) == 1        # PARAMS_READY epoch raised on the last rank


def test_topk_wire_is_exact(lib):
    """The block-wise top-k wire must select exactly the oracle's (index, value) set (4-pass radix select, ties → lower index)."""
    torch.manual_seed(1)
    code = ps.TopK(ratio=0.01)
    V = VirtualCPU(lib, [(TILE * 3 + 77,)], torch.float32, code, 1)
    g = torch.randn(TILE * 3 + 77)
    g[5] = g[9] = g[100] = 7.5
    V.encode(0, [g])
    enc = code.encode(g)
    w = V.wires[0][: V.L.ntiles * V.bpt].view(torch.int32).view(-1, 2)
    got_idx, got_val = [], []
    for t in range(V.L.ntiles):
        e = w[t * (V.bpt // 8): t * (V.bpt // 8) + V.cap]
        keep = e[:, 0] < TILE
        got_idx.append(e[keep, 0].long() + t * TILE)
        got_val.append(e[keep, 1].contiguous().view(torch.float32))
    assert torch.equal(torch.cat(got_idx), enc["idx"].long())
    assert torch.equal(torch.cat(got_val), enc["val"])


def test_topk_error_feedback(lib):
    torch.manual_seed(2)
    code, ref = ps.TopK(ratio=0.1, error_feedback=True), ps.TopK(ratio=0.1, error_feedback=True)
    V = VirtualCPU(lib, [(4000,)], torch.float32, code, 1)
    w = V.params[0].data.clone()
    for step in range(3):
        g = torch.randn(4000)
        V.encode(0, [g])
        V.update(step + 1, [sgd_h(lr=1.0, first=step == 0)])
        w = w - ref.decode(ref.encode(g, name="p0")).reshape(-1)
        assert torch.allclose(V.param_values()[0], w, atol=1e-5)


@pytest.mark.parametrize("hyper", [dict(mom=0.9, nesterov=True, wd=1e-2), dict(mom=0.8, damp=0.3, wd=1e-3)])
def test_sgd_momentum_steps(lib, hyper):
    torch.manual_seed(3)
    V = VirtualCPU(lib, SHAPES, torch.float32, ps.Identity(), 2)
    ref = [torch.nn.Parameter(p.data.clone()) for p in V.params]
    opt = torch.optim.SGD(ref, lr=0.1, momentum=hyper.get("mom", 0), dampening=hyper.get("damp", 0),
                          weight_decay=hyper.get("wd", 0), nesterov=hyper.get("nesterov", False))
    for step in range(2 if not FULL else 4):
        grads = [[torch.randn(s) for s in SHAPES] for _ in range(2)]
        for r in range(2):
            V.encode(r, grads[r])
        V.update(step + 1, [sgd_h(lr=0.1, first=step == 0, **hyper)])
        for p, a, b in zip(ref, grads[0], grads[1]):
            p.grad = a + b
        opt.step()
        for got, want in zip(V.param_values(), ref):
            assert torch.allclose(got, want.data, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("amsgrad,dtype", [(True, torch.float32), (False, torch.bfloat16)] +
                         ([(False, torch.float32), (True, torch.bfloat16)] if FULL else []))
def test_adam_steps(lib, amsgrad, dtype):
    torch.manual_seed(4)
    V = VirtualCPU(lib, SHAPES, dtype, ps.Identity(), 2, optim="adam")
    ref = [torch.nn.Parameter(p.data.float().clone()) for p in V.params]
    o = ps.Adam([(f"p{i}", p) for i, p in enumerate(ref)], ref, lr=1e-2, betas=(0.9, 0.95), eps=1e-6,
                weight_decay=1e-2, amsgrad=amsgrad, engine="host")
    for step in range(2 if not FULL else 4):
        grads = [[torch.randn(s).to(dtype) for s in SHAPES] for _ in range(2)]
        for r in range(2):
            V.encode(r, grads[r])
        V.update(step + 1, [adam_h(1e-2, 0.9, 0.95, 1e-6, 1e-2, step + 1, amsgrad)])
        for p, a, b in zip(ref, grads[0], grads[1]):
            with torch.no_grad():
                o.optim_step(p, a.float() + b.float(), amsgrad=amsgrad, betas=(0.9, 0.95), weight_decay=1e-2, eps=1e-6, lr=1e-2)
        master = [V.master[s.offset:s.offset + s.numel].view(s.param.shape) if V.master is not None else v
                  for s, v in zip((V.L.by_id[id(p)] for p in V.params), V.param_values())]
        for got, want in zip(master, ref):
            assert torch.allclose(got.float(), want.data, rtol=2e-5, atol=2e-6)
    o.close()


def test_groups_mask_and_average(lib):
    torch.manual_seed(5)
    V = VirtualCPU(lib, SHAPES, torch.float32, ps.Identity(), 4, groups=[[0, 1], [2, 3, 4]])
    w0 = [p.data.clone() for p in V.params]
    grads = [[torch.randn(s) for s in SHAPES] for _ in range(4)]
    for r in range(4):
        V.encode(r, grads[r])
    V.update(1, [sgd_h(lr=0.1), sgd_h(lr=1.0)], mask=0b1010, inv=0.5)      # only ranks 1 and 3, averaged
    for i, (got, w) in enumerate(zip(V.param_values(), w0)):
        lr = 0.1 if i < 2 else 1.0
        assert torch.allclose(got, w - lr * 0.5 * (grads[1][i] + grads[3][i]), rtol=1e-5, atol=1e-6)


def test_chunked_launches_active_mask_and_per_parameter_table(lib):
    """The round-2 launch shape: one launch per pipeline chunk (tile ranges), only the last raises PARAMS_READY; tiles of a
    parameter without a gradient are skipped through the active mask; a per-parameter {step_size, first_step} table overrides
    the group's (a late parameter starts its momentum with buf = d_p whatever the group's first_step says)."""
    torch.manual_seed(6)
    V = VirtualCPU(lib, SHAPES, torch.float32, ps.Identity(), 2)
    L = V.L
    w0 = [p.data.clone() for p in V.params]
    grads = [[torch.randn(s) for s in SHAPES] for _ in range(2)]
    for r in range(2):
        V.encode(r, grads[r])
    inactive = 1                                              # parameter index 1 (in registration order) got no gradient
    active = torch.ones(L.nparams, dtype=torch.uint8)
    active[L.by_id[id(V.params[inactive])].index] = 0
    table = torch.zeros(L.nparams, 2)
    table[:, 1] = 1.0                                         # per-parameter first_step = 1 although the group says 0
    cut = L.slots[2].first_tile                               # two chunks: arena tiles [0, cut) and [cut, ntiles)
    V.update(1, [sgd_h(lr=0.1, mom=0.9, damp=0.5, first=False)], lo=0, hi=cut, signal_mode=0, active=active, param_hyper=table)
    assert int(V.signals[1][SIG_PARAMS_READY]) == 0 and int(V.counters[1]) == 0      # a non-final chunk raises nothing
    V.update(1, [sgd_h(lr=0.1, mom=0.9, damp=0.5, first=False)], lo=cut, hi=L.ntiles, signal_mode=1, active=active, param_hyper=table)
    assert int(V.signals[1][SIG_PARAMS_READY]) == 1
    for i, (got, w) in enumerate(zip(V.param_values(), w0)):
        want = w if i == inactive else w - 0.1 * (grads[0][i] + grads[1][i])        # first step: buf = d_p, dampening ignored
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), i


def test_fused_flag_raise_and_wait(lib):
    """The last CTA of an encode launch publishes GRAD_READY; the update kernel's own wait (unpipelined path) sees it."""
    V = VirtualCPU(lib, [(TILE * 3,)], torch.float32, ps.Identity(), 2)
    w = V.params[0].data.clone()
    V.encode(0, [torch.ones(TILE * 3)])
    V.encode(1, [torch.ones(TILE * 3)], signal=(V.signals[0], SIG_GRAD_READY + 1, 7))
    assert int(V.signals[0][SIG_GRAD_READY + 1]) == 7 and int(V.counters[2]) == 0      # flag up, CTA counter back to zero
    V.update(7, [sgd_h(lr=1.0)], wait=True, wait_value=7)
    assert int(V.signals[0][SIG_ERROR]) == 0
    assert torch.allclose(V.param_values()[0], w - 2.0)


def _select(lib, sig, consumed, cand, quota, out, version, begin=(), timeout_s=0.05):
    arr = (ctypes.c_void_p * max(len(begin), 1))(*[b.data_ptr() for b in begin])
    lib.emu_select(_ptr(sig), _ptr(consumed), ctypes.c_uint32(cand), quota, _ptr(out), ctypes.c_uint64(version), arr, len(begin),
                   ctypes.c_double(timeout_s))


def test_async_select_kernel(lib):
    """``psb_select_kernel``: quota gradients from ANY source with rotating priority, finished workers reported, staleness
    recorded; with ``consistent=True`` targets it opens the sequence lock (STAGE_BEGIN = version) on every rank iff it selected."""
    sig = torch.zeros(512, dtype=torch.int64)
    peer = torch.zeros(512, dtype=torch.int64)
    consumed = torch.zeros(64, dtype=torch.int64)
    out = torch.zeros(64, dtype=torch.int64)
    sig[SIG_GRAD_READY + 1], sig[SIG_GRAD_READY + 2], sig[SIG_GRAD_READY + 3] = 1, 1, 1 << 62      # rank 3 posted DONE
    sig[320 + 1], sig[320 + 2] = 4, 2                          # SIG_GRAD_VERSION: the versions their gradients were computed on
    _select(lib, sig, consumed, 0b1110, 1, out, 6, begin=(sig, peer))
    first = int(out[0])
    assert first in (0b0010, 0b0100) and int(out[1]) == 1 and int(out[40]) == 0b1000 and int(out[41]) == 6
    assert int(sig[SIG_STAGE_BEGIN]) == 6 and int(peer[SIG_STAGE_BEGIN]) == 6
    r = 1 if first == 0b0010 else 2
    assert int(out[44 + r]) == 5 - int(sig[320 + r]) and int(consumed[r]) == 1
    _select(lib, sig, consumed, 0b1110, 1, out, 7)
    assert int(out[0]) == (0b0110 ^ first)                    # rotating priority: the other worker is served next
    _select(lib, sig, consumed, 0b1110, 2, out, 8, begin=(sig, peer))
    assert int(out[0]) == 0 and int(out[1]) == 0              # nothing new within the (emulated, short) time-out → no selection
    assert int(peer[SIG_STAGE_BEGIN]) == 6                    # ... and the sequence lock stays closed


def test_signal_and_wait_kernels(lib):
    """``psb_signal_kernel`` (flag + optional extra slot + the staleness version hand-off) and ``psb_wait_kernel`` (masked wait;
    a time-out poisons SIG_ERROR and makes every later wait return at once — what ``DeviceEngine._poll_error`` surfaces)."""
    import time
    mine, a, b = (torch.zeros(512, dtype=torch.int64) for _ in range(3))
    mine[SIG_VERSION], mine[SIG_SEEN_VERSION] = 9, 7          # latest published version / the one my last forward started on
    tg = (ctypes.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    lib.emu_signal(tg, 2, SIG_GRAD_READY + 3, ctypes.c_uint64(5), _ptr(a), SIG_ACK + 3, ctypes.c_uint64(11), _ptr(mine),
                   SIG_GRAD_VERSION + 3)
    for t in (a, b):
        assert int(t[SIG_GRAD_READY + 3]) == 5 and int(t[SIG_ACK + 3]) == 11 and int(t[SIG_GRAD_VERSION + 3]) == 7
    assert int(mine[SIG_SEEN_VERSION]) == 9                   # re-sampled for the NEXT gradient
    # wait: ranks 1 and 3 of the mask must reach 5; rank 3 has, rank 1 has not
    t0 = time.time()
    lib.emu_wait(_ptr(a), SIG_GRAD_READY, ctypes.c_uint32(0b1000), ctypes.c_uint64(5), ctypes.c_double(5.0))
    assert time.time() - t0 < 1.0 and int(a[SIG_ERROR]) == 0
    lib.emu_wait(_ptr(a), SIG_GRAD_READY, ctypes.c_uint32(0b1010), ctypes.c_uint64(5), ctypes.c_double(0.05))
    assert int(a[SIG_ERROR]) != 0                             # timed out on rank 1's flag
    t0 = time.time()
    lib.emu_wait(_ptr(a), SIG_GRAD_READY, ctypes.c_uint32(0b0010), ctypes.c_uint64(5), ctypes.c_double(30.0))
    assert time.time() - t0 < 1.0                             # poisoned: no second 30 s spin


def test_snapshot_kernels_sequence_lock(lib):
    """``psb_snapshot_fetch`` / ``psb_snapshot_commit`` (``consistent=True``): adopt staging → live only for a COMPLETE version
    (BEGIN == VERSION), newer than the adopted one; anything else leaves the live parameters untouched."""
    n = 256 * 40 + 16                                          # several CTAs, a ragged tail (16-byte vectors)
    sig = torch.zeros(512, dtype=torch.int64)
    stage = torch.arange(n, dtype=torch.float32)
    shadow, live = torch.zeros(n), torch.full((n,), -1.0)
    scratch = torch.tensor([0, -1, 0, 0, 0, 0], dtype=torch.int64)

    def snap(attempts=1):
        lib.emu_snapshot(_ptr(sig), _ptr(stage), _ptr(shadow), _ptr(live), ctypes.c_size_t(n * 4), _ptr(scratch), attempts)

    snap()                                                     # version 0: nothing published yet
    assert float(live[0]) == -1.0 and int(scratch[5]) == 0
    sig[SIG_STAGE_BEGIN], sig[SIG_VERSION] = 3, 2              # publication of version 3 in progress
    snap()
    assert float(live[5]) == -1.0 and int(scratch[5]) == 0 and int(scratch[4]) == 0
    assert scratch[:4].tolist() == [0, -1, 0, 0]               # the last CTA re-armed the voting words
    sig[SIG_VERSION] = 3                                       # complete
    snap(attempts=2)                                           # second attempt: already adopted → vetoed, live stays on v3
    assert torch.equal(live, stage) and int(scratch[5]) == 3
    stage += 1.0
    snap()                                                     # same version again: not re-copied
    assert float(live[1]) == 1.0
    sig[SIG_STAGE_BEGIN], sig[SIG_VERSION] = 4, 4
    snap()
    assert torch.equal(live, stage) and int(scratch[5]) == 4


@pytest.mark.parametrize("dtype,optim,ranks", [(torch.float32, "sgd", 4), (torch.bfloat16, "sgd", 5), (torch.float16, "adam", 4),
                                                (torch.bfloat16, "adam", 8)])
def test_nvls_reduce_and_multicast_publish(lib, dtype, optim, ranks):
    """The switch path (default at N >= 4): ``multimem.ld_reduce`` over the wire arenas (NVLS_U tiles in flight per thread, chunk
    ranges, inactive parameters) and ``multimem.st`` publication — against the fp32 sum of the ranks' wire values (16-bit wires:
    fp32 accumulation rounded ONCE to the wire type, as the switch does) and against the P2P path on the same inputs."""
    shapes = [(TILE * 9 + 5,), (60, 41), (TILE,), (300,)]          # 13 tiles: more than NVLS_U x grid for small grids
    torch.manual_seed(3)
    V = VirtualCPU(lib, shapes, dtype, ps.Identity(), ranks, optim=optim)
    torch.manual_seed(3)
    P = VirtualCPU(lib, shapes, dtype, ps.Identity(), ranks, optim=optim)
    V.bind_multicast()
    grads = [[torch.randn(s).to(dtype) for s in shapes] for _ in range(ranks)]
    for r in range(ranks):
        V.encode(r, grads[r])
        P.encode(r, grads[r])
    h = [sgd_h(lr=0.1, mom=0.9)] if optim == "sgd" else [adam_h(1e-2, 0.9, 0.999, 1e-8, 0.0, 1)]
    active = torch.ones(len(shapes), dtype=torch.uint8)
    active[V.L.by_id[id(V.params[2])].index] = 0                   # one parameter (registration index 2) sits the step out
    nt = V.L.ntiles
    V.update(1, h, lo=0, hi=5, signal_mode=0, active=active, nvls=True, grid=2)      # two chunks, ragged against NVLS_U * grid
    V.update(1, h, lo=5, hi=nt, signal_mode=1, active=active, nvls=True, grid=3)
    P.update(1, h, active=active)
    assert int(V.signals[0][SIG_ERROR]) == 0 and all(int(s[SIG_PARAMS_READY]) == 1 for s in V.signals)
    for r in range(1, ranks):
        assert torch.equal(V.param_arenas[r], V.param_arenas[0])   # multimem.st reached every rank
    for i, (a, b) in enumerate(zip(V.param_values(), P.param_values())):
        if i == 2:
            assert torch.equal(a, V.params[2].data)                # inactive: untouched
            continue
        # fp32 wire: only the summation order differs; 16-bit wires: the switch rounds the fp32 sum once, P2P never rounds it
        tol = 1e-5 if dtype == torch.float32 else 0.1 * 2.0 ** -7 * float(ranks) ** 0.5 * 4
        assert torch.allclose(a.float(), b.float(), rtol=2e-2 if dtype != torch.float32 else 1e-5, atol=tol), \
            (i, float((a.float() - b.float()).abs().max()))
    if optim == "sgd":                                             # exact model of the switch: sum in fp32, round once, then SGD
        tot = [sum(g[i].float() for g in grads) for i in range(len(shapes))]
        if dtype != torch.float32:
            tot = [t.to(dtype).float() for t in tot]
        for i, (a, t, p) in enumerate(zip(V.param_values(), tot, V.params)):
            if i != 2:
                want = (p.data.float() - 0.1 * t)
                sl = V.L.by_id[id(p)]
                got = V.master[sl.offset: sl.offset + sl.numel].view(p.shape) if V.master is not None else a.float()
                assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (i, float((got - want).abs().max()))
    lib.emu_mc_clear()
