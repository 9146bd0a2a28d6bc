"""CUDA kernel numerics on ONE GPU: N *virtual ranks* (N wire arenas in one process) drive the
encode and fused PS kernels exactly as N GPUs would (the kernels only see pointers), and every
result is compared with the pure-PyTorch fp32 oracle in :mod:`pytorch_ps_mpi_b200.codings` +
``SGD.optim_step`` / ``Adam.optim_step`` (``/root/reference/ps.py:197-261``)."""
import math

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200.codings import TILE, KIND_SCALED, KIND_TOPK

pytestmark = pytest.mark.gpu
DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class Virtual:
    """N virtual ranks sharing one GPU."""

    def __init__(self, shapes, dtype, code, nranks, optim="sgd", groups=None, master=True):
        from pytorch_ps_mpi_b200.ops import ext
        from pytorch_ps_mpi_b200.parallel.layout import FlatLayout
        self.m = ext.cuda()
        dev = torch.device("cuda", 0)
        self.dev, self.n, self.dtype, self.code, self.optim = dev, nranks, dtype, code, optim
        self.params = [torch.nn.Parameter(torch.randn(s, device=dev).to(dtype)) for s in shapes]
        groups = groups or [list(range(len(shapes)))]
        self.pg = [{"params": [self.params[i] for i in g]} for g in groups]
        self.L = L = FlatLayout(self.pg, {id(p): f"p{i}" for i, p in enumerate(self.params)})
        spec = code.device_spec()
        self.spec, self.kind, self.wire = spec, spec.kind, spec.resolved_wire(dtype)
        self.bpt, self.cap = spec.bytes_per_tile(dtype), spec.tile_capacity()
        nt, npad = L.ntiles, L.numel_padded
        self.tiles = L.tile_table_fast().to(dev)
        assert torch.equal(self.tiles.cpu(), L.tile_table())
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=dev)
        self.wires = [z(nt * self.bpt, torch.uint8) for _ in range(nranks)]
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
        P = self.m.UpdatePlan()
        P.kind, P.wire, P.opt = self.kind, self.wire, 0 if optim == "sgd" else 1
        P.grid = min(nt, self.m.update_max_grid(0, 0, 0))
        for r in range(nranks):
            P.set_rank_ptrs(r, self.wires[r].data_ptr(), self.scales[r].data_ptr(), self.param_arenas[r].data_ptr(),
                            self.signals[r].data_ptr())
        P.configure(nranks, 0, nt, self.bpt, self.cap, DT[dtype], 1, 0, 0, 0, self.param_arenas[0].data_ptr(),
                    self.master.data_ptr() if self.master is not None else 0, self.buf0.data_ptr(),
                    self.buf1.data_ptr(), self.buf2.data_ptr(), self.tiles.data_ptr(), self.signals[0].data_ptr(),
                    self.counters.data_ptr(), self.counters.data_ptr() + 4)
        self.P = P

    def encode(self, r, grads):
        L = self.L
        if self.kind == KIND_SCALED:
            self.amax.zero_()
        order = [(L.by_id[id(p)], g) for p, g in zip(self.params, grads)]
        self.m.encode(self.kind, self.wire, [g.contiguous() for _, g in order], [s.first_tile for s, _ in order],
                      [s.ntiles for s, _ in order], [s.index for s, _ in order], self.tiles.data_ptr(),
                      self.wires[r].data_ptr(), self.scales[r].data_ptr(), self.amax.data_ptr(),
                      self.residuals[r].data_ptr() if self.residuals else 0, self.bpt, self.cap, float(self.spec.ratio))

    def update(self, epoch, hypers, wait=False, mask=None, inv=1.0):
        self.P.launch(epoch, hypers, (1 << self.n) - 1 if mask is None else mask, inv, 1 if wait else 0, 1,
                      timeout_s=5.0)
        torch.cuda.synchronize()

    def param_values(self, r=0):
        return [self.param_arenas[r][s.offset:s.offset + s.numel].view(s.param.shape)
                for s in (self.L.by_id[id(p)] for p in self.params)]


def sgd_h(lr=0.1, wd=0.0, mom=0.0, damp=0.0, nesterov=False, first=True):
    return [lr, wd, mom, damp, 0, 0, 0, 0, float(nesterov), 0, float(first)]


def adam_h(lr, b1, b2, eps, wd, t, amsgrad=False):
    return [lr, wd, 0, 0, b1, b2, eps, lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t), 0, float(amsgrad), float(t == 1)]


SHAPES = [(300, 41), (5000,), (64,), (3, 3, 16, 32), (TILE * 2,)]


def _oracle_sum(code_factory, grads_per_rank, shape_of):
    tot = None
    for gs in grads_per_rank:
        code = code_factory()
        dec = [code.decode(code.encode(g, name=str(i))).reshape(shape_of[i]).float() for i, g in enumerate(gs)]
        tot = dec if tot is None else [a + b for a, b in zip(tot, dec)]
    return tot


CODES = {
    "identity": lambda: ps.Identity(), "cast_bf16": lambda: ps.Cast("bf16"), "cast_fp16": lambda: ps.Cast("fp16"),
    "cast_e4m3": lambda: ps.Cast("fp8_e4m3"), "cast_e5m2": lambda: ps.Cast("fp8_e5m2"),
    "scale_i8": lambda: ps.Scale("int8"), "scale_e4m3": lambda: ps.Scale("fp8_e4m3"),
    "scale_f16": lambda: ps.Scale("fp16"),
    "topk_f32": lambda: ps.TopK(ratio=0.05), "topk_bf16": lambda: ps.TopK(ratio=0.3, values="bf16"),
}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cname", list(CODES))
@pytest.mark.parametrize("nranks", [1, 3, 8])
def test_encode_gather_sgd(dtype, cname, nranks):
    torch.manual_seed(0)
    V = Virtual(SHAPES, dtype, CODES[cname](), nranks, master=True)
    w0 = [p.data.float().clone() for p in V.params]
    grads = [[(torch.randn(s, device=V.dev) * (1 + r)).to(dtype) for s in SHAPES] for r in range(nranks)]
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
    assert int(V.signals[nranks - 1][64]) == 1                      # PARAMS_READY epoch raised on the last rank


def test_topk_wire_is_exact():
    """The block-wise top-k wire must select exactly the oracle's (index, value) set."""
    torch.manual_seed(1)
    code = ps.TopK(ratio=0.01)
    V = Virtual([(TILE * 3 + 77,)], torch.float32, code, 1)
    g = torch.randn(TILE * 3 + 77, device=V.dev)
    g[5] = g[9] = g[100] = 7.5            # ties → lower index first
    V.encode(0, [g])
    torch.cuda.synchronize()
    enc = code.encode(g)
    w = V.wires[0].view(torch.int32).view(-1, 2)           # (idx, f32 bits) entries
    cap = V.cap
    got_idx, got_val = [], []
    for t in range(V.L.ntiles):
        e = w[t * (V.bpt // 8): t * (V.bpt // 8) + cap]
        keep = e[:, 0] < TILE
        got_idx.append(e[keep, 0].long() + t * TILE)
        got_val.append(e[keep, 1].contiguous().view(torch.float32))
    got_idx, got_val = torch.cat(got_idx), torch.cat(got_val)
    assert torch.equal(got_idx.cpu(), enc["idx"].long().cpu())
    assert torch.equal(got_val.cpu(), enc["val"].cpu())


def test_topk_error_feedback():
    torch.manual_seed(2)
    code = ps.TopK(ratio=0.1, error_feedback=True)
    ref = ps.TopK(ratio=0.1, error_feedback=True)
    V = Virtual([(4000,)], torch.float32, code, 1)
    w = V.params[0].data.clone()
    for step in range(3):
        g = torch.randn(4000, device=V.dev)
        V.encode(0, [g])
        V.update(step + 1, [sgd_h(lr=1.0, first=step == 0)])
        w = w - ref.decode(ref.encode(g, name="p0")).reshape(-1)
        assert torch.allclose(V.param_values()[0], w, atol=1e-5)


@pytest.mark.parametrize("hyper", [dict(mom=0.9), dict(mom=0.9, nesterov=True, wd=1e-2), dict(mom=0.8, damp=0.3, wd=1e-3)])
def test_sgd_momentum_steps(hyper):
    torch.manual_seed(3)
    V = Virtual(SHAPES, torch.float32, ps.Identity(), 2)
    ref = [torch.nn.Parameter(p.data.clone()) for p in V.params]
    opt = torch.optim.SGD(ref, lr=0.1, momentum=hyper.get("mom", 0), dampening=hyper.get("damp", 0),
                          weight_decay=hyper.get("wd", 0), nesterov=hyper.get("nesterov", False))
    for step in range(4):
        grads = [[torch.randn(s, device=V.dev) for s in SHAPES] for _ in range(2)]
        for r in range(2):
            V.encode(r, grads[r])
        V.update(step + 1, [sgd_h(lr=0.1, first=step == 0, **hyper)])
        for p, a, b in zip(ref, grads[0], grads[1]):
            p.grad = a + b
        opt.step()
        for got, want in zip(V.param_values(), ref):
            assert torch.allclose(got, want.data, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("amsgrad", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_adam_steps(amsgrad, dtype):
    torch.manual_seed(4)
    V = Virtual(SHAPES, dtype, ps.Identity(), 2, optim="adam")
    ref = [torch.nn.Parameter(p.data.float().clone()) for p in V.params]
    o = ps.Adam([(f"p{i}", p) for i, p in enumerate(ref)], ref, lr=1e-2, betas=(0.9, 0.95), eps=1e-6,
                weight_decay=1e-2, amsgrad=amsgrad, engine="host")
    for step in range(4):
        grads = [[torch.randn(s, device=V.dev).to(dtype) for s in SHAPES] for _ in range(2)]
        for r in range(2):
            V.encode(r, grads[r])
        V.update(step + 1, [adam_h(1e-2, 0.9, 0.95, 1e-6, 1e-2, step + 1, amsgrad)])
        for p, a, b in zip(ref, grads[0], grads[1]):
            p.grad = torch.zeros_like(p)
            with torch.no_grad():
                o.optim_step(p, a.float() + b.float(), amsgrad=amsgrad, betas=(0.9, 0.95), weight_decay=1e-2,
                             eps=1e-6, lr=1e-2)
        master = [V.master[s.offset:s.offset + s.numel].view(s.param.shape) if V.master is not None else v
                  for s, v in zip((V.L.by_id[id(p)] for p in V.params), V.param_values())]
        for got, want in zip(master, ref):
            assert torch.allclose(got.float(), want.data, rtol=2e-5, atol=2e-6)
    o.close()


def test_groups_mask_and_average():
    torch.manual_seed(5)
    V = Virtual(SHAPES, torch.float32, ps.Identity(), 4, groups=[[0, 1], [2, 3, 4]])
    w0 = [p.data.clone() for p in V.params]
    grads = [[torch.randn(s, device=V.dev) for s in SHAPES] for _ in range(4)]
    for r in range(4):
        V.encode(r, grads[r])
    V.update(1, [sgd_h(lr=0.1), sgd_h(lr=1.0)], mask=0b1010, inv=0.5)      # only ranks 1 and 3, averaged
    for i, (got, w) in enumerate(zip(V.param_values(), w0)):
        lr = 0.1 if i < 2 else 1.0
        want = w - lr * 0.5 * (grads[1][i] + grads[3][i])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_flags_wait_signal_and_timeout():
    V = Virtual([(100,)], torch.float32, ps.Identity(), 2)
    m = V.m
    V.encode(0, [torch.ones(100, device=V.dev)])
    V.encode(1, [torch.ones(100, device=V.dev)])
    # both virtual ranks post GRAD_READY=1 into rank 0's pad, then the waiting kernel proceeds
    m.signal([V.signals[0].data_ptr()], m.SIG_GRAD_READY + 0, 1)
    m.signal([V.signals[0].data_ptr()], m.SIG_GRAD_READY + 1, 1)
    V.update(1, [sgd_h(lr=1.0)], wait=True)
    assert int(V.signals[0][m.SIG_ERROR]) == 0
    assert torch.allclose(V.param_values()[0], V.params[0].data - 2.0)
    # a flag that never arrives must time out (bounded spin), not hang
    m.wait_flags(V.signals[1].data_ptr(), m.SIG_PARAMS_READY, 1, 99, 0.2)
    torch.cuda.synchronize()
    assert int(V.signals[1][m.SIG_ERROR]) == 1
