"""CPU regression test of the device engine's CONTROL FLOW (``parallel/device_engine.py``) against fakes of the CUDA
extension, the symmetric arena and ``torch.cuda`` streams / events: what is launched, in which order, over which tiles, waiting
for which progress value, raising which flag.  (Numerics of the kernels are GPU tests; the flag protocol itself is
model-checked in ``test_protocol_model.py``.  This pins the glue between the two, and runs on every CPU round.)"""
import contextlib

import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200.parallel import device_engine as de

TILE = 2048


class FakeStream:
    cuda_stream = 4242

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def query(self):
        return True

    def synchronize(self):
        pass


class FakePlan:
    def __init__(self, log):
        self.log = log
        self.kind = self.wire = self.opt = self.grid = 0
        self.window_bytes = 128 << 20

    def set_rank_ptrs(self, *a):
        pass

    def configure(self, *a):
        self.cfg = a

    def launch(self, epoch, groups, contrib_mask, inv_count, wait_grads, signal_mode, ack_mask=0, version=0, select_out=0,
               average_dynamic=0, active_ptr=0, timeout_s=30.0, wait_mask=0xffffffff, stream=0, tile_begin=0, tile_end=-1,
               wait_value=0, param_hyper=0):
        self.log.append(("update", dict(epoch=epoch, groups=groups, contrib=contrib_mask, wait_grads=wait_grads,
                                        signal_mode=signal_mode, active=bool(active_ptr), lo=tile_begin, hi=tile_end,
                                        wait_value=wait_value, param_hyper=bool(param_hyper), stream=stream)))


class FakeM:
    TILE, SIGNAL_SLOTS, MAX_RANKS, MAX_GROUPS = 2048, 512, 16, 16
    SIG_GRAD_READY, SIG_PARAMS_READY, SIG_CONSUMED, SIG_ERROR, SIG_VERSION = 0, 64, 128, 200, 201
    SIG_ACK, SIG_GRAD_VERSION, SIG_STAGE_BEGIN, SIG_SEEN_VERSION = 256, 320, 202, 203

    def __init__(self):
        self.log = []

    def UpdatePlan(self):
        return FakePlan(self.log)

    def update_max_grid(self, *a):
        return 444

    def encode(self, kind, wire, grads, first_tile, ntiles, param_idx, tiles_ptr, wire_ptr, scales_ptr, amax_ptr, residual_ptr,
               bpt, cap, ratio, sig_targets, sig_slot, sig_value, sig_counter, stream):
        self.log.append(("encode", dict(first_tile=list(first_tile), sig_targets=list(sig_targets), sig_slot=sig_slot,
                                        sig_value=sig_value, stream=stream)))

    def signal(self, targets, slot, value, *a):
        self.log.append(("signal", dict(targets=list(targets), slot=slot, value=value)))

    def wait_flags(self, signal_local, slot0, mask, want, timeout_s, stream=0):
        self.log.append(("wait", dict(slot0=slot0, mask=mask, want=want, stream=stream)))

    def launch_count(self):
        return 0


class FakeArena:
    def __init__(self, nbytes, device, world):
        self.buf = torch.zeros(nbytes + 64, dtype=torch.uint8)
        self.nbytes, self.rank, self.size = nbytes, 0, 1
        self.ptrs = [self.buf.data_ptr()] * 16
        self.mc_ptr, self.provider = 0, "fake"

    local_ptr = property(lambda self: self.ptrs[0])
    has_multicast = property(lambda self: False)

    def tensor(self, offset, nbytes, dtype, rank=None):
        return self.buf[offset: offset + nbytes].view(dtype)

    def close(self):
        pass


@pytest.fixture
def fakes(monkeypatch):
    m = FakeM()
    monkeypatch.setattr(de.ext, "cuda", lambda: m)
    monkeypatch.setattr(de, "SymmetricArena", FakeArena)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return m


def _engine(monkeypatch, rank=0, size=1, mode="ps", chunk_bytes=TILE * 4, sizes=(3 * TILE, TILE, 2 * TILE + 5, 700), **env):
    monkeypatch.setenv("PSB200_CHUNK_BYTES", str(chunk_bytes))
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    params = [torch.nn.Parameter(torch.randn(n)) for n in sizes]
    named = [(f"p{i}", p) for i, p in enumerate(params)]
    opt = ps.SGD(named, params, lr=0.1, momentum=0.9, mode=mode, engine="host")
    opt.rank, opt.size = rank, size
    eng = de.DeviceEngine(opt)
    opt._engine = eng
    return opt, eng, params


def _fire_all(eng, skip=()):
    for sl in eng.layout.slots:                      # hooks fire in arena (= backward) order
        if sl.index not in skip:
            eng.on_grad(torch.randn_like(sl.param), sl.name, sl.param)


def test_single_rank_pipeline_covers_the_arena_once_per_step(fakes, monkeypatch):
    opt, eng, _ = _engine(monkeypatch)
    assert eng.nchunks >= 2 and eng.pipeline
    for step in (1, 2):
        del fakes.log[:]
        _fire_all(eng)
        launched_in_backward = [e for e in fakes.log if e[0] == "update"]
        assert len(launched_in_backward) == eng.nchunks                  # every chunk's update is queued DURING backward
        eng.step()
        ups = [e[1] for e in fakes.log if e[0] == "update"]
        assert [(u["lo"], u["hi"]) for u in ups] == eng.chunk_tiles and ups[0]["lo"] == 0 and ups[-1]["hi"] == eng.layout.ntiles
        assert [u["wait_value"] for u in ups] == [(step - 1) * eng.nchunks + k + 1 for k in range(eng.nchunks)]
        assert all(u["signal_mode"] == 0 and u["wait_grads"] == 0 and not u["active"] for u in ups)      # N = 1: nothing to raise
        assert not any(e[0] in ("wait", "signal") for e in fakes.log)
        kinds = [e[0] for e in fakes.log]
        assert kinds == ["encode", "update"] * eng.nchunks                # per chunk: its encode, then its update
        assert all(e[1]["stream"] == FakeStream.cuda_stream for e in fakes.log)   # everything on the comm stream
    assert eng._epoch == 2


def test_server_waits_with_one_warp_kernel_and_only_last_chunk_raises_params_ready(fakes, monkeypatch):
    opt, eng, _ = _engine(monkeypatch, rank=0, size=4)
    _fire_all(eng)
    eng.step()
    kinds = [e[0] for e in fakes.log]
    assert kinds == ["encode", "wait", "update"] * eng.nchunks
    enc = [e[1] for e in fakes.log if e[0] == "encode"]
    assert all(not e["sig_targets"] for e in enc)                         # the server's own gradient is ordered by the stream
    waits = [e[1] for e in fakes.log if e[0] == "wait"]
    ups = [e[1] for e in fakes.log if e[0] == "update"]
    assert all(w["mask"] == 0b1110 and w["slot0"] == FakeM.SIG_GRAD_READY for w in waits)
    assert [w["want"] for w in waits] == [u["wait_value"] for u in ups] == [k + 1 for k in range(eng.nchunks)]
    assert [u["signal_mode"] for u in ups] == [0] * (eng.nchunks - 1) + [1]
    assert all(u["contrib"] == 0b1111 and u["wait_grads"] == 0 for u in ups)


def test_worker_raises_progress_values_and_waits_for_params(fakes, monkeypatch):
    opt, eng, _ = _engine(monkeypatch, rank=2, size=4)
    assert not eng.is_server
    _fire_all(eng)
    eng.step()
    enc = [e[1] for e in fakes.log if e[0] == "encode"]
    assert [e["sig_value"] for e in enc] == [k + 1 for k in range(eng.nchunks)]
    assert all(e["sig_slot"] == FakeM.SIG_GRAD_READY + 2 and len(e["sig_targets"]) == 1 for e in enc)   # → the server only
    assert not any(e[0] == "update" for e in fakes.log)
    waits = [e[1] for e in fakes.log if e[0] == "wait"]
    assert len(waits) == 1 and waits[0]["slot0"] == FakeM.SIG_PARAMS_READY and waits[0]["want"] == 1
    # with a gated first GEMM registered, the worker queues no wait kernel at all
    del fakes.log[:]
    eng.register_gate(object())
    _fire_all(eng)
    eng.step()
    assert not any(e[0] == "wait" for e in fakes.log)
    flag_ptr, epoch = eng.gate()
    assert flag_ptr and epoch == 2


def test_parameter_without_gradient_holds_back_its_chunk_until_step(fakes, monkeypatch):
    opt, eng, _ = _engine(monkeypatch)
    missing = eng.chunks[1][0].index                                      # first parameter of the second chunk
    _fire_all(eng, skip={missing})
    early = [e[1] for e in fakes.log if e[0] == "update"]
    assert [(u["lo"], u["hi"]) for u in early] == eng.chunk_tiles[:1]     # only chunk 0 could go during backward
    eng.step()
    ups = [e[1] for e in fakes.log if e[0] == "update"]
    assert [(u["lo"], u["hi"]) for u in ups] == eng.chunk_tiles
    assert [u["active"] for u in ups] == [False] + [True] * (eng.nchunks - 1)        # the rest run under the active mask
    assert not eng._uniform_steps and eng._param_steps[missing] == 0
    # next step every parameter fires: per-parameter hyper table in use (its step count differs from its group's)
    del fakes.log[:]
    _fire_all(eng)
    eng.step()
    ups = [e[1] for e in fakes.log if e[0] == "update"]
    assert all(u["param_hyper"] and not u["active"] for u in ups)
    assert eng._param_steps[missing] == 1 and max(eng._param_steps) == 2


def test_unpipelined_mode_is_one_launch_in_step(fakes, monkeypatch):
    opt, eng, _ = _engine(monkeypatch, rank=0, size=2, PSB200_PIPELINE="0")
    assert eng.nchunks == 1 and not eng.pipeline
    _fire_all(eng)
    ups = [e[1] for e in fakes.log if e[0] == "update"]
    assert len(ups) == 1 and (ups[0]["lo"], ups[0]["hi"]) == (0, eng.layout.ntiles) and ups[0]["signal_mode"] == 1


def test_gradient_twice_before_step_is_rejected(fakes, monkeypatch):
    opt, eng, params = _engine(monkeypatch)
    sl = eng.layout.slots[-1]
    eng.on_grad(torch.randn_like(sl.param), sl.name, sl.param)
    with pytest.raises(RuntimeError):
        eng.on_grad(torch.randn_like(sl.param), sl.name, sl.param)


def test_async_server_runs_ahead_and_harvests_in_order(fakes, monkeypatch):
    """Device-resident AsySG-InCon server: select → update pairs are queued without looking at their result; results come back
    through the pinned ring in order; a 'nothing selected' iteration rolls the predicted version / step count back; ps_done once
    every worker posted DONE."""
    script = [(0b0010, 1, 0), (0b0100, 1, 0), (0b0110, 2, 0b0010), (0, 0, 0b0110), (0, 0, 0b0110)]   # (mask, count, finished)
    state = {"i": 0}

    def select_ready(signal_local, consumed, cand_mask, quota, out, timeout_s, version=0, begin_targets=(), stream=0):
        mask, cnt, fin = script[min(state["i"], len(script) - 1)]
        state["i"] += 1
        o = eng._select_out
        o.zero_()
        o[0], o[1], o[40], o[41] = mask, cnt, fin, version
        for r in range(8):
            if mask >> r & 1:
                o[44 + r] = r                        # pretend staleness = rank
        fakes.log.append(("select", dict(cand=cand_mask, quota=quota, version=version, begin=len(begin_targets))))

    fakes.select_ready = select_ready
    opt, eng, _ = _engine(monkeypatch, rank=0, size=3, mode="async")
    opt.quota = 2
    assert not eng.pipeline and eng.is_server
    n = opt.serve()
    assert n == 3                                    # three iterations applied something
    assert eng.version == 3 and eng._async_applied == 3
    assert eng._async_done_workers == {1, 2}
    last = opt.timings[-1]
    assert last["ps_done"] and last["updates_applied"] == 3 and last["contributors"] == [1, 2] and last["staleness"] == {1: 1, 2: 2}
    sel = [e[1] for e in fakes.log if e[0] == "select"]
    ups = [e[1] for e in fakes.log if e[0] == "update"]
    assert len(sel) == len(ups) >= 4                 # every select is followed by its update launch, no host decision between
    assert [s["version"] for s in sel[:3]] == [1, 2, 3]
    assert sel[0]["cand"] == 0b110 and all(s["begin"] == 0 for s in sel)     # inconsistent reads: no sequence lock to open


def test_async_worker_posts_gradient_with_its_parameter_version(fakes, monkeypatch):
    calls = []
    fakes.signal = lambda targets, slot, value, *a: calls.append((list(targets), slot, value, a))
    opt, eng, _ = _engine(monkeypatch, rank=1, size=3, mode="async")
    for step in (1, 2):
        del fakes.log[:], calls[:]
        _fire_all(eng)
        assert [e[0] for e in fakes.log if e[0] == "update"] == []           # workers never launch updates
        eng.step()
        (targets, slot, value, extra), = calls
        assert slot == FakeM.SIG_GRAD_READY + 1 and value == step and len(targets) == 1
        assert extra[-1] == FakeM.SIG_GRAD_VERSION + 1                        # staleness accounting: posts the version it read
        waits = [e[1] for e in fakes.log if e[0] == "wait"]
        assert (len(waits) == 1 and waits[0]["slot0"] == FakeM.SIG_ACK and waits[0]["want"] == step - 1) if step > 1 else not waits
