"""L2 wire format + helpers: ``serialization`` (the finished form of the reference's
``serialization.py:8-50``), ``compress``/``decompress`` (``mpi_comms.py:18-30``), ``to_np``/``to_torch``
(``mpi_comms.py:32-58``), ``trim_msg``, ``print_summary``, ``_bytes_of`` — and the coding oracles."""
import numpy as np
import pytest
import torch

import pytorch_ps_mpi_b200 as ps
from pytorch_ps_mpi_b200 import serialization as ser
from pytorch_ps_mpi_b200.codings import TILE, tile_k

comms = ps.comms


def _obj():
    n = 1000
    x = torch.linspace(0, 6.28, n)
    return {"x": x, "y": (torch.sin(x) + 0.25).bfloat16(), "n": n, "fp8": torch.randn(33).to(torch.float8_e4m3fn),
            "nested": [torch.arange(5), {"np": np.arange(6, dtype=np.float64).reshape(2, 3)}], "s": "str", "t": (1, 2.5)}


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a.view(torch.uint8) if a.dtype.itemsize == 1 else a, b.view(torch.uint8) if b.dtype.itemsize == 1 else b)
    if isinstance(a, np.ndarray):
        return a.dtype == b.dtype and np.array_equal(a, b)
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


@pytest.mark.parametrize("level", [0, 1, 5])
def test_dumps_loads_roundtrip_all_dtypes(level):
    obj = _obj()
    raw = ser.dumps(obj)
    msg = ser.compress(raw, level=level)
    back = ser.loads(ser.decompress(msg))
    assert _same(obj, back)
    assert back["y"].dtype == torch.bfloat16            # the reference cast everything to float32 (mpi_comms.py:48)
    if level == 0:
        assert len(msg) == len(raw) + ser.HEADER_BYTES  # level 0 = framing only (blosc clevel-0 analogue)
    else:
        assert len(msg) < len(raw)


def test_compress_api_parity():
    with pytest.raises(ValueError):
        ser.compress(b"abc", name="lz4")                 # the reference refuses lz4 / snappy (mpi_comms.py:22-24)
    with pytest.raises(ValueError):
        ser.compress(b"abc", name="snappy")
    assert bytes(ser.decompress(ser.compress(b"hello world" * 100, level=2, name="zlib"))) == b"hello world" * 100


def test_unframe_detects_truncation_and_garbage():
    msg = ser.compress(ser.dumps({"a": torch.arange(100)}))
    with pytest.raises(ValueError):
        ser.decompress(msg[: len(msg) // 2])             # the reference's sentinel scan becomes an explicit length check
    with pytest.raises(ValueError):
        ser.decompress(b"\x00" * 64)


def test_native_and_python_paths_agree(monkeypatch):
    obj = _obj()
    a = bytes(ser.dumps(obj))
    monkeypatch.setattr(ser, "_native", lambda: None)
    b = bytes(ser.dumps(obj))
    assert a == b
    assert _same(ser.loads(bytearray(b)), obj)
    raw = bytes(range(256)) * 5
    assert bytes(ser._unshuffle(ser._shuffle(raw, 4), 4)) == raw


def test_to_np_to_torch():
    d = {"a": torch.ones(3, dtype=torch.float64), "b": [torch.zeros(2, dtype=torch.int32)], "c": 5,
         "h": torch.ones(2).bfloat16()}
    n = comms.to_np(d)
    assert isinstance(n["a"], np.ndarray) and n["a"].dtype == np.float64 and isinstance(n["b"][0], np.ndarray) and n["c"] == 5
    assert n["h"].dtype == np.float32                    # numpy has no bf16: exact upcast
    t = comms.to_torch(n)
    assert t["a"].dtype == torch.float64 and t["b"][0].dtype == torch.int32        # dtype preserving


def test_trim_msg():
    payload = bytearray(b"abcdef")
    assert comms.trim_msg(payload + bytearray(b"\x29" * 32) + bytearray(10)) == payload
    with pytest.raises(Exception):
        comms.trim_msg(bytearray(b"no sentinel here"))


def test_print_summary(capsys):
    s = comms.print_summary({"w": torch.zeros(2, 3), "a": np.zeros(4), "lr": 0.1})
    out = capsys.readouterr().out
    assert "(2, 3)" in out and "(4,)" in out and "lr: 0.1" in out and s.strip().startswith("{")


def test_bytes_of():
    t = torch.zeros(4, 5)                                 # 2-D: the reference's comment admits it mis-sized these
    assert ps._bytes_of(t) == 80
    assert ps._bytes_of({"a": t, "b": [np.zeros(3), t]}) == 80 + 24 + 80
    p = torch.nn.Parameter(torch.zeros(10))
    p.grad = torch.zeros(10)
    assert ps._bytes_of(p) == 80


# ---- coding oracles ----------------------------------------------------------------------------------
def test_identity_cast_scale_roundtrip():
    g = torch.randn(1000) * 3
    assert torch.equal(ps.Identity().decode(ps.Identity().encode(g)), g)
    c = ps.Cast("bf16")
    assert torch.equal(c.decode(c.encode(g)), g.bfloat16().float())
    for dt, tol in (("int8", 1 / 127), ("fp8_e4m3", 1 / 8), ("fp16", 1e-3)):
        s = ps.Scale(dt)
        back = s.decode(s.encode(g))
        assert (back - g).abs().max() <= tol * g.abs().max() + 1e-6
    big = torch.tensor([1e6, -1e6, 3.0])
    assert torch.isfinite(ps.Cast("fp8_e4m3").decode(ps.Cast("fp8_e4m3").encode(big))).all()     # saturating, not NaN


def test_topk_blockwise_semantics():
    torch.manual_seed(0)
    n = TILE * 2 + 100
    g = torch.randn(n)
    t = ps.TopK(ratio=0.1)
    code = t.encode(g)
    want = tile_k(0.1, TILE) * 2 + tile_k(0.1, 100)
    assert code["idx"].numel() == want
    dense = t.decode(code)
    nz = dense != 0
    assert nz.sum() == want and torch.equal(dense[nz], g[nz])
    for b in range(3):                                   # every kept entry beats every dropped one inside its block
        lo, hi = b * TILE, min((b + 1) * TILE, n)
        blk, keep = g[lo:hi].abs(), nz[lo:hi]
        assert blk[keep].min() >= blk[~keep].max()
    e = ps.TopK(k=7, exact=True)
    assert e.encode(g)["idx"].numel() == 7 and e.device_spec() is None
    with pytest.raises(ValueError):
        ps.TopK(ratio=0.1, k=3)


def test_topk_error_feedback_accumulates():
    t = ps.TopK(ratio=0.01, error_feedback=True)
    g = torch.ones(1000)
    total = torch.zeros(1000)
    for _ in range(100):
        total += t.decode(t.encode(g, name="w"))
    assert torch.allclose(total + t._residual["w"], torch.full((1000,), 100.0))


def test_qsgd_is_unbiased_and_bounded():
    torch.manual_seed(0)
    g = torch.randn(2000)
    c = ps.QSGD(levels=15, seed=1)
    dec = torch.stack([c.decode(c.encode(g)) for _ in range(400)])
    assert dec.shape == (400, 2000)
    assert (dec.mean(0) - g).abs().mean() < 0.12 * g.abs().mean()          # unbiased: the mean converges to g
    step = float(g.norm()) / 15
    assert (dec[0] - g).abs().max() <= step + 1e-5                           # one quantisation step at most
    assert c.encode(g)["q"].dtype == torch.int8 and ps.QSGD(levels=255).encode(g)["q"].dtype == torch.int16
    z = ps.QSGD().decode(ps.QSGD().encode(torch.zeros(5)))
    assert torch.equal(z, torch.zeros(5)) and c.device_spec() is None       # host path only


def test_svd_coding_low_rank_exact_and_passthrough():
    torch.manual_seed(0)
    a, b = torch.randn(40, 3), torch.randn(3, 7 * 5)
    g = (a @ b).reshape(40, 7, 5)                                           # exactly rank 3
    c = ps.SVD(rank=3)
    code = c.encode(g)
    assert code["us"].shape == (40, 3) and code["vh"].shape == (3, 35)
    assert torch.allclose(c.decode(code), g, atol=1e-4)
    assert (ps.SVD(rank=1).decode(ps.SVD(rank=1).encode(g)) - g).norm() < g.norm()     # best rank-1 approx
    v = torch.randn(11)
    assert torch.equal(c.decode(c.encode(v)), v)                            # 1-D gradients are sent dense


# ---- property tests ------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st   # noqa: E402


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 3 * TILE), st.floats(0.001, 1.0), st.integers(0, 2 ** 31 - 1))
def test_topk_blockwise_properties(n, ratio, seed):
    g = torch.randn(n, generator=torch.Generator().manual_seed(seed))
    code = ps.TopK(ratio=ratio).encode(g)
    idx = code["idx"].long()
    assert torch.equal(idx, torch.sort(idx).values) and idx.unique().numel() == idx.numel()      # sorted, unique
    want = sum(tile_k(ratio, min(TILE, n - t * TILE)) for t in range(-(-n // TILE)))
    assert idx.numel() == want and torch.equal(code["val"], g[idx])


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 40), min_size=0, max_size=3), st.sampled_from(["float32", "bfloat16", "float16", "int64", "uint8"]),
       st.integers(0, 3))
def test_wire_roundtrip_random_tensors(shape, dtype, level):
    dt = getattr(torch, dtype)
    t = (torch.randn(shape) * 10).to(dt) if dt.is_floating_point else torch.randint(0, 100, shape).to(dt)
    obj = {"t": t, "meta": {"shape": tuple(shape), "k": [1, 2.5, "s"]}}
    back = ser.loads(ser.decompress(ser.compress(ser.dumps(obj), level=level)))
    assert back["t"].dtype == dt and back["t"].shape == t.shape and torch.equal(back["t"], t) and back["meta"] == obj["meta"]


def test_dumps_framed_matches_compress_of_dumps_and_views_roundtrip():
    """One-buffer serialise+frame (level 0) is byte-identical to compress(dumps()), and unframe_view is zero-copy."""
    import numpy as np
    import torch
    from pytorch_ps_mpi_b200 import serialization as ser
    obj = {"w": torch.arange(10, dtype=torch.float32).bfloat16(), "big": np.arange(5000.0), "small": np.arange(3), "s": "x",
           "t": (1, 2.5)}
    for level in (0, 1):
        msg, raw_len = ser.dumps_framed(obj, level=level)
        ref = ser.compress(ser.dumps(obj), level=level)
        assert bytes(msg) == bytes(ref) and raw_len == len(ser.dumps(obj))
        view = ser.unframe_view(msg)
        back = ser.loads(view)
        assert torch.equal(back["w"], obj["w"]) and np.array_equal(back["big"], obj["big"]) and back["t"] == obj["t"]
        if level == 0:
            assert isinstance(view, memoryview) and view.obj is msg          # no copy of the payload
    import pytest
    with pytest.raises(ValueError):
        ser.unframe_view(bytes(ser.dumps_framed(obj)[0])[:40])


def test_allocator_tuning_is_idempotent_and_optional(monkeypatch):
    from pytorch_ps_mpi_b200 import runtime
    monkeypatch.setattr(runtime, "_ALLOC_TUNED", False)
    monkeypatch.setenv("PSB200_MALLOC_TUNE", "0")
    assert runtime.tune_host_allocator() is False
    monkeypatch.setenv("PSB200_MALLOC_TUNE", "1")
    assert runtime.tune_host_allocator() in (True, False)       # True on glibc
    assert runtime.tune_host_allocator() == runtime._ALLOC_TUNED


def test_dumps_is_thread_safe_and_reentrant():
    """The encode pool calls dumps() from several threads at once (ps.py::async_code); each thread has its own pickler."""
    import pickle
    import threading
    import numpy as np
    import torch
    from pytorch_ps_mpi_b200 import serialization as ser
    errs = []

    def work(seed):
        try:
            rng = np.random.default_rng(seed)
            for i in range(40):
                n = int(rng.integers(1, 5000))
                obj = {"t": torch.full((n,), float(seed)), "a": np.full(n, seed, dtype=np.float64), "i": i, "s": "x" * (i % 7),
                       "raw": pickle.PickleBuffer(bytes([seed % 251]) * n)}
                back = ser.loads(ser.dumps(obj))
                assert back["i"] == i and torch.equal(back["t"], obj["t"]) and np.array_equal(back["a"], obj["a"])
                assert bytes(back["raw"]) == bytes([seed % 251]) * n
        except Exception as e:      # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs

    class Weird:                     # dumps() called while another dumps() is on the stack of the same thread
        def __reduce__(self):
            return (bytes, (bytes(ser.dumps({"inner": torch.arange(3)})),))

    inner = ser.loads(ser.loads(ser.dumps({"w": Weird(), "after": torch.arange(4)}))["w"])
    assert torch.equal(inner["inner"], torch.arange(3))


def test_loads_from_readonly_bytes_and_memoryview():
    import numpy as np
    import torch
    from pytorch_ps_mpi_b200 import serialization as ser
    obj = {"t": torch.arange(6, dtype=torch.int32).view(2, 3), "big": np.arange(2000, dtype=np.float32), "e": torch.empty(0, 3)}
    raw = ser.dumps(obj)
    for buf in (raw, bytes(raw), memoryview(raw), memoryview(bytes(raw))):
        back = ser.loads(buf)
        assert torch.equal(back["t"], obj["t"]) and np.array_equal(back["big"], obj["big"]) and back["e"].shape == (0, 3)
