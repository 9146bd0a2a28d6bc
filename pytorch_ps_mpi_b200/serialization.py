"""L2 wire format for the generic-object (host) path: pointer-based, copy-free tensor framing.

What the reference does: ``to_np → pickle.dumps → blosc.compress`` for every message
(``/root/reference/mpi_comms.py:186-193``), with an unfinished experiment that compresses
straight out of tensor memory via ``blosc.compress_ptr(data_ptr, numel, element_size)``
(``/root/reference/serialization.py:8-36``; its ``compress`` returns an undefined name and the
``__main__`` demo calls undefined ``dumps``/``loads``).  This module finishes that idea:

``dumps(obj)`` walks the object, swaps every tensor / ndarray for a tiny placeholder, pickles
only that *skeleton*, and appends the raw tensor bytes taken from ``data_ptr()`` (one memcpy,
done by the native ``host_codec`` when the extension is built; no numpy round trip, so bf16 /
fp8 tensors survive — the reference's ``to_torch`` silently turned everything into float32,
``mpi_comms.py:48``).  ``loads`` rebuilds tensors as zero-copy views of the received buffer.

Frame (little endian)::

    magic 'PSB2' | u8 version | u8 codec | u8 typesize | u8 reserved | u64 raw_len      (16 B)
    [codec 0: raw bytes | codec 1: zlib(shuffle(raw)) | codec 2: zlib(raw)]
    raw := u32 skel_len | u32 ntensors | ntensors x (u8 dtype, u8 ndim, u16 0, u32 0, u64 nbytes,
           ndim x i64 dims) | skeleton pickle | pad16 | tensor bytes, each padded to 16

The explicit ``raw_len`` header replaces the reference's 32-byte ``0x29`` sentinel + 10x
over-allocated slots (``mpi_comms.py:80-85,96-104``) which could collide with payload bytes.
"""
from __future__ import annotations

import io
import pickle
import struct
import zlib
from typing import Any, List, Tuple

import numpy as np
import torch

__all__ = ["dumps", "loads", "compress", "decompress", "frame", "unframe", "tensor_info",
           "MAGIC", "HEADER_BYTES"]

MAGIC = b"PSB2"
HEADER_BYTES = 16
_VERSION = 1

_DTYPES: List[torch.dtype] = [
    torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int8, torch.uint8,
    torch.int16, torch.int32, torch.int64, torch.bool, torch.float8_e4m3fn, torch.float8_e5m2,
    torch.complex64,
]
_DT2CODE = {d: i for i, d in enumerate(_DTYPES)}
# dtypes numpy cannot represent travel as raw bytes and are re-viewed on load
_NP_OK = {torch.float32, torch.float64, torch.float16, torch.int8, torch.uint8, torch.int16,
          torch.int32, torch.int64, torch.bool, torch.complex64}


class _TensorRef:
    """Placeholder left in the pickled skeleton where a tensor (or ndarray) was."""

    __slots__ = ("i", "as_numpy")

    def __init__(self, i: int, as_numpy: bool):
        self.i, self.as_numpy = i, as_numpy

    def __reduce__(self):
        return (_TensorRef, (self.i, self.as_numpy))


def tensor_info(t: torch.Tensor) -> dict:
    """``numel / data_ptr / element_size`` of a tensor (``/root/reference/serialization.py:8-11``)."""
    return {"numel": t.numel(), "data_ptr": t.data_ptr(), "element_size": t.element_size()}


_NATIVE = False   # False = not resolved yet; None = unavailable


def _native():
    """The host extension (resolved once; ``None`` when it cannot be built/loaded)."""
    global _NATIVE
    if _NATIVE is False:
        try:
            from .ops import ext
            _NATIVE = ext.host() if ext.host_available() else None
        except Exception:
            _NATIVE = None
    return _NATIVE


def _split(obj: Any, tensors: List[torch.Tensor]) -> Any:
    """Replace tensors/ndarrays by placeholders, collecting them (``_predump``, ``serialization.py:14-19``)."""
    if isinstance(obj, torch.Tensor):
        t = obj.detach()
        if t.is_cuda:
            t = t.cpu()                       # device→host staging (slow path only)
        if not t.is_contiguous():
            t = t.contiguous()
        tensors.append(t)
        return _TensorRef(len(tensors) - 1, False)
    if isinstance(obj, np.ndarray) and obj.dtype != object and obj.dtype.kind in "fiubc":
        t = torch.from_numpy(np.ascontiguousarray(obj))
        tensors.append(t)
        return _TensorRef(len(tensors) - 1, True)
    if isinstance(obj, dict):
        return {k: _split(v, tensors) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_split(v, tensors) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_split(v, tensors) for v in obj)
    if isinstance(obj, map):
        return [_split(v, tensors) for v in obj]
    return obj


def _join(obj: Any, tensors: List[torch.Tensor]) -> Any:
    if isinstance(obj, _TensorRef):
        t = tensors[obj.i]
        return t.numpy() if obj.as_numpy else t
    if isinstance(obj, dict):
        return {k: _join(v, tensors) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_join(v, tensors) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_join(v, tensors) for v in obj)
    return obj


def _pad16(n: int) -> int:
    return (n + 15) & ~15


def dumps(obj: Any) -> bytearray:
    """Serialise ``obj`` to the raw (unframed) byte layout; tensors are copied exactly once."""
    tensors: List[torch.Tensor] = []
    skel = pickle.dumps(_split(obj, tensors), protocol=pickle.HIGHEST_PROTOCOL)
    head = io.BytesIO()
    head.write(struct.pack("<II", len(skel), len(tensors)))
    for t in tensors:
        head.write(struct.pack("<BBHIQ", _DT2CODE[t.dtype], t.dim(), 0, 0, t.numel() * t.element_size()))
        head.write(struct.pack(f"<{t.dim()}q", *t.shape))
    hb = head.getvalue()
    off = _pad16(len(hb) + len(skel))
    offs = []
    for t in tensors:
        offs.append(off)
        off = _pad16(off + t.numel() * t.element_size())
    out = bytearray(off)
    out[: len(hb)] = hb
    out[len(hb): len(hb) + len(skel)] = skel
    nat = _native()
    if nat is not None and tensors:
        nat.pack_ptrs(out, offs, [t.data_ptr() for t in tensors],
                      [t.numel() * t.element_size() for t in tensors])
        # `tensors` keeps the sources alive until here
    else:
        mv = memoryview(out)
        for t, o in zip(tensors, offs):
            nb = t.numel() * t.element_size()
            if nb:
                src = t.reshape(-1).view(torch.uint8).numpy()
                mv[o: o + nb] = src.data
    return out


def loads(buf) -> Any:
    """Inverse of :func:`dumps`; tensors are zero-copy views into ``buf`` (keep it alive)."""
    mv = memoryview(buf)
    skel_len, nt = struct.unpack_from("<II", mv, 0)
    p = 8
    metas: List[Tuple[int, Tuple[int, ...], int]] = []
    for _ in range(nt):
        code, ndim, _, _, nbytes = struct.unpack_from("<BBHIQ", mv, p)
        p += 16
        dims = struct.unpack_from(f"<{ndim}q", mv, p)
        p += 8 * ndim
        metas.append((code, dims, nbytes))
    skel = pickle.loads(mv[p: p + skel_len])
    off = _pad16(p + skel_len)
    tensors = []
    base = np.frombuffer(mv, dtype=np.uint8) if nt else None
    for code, dims, nbytes in metas:
        dt = _DTYPES[code]
        if nbytes:
            raw = torch.from_numpy(base[off: off + nbytes]) if not mv.readonly else \
                torch.frombuffer(bytearray(mv[off: off + nbytes]), dtype=torch.uint8)
            t = raw.view(dt).view(dims)
        else:
            t = torch.empty(dims, dtype=dt)
        tensors.append(t)
        off = _pad16(off + nbytes)
    return _join(skel, tensors)


# ---------------------------------------------------------------------------------------
# framing + optional compression (the blosc stand-in)
# ---------------------------------------------------------------------------------------
_CODEC_STORE, _CODEC_SHUF_ZLIB, _CODEC_ZLIB = 0, 1, 2


def _shuffle(raw, typesize: int) -> bytes:
    nat = _native()
    if nat is not None:
        return nat.byteshuffle(raw, typesize)
    a = np.frombuffer(raw, dtype=np.uint8)
    n = (len(a) // typesize) * typesize
    body = a[:n].reshape(-1, typesize).T.reshape(-1)
    return body.tobytes() + a[n:].tobytes()


def _unshuffle(raw, typesize: int) -> bytearray:
    nat = _native()
    if nat is not None:
        return nat.byteunshuffle(raw, typesize)
    a = np.frombuffer(raw, dtype=np.uint8)
    n = (len(a) // typesize) * typesize
    body = a[:n].reshape(typesize, -1).T.reshape(-1)
    return bytearray(body.tobytes() + a[n:].tobytes())


def frame(raw, level: int = 0, shuffle: bool = True, typesize: int = 4) -> bytearray:
    """Wrap raw bytes in the 16-byte header; ``level`` 0 = store (blosc clevel-0 analogue)."""
    if level <= 0:
        codec, body = _CODEC_STORE, raw
    elif shuffle and typesize > 1:
        codec, body = _CODEC_SHUF_ZLIB, zlib.compress(_shuffle(raw, typesize), level)
    else:
        codec, body = _CODEC_ZLIB, zlib.compress(bytes(raw), level)
    out = bytearray(HEADER_BYTES + len(body))
    out[:HEADER_BYTES] = MAGIC + struct.pack("<BBBBQ", _VERSION, codec, typesize, 0, len(raw))
    out[HEADER_BYTES:] = body
    return out


def unframe(msg) -> bytearray:
    """Validate the header and return the raw bytes (overflow / truncation raises)."""
    mv = memoryview(msg)
    if len(mv) < HEADER_BYTES or bytes(mv[:4]) != MAGIC:
        raise ValueError("unframe: bad magic (truncated or corrupted message)")
    ver, codec, typesize, _, raw_len = struct.unpack_from("<BBBBQ", mv, 4)
    if ver != _VERSION:
        raise ValueError(f"unframe: unknown frame version {ver}")
    body = mv[HEADER_BYTES:]
    if codec == _CODEC_STORE:
        if len(body) < raw_len:
            raise ValueError(f"unframe: message truncated ({len(body)} < {raw_len} bytes)")
        return bytearray(body[:raw_len])
    raw = zlib.decompress(body)
    if codec == _CODEC_SHUF_ZLIB:
        raw = _unshuffle(raw, typesize)
    if len(raw) != raw_len:
        raise ValueError("unframe: length mismatch after decompression")
    return bytearray(raw)


_BANNED = {"lz4", "snappy"}
_KNOWN = {"blosclz", "zlib", "store"}


def compress(msg, level: int = 0, name: str = "blosclz") -> bytearray:
    """``compress(msg, level=0, name='blosclz')`` with the reference signature.

    ``/root/reference/mpi_comms.py:18-26``: blosc, default level 0 (framing only), refuses
    ``lz4`` / ``snappy``.  Here level 0 stores the bytes behind a 16-byte header and levels
    1-9 byte-shuffle + deflate (blosc's shuffle filter with zlib as the entropy stage).
    """
    if name in _BANNED:
        raise ValueError("Do not specify lz4 or snappy (the reference refuses them, "
                         "mpi_comms.py:22-24); use 'blosclz' or 'zlib'")
    if name not in _KNOWN:
        raise ValueError(f"unknown compressor {name!r}")
    if name == "store":
        level = 0
    return frame(msg, level=level, shuffle=(name == "blosclz"))


def decompress(code) -> bytearray:
    """Inverse of :func:`compress` (``/root/reference/mpi_comms.py:28-30``)."""
    return unframe(code)


if __name__ == "__main__":   # the demo the reference's __main__ meant to be (serialization.py:40-50)
    n = int(1e3)
    x = torch.linspace(0, 6.28, n)
    y = torch.sin(x) + torch.randn(n) / 4
    obj = {"x": x, "y": y.bfloat16(), "n": n}
    msg = compress(dumps(obj), level=1)
    back = loads(decompress(msg))
    assert torch.equal(back["x"], x) and torch.equal(back["y"], y.bfloat16()) and back["n"] == n
    print(f"round trip ok: {len(dumps(obj))} raw bytes -> {len(msg)} framed+compressed bytes")
