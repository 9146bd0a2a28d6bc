"""L2 wire format for the generic-object (host) path: pointer-based, copy-free tensor framing.

What the reference does: ``to_np → pickle.dumps → blosc.compress`` for every message
(``/root/reference/mpi_comms.py:186-193``), with an unfinished experiment that compresses
straight out of tensor memory via ``blosc.compress_ptr(data_ptr, numel, element_size)``
(``/root/reference/serialization.py:8-36``; its ``compress`` returns an undefined name and the
``__main__`` demo calls undefined ``dumps``/``loads``).  This module finishes that idea:

``dumps(obj)`` lets the C pickler walk the object; every torch tensor is lifted out of the stream
(``reducer_override``) and every large ndarray leaves it as a protocol-5 out-of-band buffer, so the
pickle holds only a *skeleton*; the raw bytes are appended behind it, taken from ``data_ptr()``
(one memcpy, done by the native ``host_codec`` when the extension is built; no numpy round trip, so
bf16 / fp8 tensors survive — the reference's ``to_torch`` silently turned everything into float32,
``mpi_comms.py:48``).  ``loads`` rebuilds tensors and arrays as zero-copy views of the received buffer.

Frame (little endian)::

    magic 'PSB2' | u8 version | u8 codec | u8 typesize | u8 reserved | u64 raw_len      (16 B)
    [codec 0: raw bytes | codec 1: zlib(shuffle(raw)) | codec 2: zlib(raw)]
    raw := u32 skel_len | u32 nentries | nentries x (u8 dtype, u8 ndim, u16 0, u32 0, u64 nbytes,
           ndim x i64 dims) | skeleton pickle | pad16 | entry bytes, each padded to 16
           (dtype 255 = an out-of-band ndarray buffer of the pickle itself)

The explicit ``raw_len`` header replaces the reference's 32-byte ``0x29`` sentinel + 10x
over-allocated slots (``mpi_comms.py:80-85,96-104``) which could collide with payload bytes.
"""
from __future__ import annotations

import io
import os
import pickle
import struct
import threading
import zlib
from typing import Any, List, Tuple

import numpy as np
import torch

__all__ = ["dumps", "loads", "dumps_framed", "unframe_view", "compress", "decompress", "frame", "unframe", "tensor_info",
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


# Objects are walked by the C pickler itself (no Python-level recursion): ``reducer_override`` lifts torch tensors out
# of the stream (``_predump``, ``/root/reference/serialization.py:14-19``), protocol-5 out-of-band buffers lift large
# ndarrays out, small ndarrays stay in-band where a memcpy is cheaper than a header entry.
_INBAND_BYTES = int(os.environ.get("PSB200_INBAND_BYTES", 4096))   # buffers up to this size are pickled in-band (a memcpy beats a table entry)
_RAW_BUFFER = 255             # dtype code of a protocol-5 out-of-band buffer in the tensor table
_TLS = threading.local()


def _restore_tensor(i: int) -> torch.Tensor:
    return _TLS.tensors[i]


def _from_map(items: list) -> list:
    return items


class _Pickler(pickle.Pickler):
    """One instance per thread, reused across :func:`dumps` calls (constructing a pickler costs more than a small dump)."""

    def __init__(self):
        self.file = io.BytesIO()
        self.tensors: List[torch.Tensor] = []
        self.bufs: list = []
        super().__init__(self.file, protocol=5, buffer_callback=self._buffer)

    def _buffer(self, pb) -> bool:
        if pb.raw().nbytes <= _INBAND_BYTES:
            return True                   # small array: stays in the pickle stream
        self.bufs.append(pb)
        return False

    def reducer_override(self, obj):      # called by the C pickler for non-builtin objects only
        if isinstance(obj, torch.Tensor):
            t = obj.detach()
            if t.is_cuda:
                t = t.cpu()               # device→host staging (slow path only)
            if not t.is_contiguous():
                t = t.contiguous()
            self.tensors.append(t)
            return (_restore_tensor, (len(self.tensors) - 1,))
        if isinstance(obj, map):          # the reference's to_np accepts map objects (mpi_comms.py:33-43)
            return (_from_map, (list(obj),))
        return NotImplemented


def _pickler() -> "_Pickler":
    pk = getattr(_TLS, "pickler", None)
    if pk is None or getattr(_TLS, "busy", False):     # first use in this thread, or a nested dumps()
        pk = _Pickler()
        if not getattr(_TLS, "busy", False):
            _TLS.pickler = pk
    return pk


def _pad16(n: int) -> int:
    return (n + 15) & ~15


_HDR = struct.Struct("<II")
_ENT = struct.Struct("<BBHIQ")


def dumps(obj: Any, headroom: int = 0) -> bytearray:
    """Serialise ``obj`` to the raw (unframed) byte layout; tensor / large-array bytes are copied exactly once.

    ``headroom`` (a multiple of 16) leaves that many bytes free in front of the layout, so a frame header can be
    written in place (:func:`dumps_framed`) instead of copying the whole message once more."""
    pk = _pickler()
    outer_busy = getattr(_TLS, "busy", False)
    _TLS.busy = True
    try:
        f = pk.file
        f.seek(0)
        f.truncate()
        pk.clear_memo()
        pk.dump(obj)
        tensors, bufs = pk.tensors, pk.bufs
        skel = f.getvalue()
        H = headroom
        if not tensors and not bufs:           # plain Python payload (and small arrays): header + pickle
            out = bytearray(H + _pad16(8 + len(skel)))
            _HDR.pack_into(out, H, len(skel), 0)
            out[H + 8: H + 8 + len(skel)] = skel
            return out
        parts = [_HDR.pack(len(skel), len(tensors) + len(bufs))]
        sizes = []
        for t in tensors:
            nb = t.numel() * t.element_size()
            sizes.append(nb)
            parts.append(_ENT.pack(_DT2CODE[t.dtype], t.dim(), 0, 0, nb))
            parts.append(struct.pack(f"<{t.dim()}q", *t.shape))
        raws = [pb.raw() for pb in bufs]
        for r in raws:
            sizes.append(r.nbytes)
            parts.append(_ENT.pack(_RAW_BUFFER, 0, 0, 0, r.nbytes))
        hb = b"".join(parts)
        off = H + _pad16(len(hb) + len(skel))
        offs = []
        for nb in sizes:
            offs.append(off)
            off = H + _pad16(off - H + nb)
        out = bytearray(off)
        out[H: H + len(hb)] = hb
        out[H + len(hb): H + len(hb) + len(skel)] = skel
        nt = len(tensors)
        nat = _native() if nt else None
        if nat is not None:
            nat.pack_ptrs(out, offs[:nt], [t.data_ptr() for t in tensors], sizes[:nt])
        elif nt:
            mv = memoryview(out)
            for t, o, nb in zip(tensors, offs, sizes):
                if nb:
                    mv[o: o + nb] = t.reshape(-1).view(torch.uint8).numpy().data
        if raws:
            mv = memoryview(out)
            for r, o in zip(raws, offs[nt:]):
                mv[o: o + r.nbytes] = r
        return out
    finally:
        pk.tensors.clear()                     # the sources were kept alive until the copies above
        pk.bufs.clear()
        _TLS.busy = outer_busy


def loads(buf) -> Any:
    """Inverse of :func:`dumps`; tensors and large arrays are zero-copy views into ``buf`` (keep it alive)."""
    mv = memoryview(buf)
    skel_len, nent = _HDR.unpack_from(mv, 0)
    p = 8
    metas: List[Tuple[int, Tuple[int, ...], int]] = []
    for _ in range(nent):
        code, ndim, _, _, nbytes = _ENT.unpack_from(mv, p)
        p += 16
        dims = struct.unpack_from(f"<{ndim}q", mv, p) if ndim else ()
        p += 8 * ndim
        metas.append((code, dims, nbytes))
    skel = mv[p: p + skel_len]
    off = _pad16(p + skel_len)
    tensors, buffers = [], []
    base = None
    for code, dims, nbytes in metas:
        if code == _RAW_BUFFER:
            buffers.append(mv[off: off + nbytes])
        else:
            dt = _DTYPES[code]
            if nbytes:
                if mv.readonly:
                    raw = torch.frombuffer(bytearray(mv[off: off + nbytes]), dtype=torch.uint8)
                else:
                    if base is None:
                        base = np.frombuffer(mv, dtype=np.uint8)
                    raw = torch.from_numpy(base[off: off + nbytes])
                t = raw.view(dt).view(dims)
            else:
                t = torch.empty(dims, dtype=dt)
            tensors.append(t)
        off = _pad16(off + nbytes)
    prev = getattr(_TLS, "tensors", None)
    _TLS.tensors = tensors
    try:
        return pickle.loads(skel, buffers=buffers)
    finally:
        _TLS.tensors = prev


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


def dumps_framed(obj: Any, level: int = 0) -> Tuple[bytearray, int]:
    """``compress(dumps(obj), level)`` → ``(message, raw_len)``.  At level 0 the 16-byte header is written into the
    head room of the serialised buffer: one allocation and one copy of the tensor bytes per message instead of two."""
    if level <= 0:
        out = dumps(obj, headroom=HEADER_BYTES)
        raw_len = len(out) - HEADER_BYTES
        out[:HEADER_BYTES] = MAGIC + struct.pack("<BBBBQ", _VERSION, _CODEC_STORE, 4, 0, raw_len)
        return out, raw_len
    raw = dumps(obj)
    return frame(raw, level=level), len(raw)


def unframe_view(msg):
    """:func:`unframe` without the copy for stored (level-0) frames: a memoryview into ``msg`` (keep ``msg`` alive)."""
    mv = memoryview(msg)
    if len(mv) >= HEADER_BYTES and bytes(mv[:4]) == MAGIC and mv[5] == _CODEC_STORE and mv[4] == _VERSION:
        raw_len = struct.unpack_from("<Q", mv, 8)[0]
        if len(mv) - HEADER_BYTES < raw_len:
            raise ValueError(f"unframe: message truncated ({len(mv) - HEADER_BYTES} < {raw_len} bytes)")
        return mv[HEADER_BYTES: HEADER_BYTES + raw_len]
    return unframe(msg)


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
