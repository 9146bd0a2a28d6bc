#!/usr/bin/env python
"""Serialization micro-benchmark — the script form of ``/root/reference/Serialization-timing.ipynb``.

Same payload and sweep as the notebook (``ipynb:67-73,109-113``): a dict with a float64 ``linspace``
of length ``n``, a 10-element random array, a shape tuple, a float and a short string;
``n`` in 30 log-spaced points of [10, 1e4]; compression levels 0/1/2; 100 repeats; it times dump,
load, compress and decompress and records the byte counts — for three serializers:

* ``pickle``      (the reference's wire format, ``mpi_comms.py:188``)
* ``psb2``        this repo's format (:mod:`pytorch_ps_mpi_b200.serialization`: pickled skeleton +
                  raw tensor bytes by pointer, 16-byte length header)
* ``msgpack``     only if importable (the notebook's second contender)

Prints a table next to the notebook's published numbers and writes a CSV.
"""
from __future__ import annotations

import os
import pickle
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_ps_mpi_b200 import serialization as psb   # noqa: E402

try:
    import msgpack
    import msgpack_numpy
    msgpack_numpy.patch()
    HAVE_MSGPACK = True
except Exception:
    HAVE_MSGPACK = False


def payload(n):
    return {"x": np.linspace(0, 1, int(n)), "y": np.random.rand(10), "shape": (int(n), 1), "lr": 0.1, "name": "w"}


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e6, sum(ts) / len(ts) * 1e6, out


def main(reps=100, out_csv=None):
    ns = np.unique(np.logspace(1, 4, 30).astype(int))
    ser = {"pickle": (pickle.dumps, pickle.loads), "psb2": (psb.dumps, psb.loads)}
    if HAVE_MSGPACK:
        ser["msgpack"] = (msgpack.packb, msgpack.unpackb)
    rows = []
    for n in ns:
        obj = payload(n)
        for name, (dump, load) in ser.items():
            d_min, d_mean, blob = best(lambda: dump(obj), reps)
            l_min, l_mean, back = best(lambda: load(blob), reps)
            assert np.allclose(np.asarray(back["x"]), obj["x"])
            for level in (0, 1, 2):
                if name == "psb2":
                    c_min, c_mean, comp = best(lambda: psb.compress(blob, level=level), reps)
                    x_min, x_mean, _ = best(lambda: psb.decompress(comp), reps)
                else:
                    c_min, c_mean, comp = best(lambda: zlib.compress(bytes(blob), level), reps)
                    x_min, x_mean, _ = best(lambda: zlib.decompress(comp), reps)
                rows.append(dict(n=int(n), serializer=name, level=level, bytes=len(blob), packed_bytes=len(comp),
                                 dump_us=d_min, load_us=l_min, compress_us=c_mean, decompress_us=x_mean))
    import csv
    out_csv = out_csv or os.path.join(ROOT, "profiles", "serialization_timing.csv")
    with open(out_csv, "w", newline="") as f:
        wri = csv.DictWriter(f, fieldnames=list(rows[0]))
        wri.writeheader()
        wri.writerows(rows)

    def pick(name, n, level=0):
        return next(r for r in rows if r["serializer"] == name and r["n"] == n and r["level"] == level)

    print(f"{'metric':46s} {'reference notebook (py3.6, CPU n/a)':>36s} {'here: pickle':>14s} {'here: psb2':>12s}")
    ref = {10: ("97.0 / 47.9 us (dump/load, 436 B)"), 12: ("68.2 / 32.9 us (452 B)"), 16: ("48.9 / 27.9 us (484 B)")}
    for n, txt in ref.items():
        p, q = pick("pickle", n), pick("psb2", n)
        print(f"dump/load n={n:<5d}                             {txt:>36s} {p['dump_us']:6.1f}/{p['load_us']:5.1f} {q['dump_us']:6.1f}/{q['load_us']:5.1f}")
    for n in (10, 10000):
        p, q = pick("pickle", n), pick("psb2", n)
        print(f"min dump / load us at n={n:<6d}                 {'~20->33 / ~11->19 us over the sweep':>36s} {p['dump_us']:6.1f}/{p['load_us']:5.1f} {q['dump_us']:6.1f}/{q['load_us']:5.1f}")
    p0, q0, q1 = pick("pickle", 10000, 0), pick("psb2", 10000, 0), pick("psb2", 10000, 1)
    print(f"level-0 compress us at n=1e4 (bytes in -> out)      {'~90 us (80 KB -> 80 KB + 11 B)':>36s} {p0['compress_us']:14.1f} {q0['compress_us']:12.1f}   ({q0['bytes']} -> {q0['packed_bytes']} B)")
    print(f"level-1 bytes ratio at n=1e4                        {'~0.72x (zlib-1/2)':>36s} {pick('pickle', 10000, 1)['packed_bytes'] / p0['bytes']:14.2f} {q1['packed_bytes'] / q1['bytes']:12.2f}   (psb2 = byte-shuffle + deflate)")
    thr = q0["bytes"] / ((q0["dump_us"] + q0["compress_us"]) * 1e-6) / 1e9
    print(f"serialize+frame throughput at 80 KB                 {'~0.65 GB/s per thread (derived)':>36s} {p0['bytes'] / ((p0['dump_us'] + p0['compress_us']) * 1e-6) / 1e9:11.2f} GB/s {thr:8.2f} GB/s")
    print(f"\nwrote {out_csv} ({len(rows)} rows; msgpack {'included' if HAVE_MSGPACK else 'not installed'})")


if __name__ == "__main__":
    main()
