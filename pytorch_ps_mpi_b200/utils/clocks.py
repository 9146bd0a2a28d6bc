"""Sample SM clocks / throttle reasons while a timed region runs (bench hygiene).

:class:`NvmlClockSampler` reads NVML **in-process** from a background thread (no child process inside the timed region —
round 1's ``nvidia-smi -lms`` child cost ~1 ms/step while it started up); :class:`ClockSampler` is the ``nvidia-smi``
fallback for boxes without ``pynvml``.  Both expose ``mark()`` / ``summary()`` → ``{"sm_mhz", "sm_max_mhz", "reasons"}``.
"""
from __future__ import annotations

import shutil
import statistics
import subprocess
import threading
import time
from typing import Dict, List, Optional

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
_REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
# nvmlClocksEventReasons bit masks (nvml.h)
_NVML_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}


class NvmlClockSampler:
    """``with NvmlClockSampler(cuda_index) as c: ...; c.mark(); ...; c.summary()``.

    A daemon thread polls ``nvmlDeviceGetClockInfo`` / ``...CurrentClocksEventReasons`` / power every ``period_ms``
    (each call is a few microseconds of driver ioctl; no subprocess, no CUDA call)."""

    def __init__(self, cuda_index: int = 0, period_ms: int = 25):
        self.cuda_index, self.period = cuda_index, period_ms / 1e3
        self._samples: List[tuple] = []
        self._mark = 0
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._h = None
        self._nv = None
        self.sm_max = None

    def _open(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.cuda_index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(self.cuda_index)
        self._nv, self._h = nv, h
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))

    def _poll_once(self):
        nv, h = self._nv, self._h
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        try:
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
        except Exception:
            pw = None
        self._samples.append((sm, bits, pw))

    def __enter__(self):
        try:
            self._open()
            self._poll_once()
        except Exception:
            self._h = None
            return self

        def run():
            while not self._stop.wait(self.period):
                try:
                    self._poll_once()
                except Exception:
                    return

        self._thread = threading.Thread(target=run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        return False

    @property
    def ok(self) -> bool:
        return self._h is not None

    def mark(self) -> None:
        """Samples taken before this call (warm-up) are ignored by :meth:`summary`."""
        if self._h is not None:
            try:
                self._poll_once()
            except Exception:
                pass
        self._mark = max(0, len(self._samples) - 1)

    def summary(self) -> Dict:
        if self._h is not None:
            try:
                self._poll_once()
            except Exception:
                pass
        rows = self._samples[self._mark:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({name for _, bits, _ in rows for name, bit in _NVML_BITS.items() if bits & bit})
        power = [p for _, _, p in rows if p is not None]
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(rows), "power_w_max": max(power) if power else None, "source": "nvml in-process"}


class ClockSampler:
    """``nvidia-smi -lms`` child-process fallback: ``with ClockSampler(gpu_index) as c: ...; c.summary()``."""

    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self._proc: Optional[subprocess.Popen] = None
        self._lines: List[str] = []
        self._thread: Optional[threading.Thread] = None
        self._mark = 0

    def __enter__(self):
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        try:
            self._proc = subprocess.Popen(
                [exe, f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index),
                 "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return self

        def pump():
            for line in self._proc.stdout:
                self._lines.append(line.strip())

        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._proc is not None:
            self._proc.terminate()       # exact PID we started
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
            if self._thread is not None:
                self._thread.join(timeout=2)
        return False

    def mark(self) -> None:
        """Samples taken before this call (e.g. during warm-up) are ignored by :meth:`summary`."""
        self._mark = max(0, len(self._lines) - 1)      # keep the most recent one so short regions still have a sample

    def summary(self) -> Dict:
        sm, smax, power = [], [], []
        reasons = set()
        for ln in self._lines[self._mark:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(_REASONS, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvidia-smi child"}
