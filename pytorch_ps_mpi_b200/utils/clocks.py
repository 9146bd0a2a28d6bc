"""Sample ``nvidia-smi`` clocks / throttle reasons while a timed region runs (bench hygiene)."""
from __future__ import annotations

import shutil
import statistics
import subprocess
import threading
from typing import Dict, List, Optional

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
_REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]


class ClockSampler:
    """``with ClockSampler(gpu_index) as c: ...; c.summary()`` → ``{"sm_mhz", "sm_max_mhz", "reasons"}``."""

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
                "samples": len(sm), "power_w_max": max(power) if power else None}
