"""Sample nvidia-smi clocks / throttle reasons during a timed region (B200_PROFILING.md recipe)."""
from __future__ import annotations

import shutil
import statistics
import subprocess
import tempfile

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self._proc = None
        self._file = None

    def __enter__(self):
        if shutil.which("nvidia-smi") is None:
            return self
        self._file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index), "-lms", str(self.period_ms)],
                                      stdout=self._file, stderr=subprocess.DEVNULL)
        return self

    def __exit__(self, *exc):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        return False

    def summary(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        if self._file is None:
            return out
        self._file.flush()
        sm, mx, pw, reasons = [], [], [], set()
        with open(self._file.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if sm:
            # "under load": the upper half of the samples (idle gaps between sizes pull the median down otherwise)
            hi = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons), samples=len(sm))
        return out
