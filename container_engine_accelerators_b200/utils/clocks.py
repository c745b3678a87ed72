"""Sample SM clocks / power / throttle reasons during a timed region (B200_PROFILING.md recipe).

Primary path: NVML in-process (pynvml) on a 50 ms thread — `nvidia-smi -lms` takes seconds to start on an 8-GPU box,
longer than a whole sweep. Fallback: the recipe's nvidia-smi command line.
"""
from __future__ import annotations

import shutil
import statistics
import subprocess
import tempfile
import threading
import time

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
# nvmlClocksEventReason bits
_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 50):
        self.gpu_index, self.period_ms = gpu_index, period_ms
        self._samples: list = []
        self._reasons: set = set()
        self._max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._proc = None
        self._file = None
        self.source = "none"

    # ------------------------------------------------------------------ NVML thread
    def _nvml_loop(self, pynvml, handle) -> None:
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                pw = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                try:
                    bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self._samples.append((float(sm), pw))
                for bit, name in _REASONS.items():
                    if bits & bit:
                        self._reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period_ms / 1000.0)

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self._max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self._thread.start()
            self.source = "nvml"
            return self
        except Exception:
            pass
        if shutil.which("nvidia-smi") is not None:
            self._file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index), "-lms", "200"],
                                          stdout=self._file, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
            time.sleep(0.5)
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        return False

    def summary(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": self._max_mhz, "power_w_max": None, "reasons": [], "samples": 0, "source": self.source}
        sm = [s[0] for s in self._samples]
        pw = [s[1] for s in self._samples]
        reasons = set(self._reasons)
        mx = [self._max_mhz] if self._max_mhz else []
        if self._file is not None:
            self._file.flush()
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
            hi = sorted(sm)[len(sm) // 2:]   # "under load": the upper half (launch gaps between sizes pull a plain median down)
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx) if mx else None, power_w_max=max(pw) if pw else None, reasons=sorted(reasons), samples=len(sm))
        return out
