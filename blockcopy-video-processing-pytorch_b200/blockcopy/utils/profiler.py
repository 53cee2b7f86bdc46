"""Named, levelled timing regions -- ``blockcopy.utils.profiler.timings``.

Same API and report format as the reference (utils/profiler.py:7-61), because the consumers
import it by path (swiftnet.py:11, swiftnet/util.py:6, csp_blockcopy.py:7): a region declared as
``with timings.env(name, level)`` is measured only when ``level <= timings.level``; the default
level 0 makes every region of the package (all declared at level >= 1) free.

Measured regions synchronise the device on both sides, as in the reference, so a report taken at
level >= 1 is a breakdown, not a throughput number.  ``timings.use_events = True`` switches to
CUDA events recorded on the current stream (no device-wide synchronisation while measuring;
durations are resolved when the report is printed).
"""
from __future__ import annotations

import time
from collections import defaultdict
from contextlib import contextmanager

import torch


class Timings:
    def __init__(self, level: int = 0):
        self.level = level
        self.average = True
        self.use_events = False
        self.cnt = 0
        self.reset()

    def reset(self):
        self.records = defaultdict(float)  # name -> seconds
        self.starts = {}                   # name -> perf_counter value or cuda event
        self.counts = defaultdict(int)
        self._pending = []                 # (name, start_event, stop_event)
        self.cnt = 0

    def add_cnt(self, cnt: int = 1):
        if self.level >= 0:
            self.cnt += cnt

    def set_level(self, level: int):
        self.level = level

    # ------------------------------------------------------------------ start / stop
    def start(self, name: str, level: int = 0):
        if level > self.level:
            return
        if self.use_events and torch.cuda.is_available():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.starts[name] = ev
        else:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.starts[name] = time.perf_counter()
        self.counts[name] += 1

    def stop(self, name: str, level: int = 0):
        begin = self.starts.pop(name, None)
        if begin is None:
            return
        if isinstance(begin, float):
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.records[name] += time.perf_counter() - begin
        else:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self._pending.append((name, begin, end))

    def _resolve(self):
        if self._pending:
            torch.cuda.synchronize()
            for name, a, b in self._pending:
                self.records[name] += a.elapsed_time(b) * 1e-3
            self._pending = []

    @contextmanager
    def env(self, name: str, level: int = 0):
        self.start(name, level)
        try:
            yield
        finally:
            self.stop(name)

    def __repr__(self) -> str:
        self._resolve()
        if self.cnt == 0:
            return "## Profiler: no batches registered"
        if self.cnt < 0:
            return "## Profiler: disabled"
        lines = [f"### Profiler (images: {self.cnt})###"]
        for name in sorted(self.records):
            ms = self.records[name] * 1000
            lines.append(
                f"# {name:20}: {ms / self.cnt:4.3f} ms per image (number of calls: {self.counts[name]}, "
                f"per call: {ms / max(1, self.counts[name]):4.3f} ms) ")
        return "\n".join(lines) + "\n"


timings = Timings(level=0)
