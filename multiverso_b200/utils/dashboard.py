"""Named cumulative monitors, timed on the device.

Reference: Dashboard / Monitor / MONITOR_BEGIN / MONITOR_END
(include/multiverso/dashboard.h:16-76, src/dashboard.cpp:14-49): wall-clock timers with
count / total ms / average.  Here a monitor can also be bracketed by CUDA events on the
op's stream (device time, which is what every multi-GPU number must be), with
an optional byte count so ``Display`` reports achieved GB/s per table op, and NVTX ranges.
"""
from __future__ import annotations

import threading
import time
from contextlib import contextmanager
from typing import Dict, List

from .log import Log


class Monitor:
    def __init__(self, name: str):
        self.name = name
        self.count = 0
        self.elapse_ms = 0.0
        self.bytes = 0
        self._t0 = 0.0
        self._pending: List = []   # (start_event, end_event, bytes)
        self._lock = threading.Lock()

    # host wall-clock flavour (same contract as the reference)
    def begin(self) -> None:
        self._t0 = time.perf_counter()

    def end(self) -> None:
        with self._lock:
            self.elapse_ms += (time.perf_counter() - self._t0) * 1e3
            self.count += 1

    # device flavour
    def begin_cuda(self):
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def end_cuda(self, start_ev, nbytes: int = 0) -> None:
        import torch
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        with self._lock:
            self._pending.append((start_ev, ev, nbytes))

    def _drain(self) -> None:
        keep = []
        for s, e, nb in self._pending:
            if e.query():
                self.elapse_ms += s.elapsed_time(e)
                self.count += 1
                self.bytes += nb
            else:
                keep.append((s, e, nb))
        self._pending = keep

    def average(self) -> float:
        return self.elapse_ms / self.count if self.count else 0.0

    def info_string(self) -> str:
        with self._lock:
            self._drain()
        s = (f"[Monitor] {self.name}: count = {self.count} elapse = {self.elapse_ms:.3f}ms "
             f"average = {self.average():.4f}ms")
        if self.bytes and self.elapse_ms > 0:
            s += f" bandwidth = {self.bytes / self.elapse_ms / 1e6:.1f} GB/s"
        return s


class _Dashboard:
    def __init__(self):
        self._record: Dict[str, Monitor] = {}
        self._lock = threading.Lock()

    def monitor(self, name: str) -> Monitor:
        with self._lock:
            m = self._record.get(name)
            if m is None:
                m = self._record[name] = Monitor(name)
            return m

    def watch(self, name: str) -> str:
        m = self._record.get(name)
        return m.info_string() if m else f"[Monitor] {name}: not found"

    def display(self) -> None:
        Log.info("--------------Show dashboard monitor information--------------")
        for name in sorted(self._record):
            Log.info("%s", self._record[name].info_string())
        for tid, st in self.staleness().items():
            if st.get("adds"):
                Log.info("[Staleness] table %s: %d adds, mean %.2f, p50 %d, p99 %d (other workers' adds between "
                         "my Get and my Add)", tid, st["adds"], st["mean"], st["p50"], st["p99"])
        Log.info("---------------------------------------------------------------")

    def snapshot(self) -> dict:
        """{name: {count, total_ms, avg_ms, gbs}} of every monitor (device-timed where applicable)."""
        out = {}
        for name, m in sorted(self._record.items()):
            m.info_string()   # drains finished CUDA events
            out[name] = {"count": m.count, "total_ms": round(m.elapse_ms, 3), "avg_ms": round(m.average(), 4)}
            if m.bytes and m.elapse_ms > 0:
                out[name]["gbs"] = round(m.bytes / m.elapse_ms / 1e6, 1)
        return out

    def staleness(self) -> dict:
        """Staleness histograms of every table created with ``-staleness=true``:
        {table_id: {"adds": n, "mean": m, "p50": .., "p99": .., "max_bin": .., "hist": [64 bins]}} where bin k counts
        this worker's shard updates (one per Add and shard) that were applied after k Adds of OTHER workers to that
        shard since this worker last pulled it (last bin: >= 63)."""
        from ..runtime import Runtime
        out = {}
        rt = Runtime._inst
        if rt is None:
            return out
        for t in getattr(rt, "tables", []):
            h = getattr(t, "staleness_hist", None)
            if h is None:
                continue
            hist = [int(v) for v in h.cpu().tolist()]
            n = sum(hist)
            if n == 0:
                out[getattr(t, "table_id", len(out))] = {"adds": 0, "hist": hist}
                continue
            cum, p50, p99 = 0, None, None
            for k, v in enumerate(hist):
                cum += v
                if p50 is None and cum >= 0.5 * n:
                    p50 = k
                if p99 is None and cum >= 0.99 * n:
                    p99 = k
            out[getattr(t, "table_id", len(out))] = {
                "adds": n, "mean": round(sum(k * v for k, v in enumerate(hist)) / n, 3), "p50": p50, "p99": p99,
                "max_bin": max(k for k, v in enumerate(hist) if v), "hist": hist}
        return out

    def reset(self) -> None:
        with self._lock:
            self._record.clear()


Dashboard = _Dashboard()


@contextmanager
def monitor(name: str, cuda: bool = False, nbytes: int = 0, nvtx: bool = False):
    """``with monitor("WORKER_TABLE_SYNC_ADD", cuda=True, nbytes=n):`` == MONITOR_BEGIN/END."""
    m = Dashboard.monitor(name)
    rng = None
    if nvtx:
        try:
            import torch
            torch.cuda.nvtx.range_push(name)
            rng = True
        except Exception:
            rng = None
    if cuda:
        ev = m.begin_cuda()
        try:
            yield m
        finally:
            m.end_cuda(ev, nbytes)
    else:
        m.begin()
        try:
            yield m
        finally:
            m.end()
    if rng:
        import torch
        torch.cuda.nvtx.range_pop()
