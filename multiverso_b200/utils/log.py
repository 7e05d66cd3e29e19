"""Process-global logger with the reference's levels and line format.

Reference: multiverso::Log (include/multiverso/util/log.h:9-142, src/util/log.cpp:82-111):
levels Debug < Info < Error < Fatal, ``[LEVEL] [time] msg``, optional tee to a file,
``-logtostderr``.  Added here: a rank prefix and an optional JSONL metrics sink.
"""
from __future__ import annotations

import json
import os
import sys
import time
from typing import Optional

DEBUG, INFO, ERROR, FATAL = 0, 1, 2, 3
_NAMES = {DEBUG: "DEBUG", INFO: "INFO", ERROR: "ERROR", FATAL: "FATAL"}


class FatalError(RuntimeError):
    """Raised by Log.fatal when kill-on-fatal is disabled (tests)."""


class Logger:
    def __init__(self):
        self.level = INFO
        self.file = None
        self.to_stderr = False
        self.kill_fatal = False
        self.rank: Optional[int] = None
        self._metrics = None

    def reset_log_file(self, path: str) -> None:
        if self.file:
            self.file.close()
        self.file = open(path, "a") if path else None

    def reset_log_level(self, level: int) -> None:
        self.level = level

    def reset_kill_fatal(self, kill: bool) -> None:
        self.kill_fatal = kill

    def write(self, level: int, fmt: str, *args) -> None:
        if level < self.level:
            return
        msg = fmt % args if args else fmt
        stamp = time.strftime("%Y-%m-%d %H:%M:%S")
        rk = f" [rank {self.rank}]" if self.rank is not None else ""
        line = f"[{_NAMES[level]}] [{stamp}]{rk} {msg}"
        if not line.endswith("\n"):
            line += "\n"
        (sys.stderr if self.to_stderr else sys.stdout).write(line)
        if self.file:
            self.file.write(line)
            self.file.flush()

    def debug(self, fmt, *a): self.write(DEBUG, fmt, *a)
    def info(self, fmt, *a): self.write(INFO, fmt, *a)
    def error(self, fmt, *a): self.write(ERROR, fmt, *a)

    def fatal(self, fmt, *a):
        self.write(FATAL, fmt, *a)
        if self.kill_fatal:
            os._exit(1)
        raise FatalError(fmt % a if a else fmt)

    # ---- JSONL metrics sink -------------------------------------------------------------
    def open_metrics(self, path: str) -> None:
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self._metrics = open(path, "a")

    def metric(self, name: str, value, **extra) -> None:
        if self._metrics is None:
            return
        rec = {"ts": time.time(), "rank": self.rank, "name": name, "value": value}
        rec.update(extra)
        self._metrics.write(json.dumps(rec) + "\n")
        self._metrics.flush()


Log = Logger()


def CHECK(cond, msg: str = "CHECK failed"):
    if not cond:
        Log.fatal(msg)


def CHECK_NOTNULL(x, msg: str = "CHECK_NOTNULL failed"):
    if x is None:
        Log.fatal(msg)
    return x
