"""Timer (reference include/multiverso/util/timer.h:9-24): Start(), elapse() in ms."""
import time


class Timer:
    def __init__(self):
        self.start()

    def start(self) -> None:
        self._t0 = time.perf_counter()

    def elapse(self) -> float:
        return (time.perf_counter() - self._t0) * 1e3
