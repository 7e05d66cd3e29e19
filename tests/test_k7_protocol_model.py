"""Discrete-event model of the mbarrier protocol of the window-batched K7 (csrc/cuda/sgns_win.cu): the
input-row ring (in_full / in_empty, 2W+1 arrivals per slot), the per-consumer double-buffered output stages
(out_full / out_empty, deferred release), phase-parity waits exactly as the kernel issues them, the input-ring
producer, the output-stage producer with its multi-position passes, the end markers.  Random schedules must
always run to completion: a schedule that stops making progress is a deadlock (or a parity alias) in the
protocol itself, independent of the hardware.  (This model found the pass-width bug: two positions of the same
consumer in one pass wait on each other -- `PW = min(4, consumer warps)` in the kernel.)"""
import random

import pytest


class Bar:
    """mbarrier: `count` arrivals complete a phase; test(parity) is true once the phase of that parity has
    completed, i.e. the CURRENT phase has the other parity (mbarrier.try_wait.parity semantics)."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != parity


def run_model(NW, W, LEN, pass_width, seed):
    rng = random.Random(seed)
    R = 2 * NW + 2 * W + 2                       # ring slots (make_layout)
    nv = LEN + 4 * W                             # virtual centres of the CTA's chunk
    in_full = [Bar(1) for _ in range(R)]
    in_empty = [Bar(2 * W + 1) for _ in range(R)]
    out_full = [Bar(1) for _ in range(2 * NW)]
    out_empty = [Bar(1) for _ in range(2 * NW)]  # kernel: 8 release lanes; one logical arrival here
    meta = [None] * (2 * NW)
    done = set()

    def in_producer():
        for i0 in range(0, nv, 32):
            nb = min(32, nv - i0)
            for l0 in range(0, nb, R):           # groups of at most R lanes, each lane its own slot
                pending = list(range(l0, min(l0 + R, nb)))
                while pending:                   # the lanes of a group wait and issue independently
                    rng.shuffle(pending)
                    for l in pending[:]:
                        J = i0 + l
                        slot, rnd = J % R, J // R
                        if not in_empty[slot].test((rnd & 1) ^ 1):
                            continue
                        if rnd > 0 and not in_full[slot].test((rnd - 1) & 1):
                            continue
                        in_full[slot].arrive()   # copy lands (or plain arrive for an empty entry)
                        pending.remove(l)
                    yield

    def out_producer():
        for i0 in range(0, nv, 32):
            nb = min(32, nv - i0)
            for l4 in range(0, nb, pass_width):
                grp = [i0 + l4 + g for g in range(pass_width) if l4 + g < nb]
                for ii in grp:                   # every group waits for its stage ...
                    cw, n = ii % NW, ii // NW
                    st = cw * 2 + (n & 1)
                    while not out_empty[st].test(((n >> 1) & 1) ^ 1):
                        yield
                for ii in grp:                   # ... the warp converges, then arms + issues
                    cw, n = ii % NW, ii // NW
                    st = cw * 2 + (n & 1)
                    meta[st] = 1 if 2 * W <= ii < 2 * W + LEN else 0
                    out_full[st].arrive()
                yield
        for k in range(NW):                      # end markers, in stream order
            ii = nv + k
            cw, n = ii % NW, ii // NW
            st = cw * 2 + (n & 1)
            while not out_empty[st].test(((n >> 1) & 1) ^ 1):
                yield
            meta[st] = -1
            out_full[st].arrive()

    def consumer(cw):
        n, i = 0, cw
        while True:
            st = cw * 2 + (n & 1)
            while not out_full[st].test((n >> 1) & 1):
                yield
            if n > 0:
                out_empty[st ^ 1].arrive()       # deferred release of the previous position's stage
            act = meta[st]
            if act < 0:
                return
            if act:
                hw = rng.randint(1, W)
                for d in range(-hw, hw + 1):
                    if d == 0:
                        continue
                    j = i - W + d
                    while not in_full[j % R].test((j // R) & 1):
                        yield
                for _ in range(rng.randint(0, 6)):   # the maths
                    yield
            for lane in range(2 * W + 1):
                j = i - 2 * W + lane
                if j >= 0:
                    in_empty[j % R].arrive()
            done.add(i)
            i += NW
            n += 1

    procs = [in_producer(), out_producer()] + [consumer(c) for c in range(NW)]
    alive = list(range(len(procs)))
    idle = 0
    while alive:
        before = tuple(b.phase for b in in_full + in_empty + out_full + out_empty) + tuple(b.pending for b in in_empty)
        order = alive[:]
        rng.shuffle(order)
        for k in order:
            try:
                next(procs[k])
            except StopIteration:
                alive.remove(k)
        after = tuple(b.phase for b in in_full + in_empty + out_full + out_empty) + tuple(b.pending for b in in_empty)
        idle = idle + 1 if before == after else 0
        if idle > 500:
            return False, sorted(set(range(nv)) - done)[:4]
    return len(done) == nv, []


@pytest.mark.parametrize("NW", [2, 3, 4, 5, 7, 10])
@pytest.mark.parametrize("W", [1, 5, 15])
def test_k7_protocol_has_no_deadlock(NW, W):
    for seed in range(3):
        ok, stuck = run_model(NW, W, LEN=120, pass_width=min(4, NW), seed=seed)
        assert ok, (NW, W, seed, stuck)


def test_k7_four_wide_pass_with_three_consumers_deadlocks():
    """The bug the model found: with fewer than 4 consumer warps a 4-wide pass holds positions (n, c) and
    (n+1, c); the second waits for a stage that is only released after the first has been handed over."""
    ok, stuck = run_model(3, 5, LEN=120, pass_width=4, seed=0)
    assert not ok and stuck
