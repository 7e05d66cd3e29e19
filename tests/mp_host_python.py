"""Python API, host backend, 3 processes, BSP mode: exact-integer array / matrix / kv scenarios."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import multiverso_b200 as mv

mv.init(sync=True)
W = mv.workers_num()
t = mv.ArrayTable(1000, "float32", init_value=np.arange(1000))
d = np.ones(1000, np.float32)
for it in range(1, 6):
    t.add(d)
    assert np.array_equal(t.get(), np.arange(1000, dtype=np.float32) + it * W)
m = mv.MatrixTable(11, 10, "int32")
base = (np.arange(110, dtype=np.int32) + 1).reshape(11, 10)
for count in range(1, 4):
    m.add(base)
    m.add_rows([0, 1, 5, 10], base[[0, 1, 5, 10]])
    exp = base * count * W
    exp[[0, 1, 5, 10]] *= 2
    assert np.array_equal(m.get(), exp)
# sparse (delta-pull) table under BSP: a worker whose whole-table delta is all zero must still be SEEN by
# every server (one Add per worker per step) -- otherwise the others' Gets stay parked until FinishTrain
sp = mv.MatrixTable(12, 4, "float32", is_sparse=True)
for it in range(1, 4):
    delta = np.zeros((12, 4), np.float32)
    if mv.rank() != 1:                       # rank 1 contributes nothing at all
        delta[mv.rank() % 12] = 1.0
    sp.add(delta)
    got = sp.get()
    exp_sp = np.zeros((12, 4), np.float32)
    for r in range(W):
        if r != 1:
            exp_sp[r % 12] += it
    assert np.array_equal(np.asarray(got).reshape(12, 4), exp_sp), (it, got)
kv = mv.KVTable("int64", "float32")
kv.add([1, 2, 1000003], [1.0, 2.0, 0.5])
mv.barrier()
assert np.allclose(kv.get([1, 2, 1000003, 9]), [W, 2 * W, 0.5 * W, 0])
# checkpoint: every server writes its shard; a later load restores the table on every rank
import tempfile
ckpt = os.path.join(tempfile.gettempdir(), f"mv_mp_ckpt_{os.environ.get('MV_PORT', '0')}")
before = t.get().copy()
assert mv.save_table(t, ckpt)
t.add(d)
mv.barrier()
assert not np.array_equal(t.get(), before)
assert mv.load_table(t, ckpt)
assert np.array_equal(t.get(), before)
mv.barrier()
if mv.rank() == 0:
    for s in range(mv.num_servers()):
        os.remove(f"{ckpt}.shard{s}")
# the device backend's async spellings on the host backend
h, rows = m.get_rows_async([0, 10])
m.wait(h)
assert np.array_equal(rows, exp[[0, 10]])
x = np.full(10, float(mv.rank() + 1), np.float64)
mv.aggregate(x)
assert np.allclose(x, W * (W + 1) / 2)
mv.barrier()
mv.shutdown()
print("python mp ok")
