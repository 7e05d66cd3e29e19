"""-ps_role separation on the host backend: rank 0 = server only, ranks 1.. = workers only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import multiverso_b200 as mv

rank = int(os.environ["MV_RANK"])
mv.init(ps_role="server" if rank == 0 else "worker")
assert mv.num_servers() == 1 and mv.num_workers() == mv.size() - 1
assert (mv.server_id() == 0) == (rank == 0)
assert mv.worker_id() == (rank - 1 if rank > 0 else -1)
t = mv.ArrayTable(100, "float32")
if rank > 0:
    t.add(np.ones(100, np.float32) * rank)
mv.barrier()
if rank > 0:
    W = mv.num_workers()
    assert np.allclose(t.get(), W * (W + 1) / 2.0)
mv.barrier()
mv.shutdown()
print("roles ok")
