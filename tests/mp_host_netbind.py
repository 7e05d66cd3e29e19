"""Explicit-endpoint bootstrap of the control plane (MV_NetBind / MV_NetConnect, the path the
C# binding uses; reference: zmq_net.h Bind/Connect): every process binds its own endpoint,
connects to the others, and only then calls init()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import multiverso_b200 as mv

rank, size = int(os.environ["MV_RANK"]), int(os.environ["MV_SIZE"])
base = int(os.environ["MV_PORT"]) + 500
# the launcher's variables must not be what makes this work
for k in ("MV_RANK", "MV_SIZE", "MV_PORT", "RANK", "WORLD_SIZE"):
    os.environ.pop(k, None)
endpoints = [f"127.0.0.1:{base + r}" for r in range(size)]
mv.net_bind(rank, endpoints[rank])
others = [r for r in range(size) if r != rank]
mv.net_connect(others, [endpoints[r] for r in others])
mv.init()
assert mv.rank() == rank and mv.size() == size and mv.workers_num() == size
t = mv.ArrayTable(64, "float32")
t.add(np.full(64, rank + 1.0, np.float32))
mv.barrier()
assert np.allclose(t.get(), size * (size + 1) / 2.0)
mv.barrier()
mv.shutdown()
print("netbind ok")
