"""BSP with unequal step counts: rank 1 issues an extra Add + Get while rank 0 sleeps before shutting down."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import multiverso_b200 as mv

mv.init(sync=True, request_stall_warn_s=0.5)
t = mv.ArrayTable(16, "float32")
t.add(np.ones(16, np.float32))
t.get()
if mv.rank() == 1:
    t.add(np.ones(16, np.float32))
    t.get()                                  # waits for rank 0's matching Add -- which never comes ...
else:
    time.sleep(2.0)                          # ... until rank 0 finishes training (shutdown sends FinishTrain)
print("stall ok", flush=True)
mv.shutdown()
