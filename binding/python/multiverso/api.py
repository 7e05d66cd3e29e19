"""mv.init / shutdown / barrier / workers_num / worker_id / server_id / is_master_worker
(reference: binding/python/multiverso/api.py:12-75; the swapped shutdown/barrier docstrings
of the reference, SURVEY Q21, are fixed)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import multiverso_b200 as _mv

__all__ = ["init", "shutdown", "barrier", "workers_num", "worker_id", "server_id", "is_master_worker"]


def init(sync=False):
    """Initialize multiverso. Must be called before any other call.

    sync=True creates a BSP server: every worker's i-th get returns identical parameters,
    computed after all workers' matching adds; all workers must issue the same number of
    add/get calls. sync=False is the asynchronous parameter server."""
    _mv.init(sync=bool(sync))


def shutdown():
    """Shut multiverso down. Call it when training is finished."""
    _mv.shutdown()


def barrier():
    """Block until every process has reached the barrier (all previous table operations of
    every worker are complete afterwards)."""
    _mv.barrier()


def workers_num():
    """Total number of workers."""
    return _mv.workers_num()


def worker_id():
    """Id (0-based, dense) of this worker, -1 if this process is not a worker."""
    return _mv.worker_id()


def server_id():
    return _mv.server_id()


def is_master_worker():
    """The master worker (id 0) does the one-off jobs: initial values, validation, saving."""
    return _mv.worker_id() == 0
