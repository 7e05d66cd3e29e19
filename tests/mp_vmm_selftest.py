"""One rank of the VMM / multicast allocation protocol test (csrc/device_rt/vmm.cpp) against the driver test
double tests/fake_libcuda.c.  Launched by tools/mvrun.py; LD_LIBRARY_PATH points at the directory holding the
fake libcuda.so.1.  argv: <fail_rank> <function to fail on that rank | none> <expected: ok | fallback>."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ["MV_RANK"])
fail_rank, fail_fn, expect = int(sys.argv[1]), sys.argv[2], sys.argv[3]
if fail_fn != "none" and rank == fail_rank:
    os.environ["FAKE_CUDA_FAIL"] = fail_fn

import multiverso_b200 as mv  # noqa: E402

mv.init()                                     # host backend: the TCP control plane the allocation protocol rides on
lib = ctypes.CDLL(os.path.join(ROOT, "multiverso_b200", "_lib", "libmvdevice.so"))
lib.mvd_vmm_selftest_hostmapped.argtypes = [ctypes.c_longlong, ctypes.c_char_p, ctypes.c_int]


def open_fds():
    return len(os.listdir("/proc/self/fd"))


msg = ctypes.create_string_buffer(256)
before = open_fds()
for rep in range(2):                          # twice: the teardown leaves nothing behind
    rc = lib.mvd_vmm_selftest_hostmapped(300000, msg, 256)
    if expect == "ok":
        assert rc == 1, (rank, rc, msg.value)
    else:
        assert rc == 0 and msg.value, (rank, rc, msg.value)
after = open_fds()
assert after == before, f"rank {rank}: descriptor leak {before} -> {after}"
mv.barrier()
mv.shutdown()
print(f"vmm selftest {expect}: {msg.value.decode()}")
