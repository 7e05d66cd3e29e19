"""CPU-side tests of the pieces of the C++ runtime's NVLS allocation path (csrc/device_rt/vmm.cpp) that do not
need a GPU: the descriptor hand-over between sibling processes (pidfd_open + pidfd_getfd, what carries the
cuMemExportToShareableHandle descriptors from one rank to the others) and the graceful "not available" answer on
a machine without a driver."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multiverso_b200", "_lib", "libmvdevice.so")


def _lib():
    if not os.path.exists(LIB):      # build() made it; never relink libraries this process may have loaded
        from multiverso_b200 import _build
        _build.build_device_rt()
    if not os.path.exists(LIB):
        pytest.skip("libmvdevice.so not built (no g++ / libmvb200.so)")
    return ctypes.CDLL(LIB)


CHILD = r"""
import ctypes, os, sys
lib = ctypes.CDLL(sys.argv[1])
fd = lib.mvd_dup_fd_from_pid(int(sys.argv[2]), int(sys.argv[3]))
if fd < 0:
    print("ERR", ctypes.get_errno()); sys.exit(3)
os.lseek(fd, 0, os.SEEK_SET)
sys.stdout.write(os.read(fd, 64).decode())
os.write(fd, b"+child")
"""


def test_fd_duplication_between_sibling_processes():
    lib = _lib()
    if not hasattr(os, "memfd_create"):
        pytest.skip("no memfd_create")
    lib.mvd_allow_fd_duplication()
    fd = os.memfd_create("mvb200-test")
    os.write(fd, b"slab-of-rank-0")
    r = subprocess.run([sys.executable, "-c", CHILD, LIB, str(os.getpid()), str(fd)], capture_output=True, text=True,
                       timeout=60)
    if r.returncode == 3:
        pytest.skip("pidfd_getfd is not permitted in this sandbox: " + r.stdout.strip())
    assert r.returncode == 0, r.stderr
    assert r.stdout == "slab-of-rank-0"
    # same open file description: the child's write is visible through the parent's descriptor
    os.lseek(fd, 0, os.SEEK_SET)
    assert os.read(fd, 64) == b"slab-of-rank-0+child"
    # own pid: plain dup
    own = lib.mvd_dup_fd_from_pid(os.getpid(), fd)
    assert own >= 0 and own != fd
    os.close(own)
    os.close(fd)


def test_vmm_reports_unavailable_without_a_driver():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box: availability depends on the fabric")
    lib = _lib()
    why = ctypes.create_string_buffer(256)
    assert lib.mvd_vmm_available(0, why, 256) == 0
    assert why.value        # a reason, not a crash
