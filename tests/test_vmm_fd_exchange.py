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


# ---- the whole collective allocation protocol, executed against a test double of the driver -------------------
@pytest.fixture(scope="module")
def fake_driver_dir(tmp_path_factory):
    import shutil
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    d = tmp_path_factory.mktemp("fakecuda")
    subprocess.run([cc, "-shared", "-fPIC", "-O1", "-o", str(d / "libcuda.so.1"),
                    os.path.join(ROOT, "tests", "fake_libcuda.c")], check=True)
    return str(d)


def _run_selftest(fake_dir, n, fail_rank, fail_fn, expect):
    _lib()
    env = dict(os.environ, LD_LIBRARY_PATH=fake_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", str(n), "--timeout", "90", "--",
                        sys.executable, os.path.join(ROOT, "tests", "mp_vmm_selftest.py"), str(fail_rank), fail_fn, expect],
                       capture_output=True, text=True, timeout=150, env=env)
    if "Operation not permitted" in r.stdout + r.stderr:
        pytest.skip("pidfd_getfd is not permitted in this sandbox")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(f"vmm selftest {expect}") == n, r.stdout[-2000:]
    return r.stdout


@pytest.mark.parametrize("n", [2, 3])
def test_vmm_protocol_allocates_maps_and_releases(fake_driver_dir, n):
    """Every rank sees every peer's slab through its own mapping (descriptors really crossed the processes), the
    multicast view is mapped, two allocate / release rounds leave no descriptor behind."""
    _run_selftest(fake_driver_dir, n, 0, "none", "ok")


@pytest.mark.parametrize("fail_rank,fail_fn", [(1, "cuMemCreate"), (0, "cuMulticastCreate"), (1, "cuMemImportFromShareableHandle"),
                                               (0, "cuMulticastAddDevice"), (1, "cuMulticastBindMem"), (0, "cuMemMap"),
                                               (1, "multicast_attr")])
def test_vmm_protocol_falls_back_collectively(fake_driver_dir, fail_rank, fail_fn):
    """A failure of any step on ONE rank makes EVERY rank return "not available" (nobody hangs in a rendezvous,
    nobody keeps a mapping or a descriptor): the caller then uses the cudaIpc path on all ranks alike."""
    out = _run_selftest(fake_driver_dir, 2, fail_rank, fail_fn, "fallback")
    assert "vmm selftest fallback" in out
