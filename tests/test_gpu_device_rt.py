"""GPU tests of the native C++ device runtime (include/multiverso/device/device.h,
csrc/device_rt -> libmvdevice.so): build/bin/mv_device_test runs the reference's exact-integer
array / matrix / kv / allreduce scenarios, the updater numerics against the host arithmetic and
a checkpoint replay, all from C++ over the kernel library's C ABI (no torch in the process)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "bin", "mv_device_test")


@pytest.fixture(scope="module", autouse=True)
def built():
    """build() (the driver's build step) already produced the executable; only a missing one is
    built here -- never relink libmvb200.so / libmultiverso.so while this process has them loaded."""
    if not os.path.exists(EXE):
        sys.path.insert(0, ROOT)
        from multiverso_b200 import _build
        _build.build_device_rt()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_device_runtime_single_gpu():
    r = subprocess.run([EXE, "all"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("sync", ["true", "false"])
def test_cpp_device_runtime_two_gpus(sync):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "2", "--timeout", "200", "--",
                        EXE, "all", f"-sync={sync}"], capture_output=True, text=True, timeout=260)
    assert r.returncode == 0 and r.stdout.count("PASS") == 2, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_cpp_device_runtime_nvls_aggregate_four_gpus():
    """-device_nvls=2: the C++ runtime's large float MV_Aggregate reduces in the NVSwitch on a VMM / multicast-bound
    staging buffer (csrc/device_rt/vmm.cpp); on a platform without multicast objects the run must still pass on
    the two-shot fallback."""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "4", "--timeout", "200", "--",
                        EXE, "all", "-device_nvls=2"], capture_output=True, text=True, timeout=260)
    assert r.returncode == 0 and r.stdout.count("PASS") == 4, r.stdout[-3000:] + r.stderr[-3000:]
