"""CPU tests of the C++ host runtime (no GPU): the reference's unit tier
(Test/unittests/*.cpp, world size 1 loop-back) and end-to-end tier
(`mpirun -np 4 ./multiverso.test kv|array|net|matrix|allreduce`, Test/main.cpp) re-created
on the TCP backend with a local rank forker, plus the extra scenarios (sparse wire
compression, async mode, stateful updaters + checkpoint)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "bin", "mv_test")


@pytest.fixture(scope="module", autouse=True)
def built():
    sys.path.insert(0, ROOT)
    from multiverso_b200 import _build
    _build.build_host()
    assert os.path.exists(BIN)


def run_mp(n, *cmd, timeout=120):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", str(n),
                        "--timeout", str(timeout), "--", *cmd], capture_output=True, text=True,
                       timeout=timeout + 30)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_native_unit_suite():
    r = subprocess.run([BIN, "unit"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("scenario", ["net", "kv", "array", "array_async", "matrix", "sparse", "allreduce", "apptables"])
@pytest.mark.parametrize("nproc", [2, 4])
def test_native_multiprocess(scenario, nproc):
    out = run_mp(nproc, BIN, scenario)
    assert out.count("PASS") == nproc


def test_allreduce_non_power_of_two():
    for n in (3, 5):
        assert run_mp(n, BIN, "allreduce").count("PASS") == n


@pytest.mark.parametrize("updater", ["default", "sgd", "momentum_sgd", "adagrad", "dcasgd", "dcasgda"])
def test_native_updaters_and_checkpoint(updater):
    assert run_mp(2, BIN, f"updater:{updater}").count("PASS") == 2


@pytest.mark.parametrize("scenario", ["dense_perf", "sparse_perf"])
def test_matrix_perf_scenarios(scenario):
    # Test/test_matrix_perf.cpp (reference "perf" tier), small row count: values verified, timings printed
    out = run_mp(3, BIN, scenario, "6000")
    assert out.count("PASS") == 3 and "add 100% of the rows" in out


def test_backup_worker_ratio_flag():
    # with 4 workers and 25% backup workers the BSP quorum is 3: the scenario still completes
    out = run_mp(4, BIN, "array", "-backup_worker_ratio=25")
    assert out.count("PASS") == 4


def test_role_separation():
    # rank 0 server-only, others worker-only (ps_role flag through the environment is per
    # process, so drive it with a small Python script)
    script = os.path.join(ROOT, "tests", "mp_host_roles.py")
    out = run_mp(3, sys.executable, script)
    assert out.count("roles ok") == 3


def test_local_stream_roundtrip(tmp_path):
    r = subprocess.run([BIN, "stream", str(tmp_path / "blob.bin")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([BIN, "stream", "file://" + str(tmp_path / "blob2.bin")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


def test_hdfs_stream(tmp_path):
    """hdfs:// URIs go through a run-time loaded libhdfs (io/hdfs_stream.h).  A stand-in
    libhdfs backed by a local directory exercises connect / open / short reads / append;
    without any libhdfs the open fails with a diagnostic instead of crashing."""
    fake = tmp_path / "libhdfs.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-o", str(fake), os.path.join(ROOT, "tests", "fake_libhdfs.c")],
                   check=True)
    root = tmp_path / "dfs"
    (root / "models").mkdir(parents=True)
    env = dict(os.environ, MV_LIBHDFS=str(fake), FAKE_HDFS_ROOT=str(root))
    r = subprocess.run([BIN, "stream", "hdfs://namenode:9000/models/shard0.bin"], capture_output=True, text=True,
                       timeout=60, env=env)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    assert (root / "models" / "shard0.bin").stat().st_size == 750000 * 4 + 24
    assert (root / ".connected").read_text().split() == ["namenode", "9000"]
    env = {k: v for k, v in os.environ.items() if k != "MV_LIBHDFS"}
    r = subprocess.run([BIN, "stream", "hdfs://namenode:9000/models/x.bin"], capture_output=True, text=True,
                       timeout=60, env=env)
    assert r.returncode != 0 and "libhdfs.so not found" in (r.stdout + r.stderr)


def test_explicit_endpoint_bootstrap():
    # MV_NetBind / MV_NetConnect before MV_Init (C# binding path)
    out = run_mp(3, sys.executable, os.path.join(ROOT, "tests", "mp_host_netbind.py"))
    assert out.count("netbind ok") == 3


def test_machine_file_bootstrap(tmp_path):
    # -machine_file / -port (ZeroMQ-style bootstrap of the reference, zmq_net.h:25-61): one
    # address per line; several ranks on one host take consecutive ports and MV_RANK picks the line
    import random
    mf = tmp_path / "machines.txt"
    mf.write_text("127.0.0.1\n127.0.0.1\n127.0.0.1\n")
    port = random.randrange(10000, 30000, 16)      # below the ephemeral port range
    out = run_mp(3, BIN, "array", f"-machine_file={mf}", f"-port={port}")
    assert out.count("PASS") == 3


def test_cpp_examples_compile_and_run(tmp_path):
    """examples/cpp/host_tables.cpp is the documented minimal C++ program: build it against the
    in-tree library and run it on 3 ranks, BSP + sgd; device_tables.cpp must at least compile and
    link against libmvdevice (it needs GPUs to run)."""
    lib = os.path.join(ROOT, "multiverso_b200", "_lib")
    exe = str(tmp_path / "host_tables")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "cpp", "host_tables.cpp"), "-o", exe, f"-L{lib}", "-lmultiverso",
                    f"-Wl,-rpath,{lib}", "-pthread", "-fopenmp"], check=True)
    out = run_mp(3, exe, "-sync=true", "-updater_type=sgd")
    assert out.count("w[0] = -1.5, emb[42][0] = -3, counter = 3003") == 3, out
    if os.path.exists(os.path.join(lib, "libmvdevice.so")):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "cpp", "device_tables.cpp"), "-o", str(tmp_path / "device_tables"),
                        f"-L{lib}", "-lmvdevice", "-lmultiverso", "-lmvb200", f"-Wl,-rpath,{lib}"], check=True)


def test_stalled_request_is_reported():
    """A BSP schedule with unequal step counts (rank 1 issues one Add more than rank 0 before rank 0
    reaches shutdown) makes a request wait; the waiting rank must say which table / request / how many
    servers it is waiting for (-request_stall_warn_s) instead of hanging silently, and the run still
    completes once the other rank's FinishTrain arrives."""
    script = os.path.join(ROOT, "tests", "mp_host_stall.py")
    out = run_mp(2, sys.executable, script)
    assert "still waits for" in out and out.count("stall ok") == 2
