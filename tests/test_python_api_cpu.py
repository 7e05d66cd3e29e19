"""Python API on the host backend (no GPU): the reference binding's tests
(binding/python/multiverso/tests/test_multiverso.py) at world size 1 and 3."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture
def mv_host():
    import torch
    if torch.cuda.is_available():
        pytest.skip("host-backend test (runs where there is no GPU)")
    import multiverso_b200 as mv
    mv.FLAGS.reset()
    mv.init()
    yield mv
    mv.shutdown(finalize_net=False)
    mv.FLAGS.reset()


def test_flags_parse_and_compact():
    from multiverso_b200.utils import FlagRegister
    f = FlagRegister()
    rest = f.parse(["prog", "-sync=true", "-updater_type=sgd", "pos", "-unknown=1", "-omp_threads=8", "-x"])
    assert rest == ["prog", "pos", "-unknown=1", "-x"]
    assert f.get("sync") is True and f.get("updater_type") == "sgd" and f.get("omp_threads") == 8
    with pytest.raises(KeyError):
        f.set("no_such_flag", 1)


def test_array_table_scenario(mv_host):
    """_test_array: 10000 elements x 100 iterations, two adds -> (j+1)(i+1)*2*workers."""
    mv = mv_host
    size = 10000
    t = mv.ArrayTable(size, "float32")
    base = np.arange(1, size + 1, dtype=np.float32)
    for i in range(20):
        t.add(base)
        t.add(base)
        mv.barrier()
        assert np.array_equal(t.get(), base * (i + 1) * 2 * mv.workers_num())
        mv.barrier()


def test_matrix_table_scenario(mv_host):
    """test_matrix: 11x10, whole + row [0,1,5,10] adds, whole and row gets."""
    mv = mv_host
    R, C = 11, 10
    t = mv.MatrixTable(R, C, "float32")
    base = np.arange(R * C, dtype=np.float32).reshape(R, C)
    rows = [0, 1, 5, 10]
    for count in range(1, 6):
        t.add(base)
        t.add_rows(rows, base[rows])
        mv.barrier()
        exp = base * count * mv.workers_num()
        exp[rows] *= 2
        assert np.array_equal(t.get(), exp)
        assert np.array_equal(t.get_rows(rows), exp[rows])
        mv.barrier()


def test_master_init_and_kv_and_aggregate(mv_host):
    mv = mv_host
    t = mv.ArrayTable(16, "float32", init_value=np.full(16, 3.0))
    assert np.array_equal(t.get(), np.full(16, 3.0, np.float32))
    kv = mv.KVTable("int64", "int64")
    kv.add(4, 10 ** 12)
    kv.add(4, 5)
    assert kv.get(4) == 10 ** 12 + 5 and kv.raw()[4] == 10 ** 12 + 5
    x = np.ones(8, np.float32)
    mv.aggregate(x)
    assert np.array_equal(x, np.ones(8, np.float32) * mv.size())
    assert mv.is_master_worker() and mv.worker_id() == 0 and mv.server_id() == 0


def test_sparse_matrix_delta_pull(mv_host):
    mv = mv_host
    t = mv.MatrixTable(8, 4, "float32", is_sparse=True)
    full = t.get()                       # first pull: everything is stale
    assert full.shape == (8, 4) and not full.any()
    d = np.zeros((8, 4), np.float32)
    d[2] = 1.0
    t.add(d)                             # only row 2 is shipped / marked stale
    buf = np.full((8, 4), -1.0, np.float32)
    t.get(out=buf)                       # delta pull touches only row 2
    assert np.array_equal(buf[2], np.ones(4)) and (buf[[0, 1, 3, 4, 5, 6, 7]] == -1).all()


def test_checkpoint_roundtrip_host(mv_host, tmp_path):
    mv = mv_host
    t = mv.ArrayTable(100, "float32")
    t.add(np.arange(100, dtype=np.float32))
    assert t.store(str(tmp_path / "ckpt"))
    t.add(np.ones(100, np.float32))
    assert t.load(str(tmp_path / "ckpt"))
    assert np.array_equal(t.get(), np.arange(100, dtype=np.float32))
    # the backend-independent spelling
    assert mv.save_table(t, str(tmp_path / "ckpt2")) and os.path.exists(tmp_path / "ckpt2.shard0")
    t.add(np.ones(100, np.float32))
    assert mv.load_table(t, str(tmp_path / "ckpt2"))
    assert np.array_equal(t.get(), np.arange(100, dtype=np.float32))
    assert not mv.load_table(t, str(tmp_path / "no_such_checkpoint"))


def test_multiprocess_python_api():
    script = os.path.join(ROOT, "tests", "mp_host_python.py")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "3", "--", sys.executable,
                        script], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and r.stdout.count("python mp ok") == 3, r.stdout[-2000:] + r.stderr[-2000:]


def test_compat_binding_suite_single_and_multi_process():
    """binding/python/multiverso (drop-in for the reference binding) at 1 and 2 processes."""
    suite = os.path.join(ROOT, "binding", "python", "multiverso", "tests", "test_multiverso.py")
    r = subprocess.run([sys.executable, suite], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "2", "--", sys.executable, suite],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("example,args", [
    ("logistic_regression.py", []),
    ("resnet_cifar10.py", ["--epochs", "1", "--n", "1", "--samples", "256"]),
    ("addition_rnn.py", ["--iters", "1", "--samples", "1500", "--digits", "2"]),
])
def test_binding_examples_two_workers(example, args):
    """The binding examples (reference: binding/python/examples/theano/*) run with 2 workers on the host backend."""
    script = os.path.join(ROOT, "binding", "python", "examples", "torch", example)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "2", "--timeout", "280", "--",
                        sys.executable, script, *args], capture_output=True, text=True, timeout=320, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
