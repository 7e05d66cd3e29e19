"""CPU tests of the Python package's utilities and option plumbing: Dashboard / monitor
(reference: include/multiverso/dashboard.h), Log levels / file tee / JSONL metrics
(util/log.h), AddOption's 20-byte layout (updater/updater.h:13-69), MV_CreateTable over option
structs, set_flag / flag errors, and the C-level dashboard of the host runtime."""
import json
import os
import struct
import sys
import time

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


def test_dashboard_monitor_accumulates():
    from multiverso_b200.utils import Dashboard
    from multiverso_b200.utils.dashboard import monitor
    Dashboard.reset()
    for _ in range(3):
        with monitor("UNIT_TEST_REGION"):
            time.sleep(0.002)
    snap = Dashboard.snapshot()
    assert snap["UNIT_TEST_REGION"]["count"] == 3 and snap["UNIT_TEST_REGION"]["total_ms"] >= 5.0
    assert "UNIT_TEST_REGION: count = 3" in Dashboard.watch("UNIT_TEST_REGION")
    assert "not found" in Dashboard.watch("NO_SUCH_MONITOR")
    Dashboard.reset()
    assert Dashboard.snapshot() == {}


def test_log_levels_file_and_metrics(tmp_path, capsys):
    from multiverso_b200.utils import log as L
    lg = L.Logger()
    lg.rank = 3
    lg.reset_log_file(str(tmp_path / "mv.log"))
    lg.debug("hidden %d", 1)
    lg.info("shown %d", 2)
    lg.reset_log_level(L.DEBUG)
    lg.debug("now shown")
    with pytest.raises(L.FatalError):
        lg.fatal("boom %s", "x")
    out = capsys.readouterr().out
    assert "hidden" not in out and "[INFO]" in out and "[rank 3] shown 2" in out and "[FATAL]" in out
    text = (tmp_path / "mv.log").read_text()
    assert "now shown" in text and "boom x" in text
    lg.to_stderr = True
    lg.error("to stderr")
    assert "to stderr" in capsys.readouterr().err
    lg.metric("ignored_before_open", 1)
    lg.open_metrics(str(tmp_path / "m" / "metrics.jsonl"))
    lg.metric("words_per_sec", 123.5, step=7)
    rec = json.loads((tmp_path / "m" / "metrics.jsonl").read_text().splitlines()[0])
    assert rec["name"] == "words_per_sec" and rec["value"] == 123.5 and rec["step"] == 7 and rec["rank"] == 3


def test_add_option_wire_layout():
    from multiverso_b200.tables.options import AddOption, GetOption
    o = AddOption(worker_id=2, momentum=0.5, learning_rate=0.25, rho=0.125, lambda_=2.0)
    raw = o.pack()
    assert len(raw) == 20 and struct.unpack("<iffff", raw) == (2, 0.5, 0.25, 0.125, 2.0)
    back = AddOption.unpack(raw)
    assert (back.worker_id, back.momentum, back.learning_rate, back.rho, back.lambda_) == (2, 0.5, 0.25, 0.125, 2.0)
    d = AddOption(worker_id=0)
    assert (d.momentum, round(d.learning_rate, 6), round(d.rho, 6), round(d.lambda_, 6)) == (0.0, 0.01, 0.1, 0.1)
    assert GetOption(worker_id=5).pack() == struct.pack("<i", 5)


def test_create_table_from_options(mv_host):
    mv = mv_host
    a = mv.create_table(mv.ArrayTableOption(17, "float64"))
    a.add(np.arange(17, dtype=np.float64))
    assert np.array_equal(a.get(), np.arange(17, dtype=np.float64))
    m = mv.create_table(mv.MatrixTableOption(5, 3, "float32", min_value=-0.5, max_value=0.5))
    w = m.get().reshape(5, 3)
    assert w.min() >= -0.5 and w.max() <= 0.5 and np.abs(w).sum() > 0      # server-side random init
    s = mv.create_table(mv.MatrixOption(6, 2, "float32", is_sparse=True))
    s.add_rows([1, 4], np.ones((2, 2), np.float32))
    assert np.array_equal(s.get().reshape(6, 2)[[1, 4]], np.ones((2, 2), np.float32))
    kv = mv.create_table(mv.KVTableOption("int64", "float32"))
    kv.add([7], [1.5])
    assert np.allclose(kv.get([7]), [1.5])
    with pytest.raises(TypeError):
        mv.create_table(object())


def test_set_flag_and_identity(mv_host):
    mv = mv_host
    assert mv.rank() == 0 and mv.size() == 1 and mv.num_workers() == 1 and mv.num_servers() == 1
    assert mv.worker_id() == 0 and mv.server_id() == 0 and mv.is_master_worker()
    assert mv.worker_id_to_rank(0) == 0 and mv.server_id_to_rank(0) == 0
    mv.set_flag("omp_threads", 2)
    assert mv.FLAGS.get("omp_threads") == 2
    with pytest.raises(KeyError):
        mv.set_flag("definitely_not_a_flag", 1)
    x = np.arange(6, dtype=np.float32)
    mv.aggregate(x)                                   # world size 1: identity
    assert np.array_equal(x, np.arange(6, dtype=np.float32))
    mv.dashboard_display()                            # native Dashboard::Display must not raise


def test_host_backend_has_the_async_spellings(mv_host):
    """Scripts written against the device backend (handles + wait) run unchanged on the host backend."""
    mv = mv_host
    a = mv.ArrayTable(8, "float32")
    a.wait(a.add_async(np.ones(8, np.float32)))
    mv.barrier()
    h, out = a.get_async()
    a.wait(h)
    assert np.array_equal(out, np.ones(8, np.float32))
    m = mv.MatrixTable(6, 3, "float32", is_sparse=True)
    m.wait(m.add_rows_async([2, 5], np.full((2, 3), 2.0, np.float32)))
    mv.barrier()
    h, rows = m.get_rows_async([5, 2, 0])
    m.wait(h)
    assert np.array_equal(rows, np.array([[2, 2, 2], [2, 2, 2], [0, 0, 0]], np.float32))
    ids, vals = m.get_stale()                      # first pull: everything
    assert ids.size == 6
    m.add_rows([1], np.ones((1, 3), np.float32))
    mv.barrier()
    ids, vals = m.get_stale()
    assert ids.tolist() == [1] and np.array_equal(vals, np.ones((1, 3), np.float32))
    ids, vals = m.get_stale()
    assert ids.size == 0 and vals.shape == (0, 3)
    m.finish_train()


@pytest.mark.parametrize("objective,out", [(2, 7), (1, 1), (1, 5), (0, 1)])
def test_logreg_dense_gemm_step_matches_reference_maths(objective, out):
    """The GEMM formulation of the dense LogisticRegression step (used for many-class / wide problems)
    against a direct per-sample fp64 implementation of Objective::Predict + Gradient."""
    import torch
    from multiverso_b200.models.logreg import dense_gemm_step
    rng = np.random.default_rng(3)
    n, dim = 33, 19
    x = rng.normal(size=(n, dim)).astype(np.float32)
    w = (0.3 * rng.normal(size=(out, dim))).astype(np.float32)
    labels = rng.integers(0, out if out > 1 else 2, size=n).astype(np.float32)
    if objective == 0:
        labels = rng.normal(size=n).astype(np.float32)
    grad = torch.zeros(out * dim)
    loss, correct, p = dense_gemm_step(torch.from_numpy(x), torch.from_numpy(labels), torch.from_numpy(w).view(-1),
                                       grad, objective, out)
    # reference maths, one sample at a time
    exp_loss, exp_correct, exp_grad = 0.0, 0, np.zeros((out, dim))
    for i in range(n):
        logit = w.astype(np.float64) @ x[i].astype(np.float64)
        y = np.zeros(out)
        if out == 1:
            y[0] = labels[i]
        else:
            y[int(labels[i])] = 1
        if objective == 2 and out > 1:
            e = np.exp(logit - logit.max()); pr = e / e.sum()
            exp_loss -= np.log(pr[int(labels[i])])
        elif objective >= 1:
            pr = 1 / (1 + np.exp(-logit))
            exp_loss -= (y * np.log(pr) + (1 - y) * np.log(1 - pr)).sum()
        else:
            pr = logit
            exp_loss += 0.5 * ((pr - y) ** 2).sum()
        if out > 1:
            exp_correct += int(pr.argmax() == int(labels[i]))
        elif objective >= 1:
            exp_correct += int((pr[0] > 0.5) == (labels[i] > 0.5))
        else:
            exp_correct += int(abs(pr[0] - labels[i]) < 0.5)
        exp_grad += np.outer(pr - y, x[i]) / n
    assert abs(float(loss) - exp_loss) < 1e-3 * max(1.0, abs(exp_loss))
    assert int(correct) == exp_correct
    assert np.allclose(grad.view(out, dim).numpy(), exp_grad, atol=1e-5)
    assert p.shape == (n, out)


def test_get_gemm_on_the_host_backend(mv_host):
    import torch
    mv = mv_host
    t = mv.MatrixTable(12, 8, "float32")
    w = np.arange(96, dtype=np.float32).reshape(12, 8) / 10
    t.add(w)
    mv.barrier()
    x = torch.arange(24, dtype=torch.float32).view(3, 8)
    y = mv.ops.get_gemm(t, x)
    assert torch.allclose(y, x @ torch.from_numpy(w).t())
    assert not mv.ops.get_gemm_supported()


def test_ps_linear_autograd_and_push(mv_host):
    """PSLinear: forward = x @ W^T against the table, backward returns dL/dx and pushes dL/dW * scale
    into the table (default updater adds, so -lr performs the SGD step on the servers)."""
    import torch
    mv = mv_host
    out_f, in_f, lr = 6, 4, 0.5
    t = mv.MatrixTable(out_f, in_f, "float32")
    w0 = torch.linspace(-1, 1, out_f * in_f).view(out_f, in_f)
    t.add(w0.numpy())
    mv.barrier()
    layer = mv.ops.PSLinear(t, push_scale=-lr)
    x = torch.randn(5, in_f, requires_grad=True)
    target = torch.randn(5, out_f)
    loss = ((layer(x) - target) ** 2).sum()
    loss.backward()
    # the same computation with a local weight
    w_ref = w0.clone().requires_grad_(True)
    x_ref = x.detach().clone().requires_grad_(True)
    ((x_ref @ w_ref.t() - target) ** 2).sum().backward()
    assert torch.allclose(x.grad, x_ref.grad, atol=1e-5)
    mv.barrier()
    w_new = torch.as_tensor(t.get()).view(out_f, in_f)
    assert torch.allclose(w_new, w0 - lr * w_ref.grad, atol=1e-5)
    # batched leading dimensions, no push
    frozen = mv.ops.PSLinear(t, push_scale=None)
    y = frozen(torch.ones(2, 3, in_f))
    assert y.shape == (2, 3, out_f) and torch.allclose(y[0, 0], w_new.sum(dim=1), atol=1e-5)
    assert "out_features=6" in repr(frozen)
