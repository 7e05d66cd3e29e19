"""`-m gpu` coverage of the native device-plane pieces that round 1 compiled but never ran on a GPU:
the C++ GPU applications (`wordembedding_gpu`, `logreg_gpu`), the GPU-served C API
(`libmultiverso_gpu.so` through tools/check_gpu_c_api.py), `PSLinear` against `nn.Linear`, and device
checkpoints (`mv.save_table` / `mv.load_table`).  Multi-rank variants run through tools/mvrun.py when
two GPUs are visible."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "bin")


def _launcher(n):
    return [] if n == 1 else [sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", str(n), "--timeout", "200", "--"]


def _ranks():
    n = torch.cuda.device_count()
    return [1] + ([2] if n >= 2 else [])


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("n", [1, 2])
def test_wordembedding_gpu_binary(tmp_path, n):
    if n not in _ranks():
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(0)
    corpus = tmp_path / "topics.txt"
    with open(corpus, "w") as f:
        for _ in range(20000):
            t = rng.integers(20)
            f.write(" ".join(f"t{t}w{w}" for w in rng.integers(50, size=rng.integers(5, 20))) + "\n")
    out = tmp_path / "vec.txt"
    r = subprocess.run(_launcher(n) + [os.path.join(BIN, "wordembedding_gpu"), "-train_file", str(corpus), "-output", str(out),
                                       "-size", "32", "-cbow", "0", "-negative", "5", "-epoch", "3", "-min_count", "1",
                                       "-data_block_size", "300000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = _json_lines(r.stdout)
    assert len(res) == n
    for x in res:
        assert x["app"] == "wordembedding_gpu" and x["k7_launches"] > 0
        assert x["epoch_loss"][-1] < x["epoch_loss"][0] and all(np.isfinite(x["epoch_loss"]))
    assert out.exists() and out.stat().st_size > 1000


@pytest.mark.parametrize("n,use_ps", [(1, "false"), (1, "true"), (2, "true")])
def test_logreg_gpu_binary(tmp_path, n, use_ps):
    if n not in _ranks():
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(1)
    D, Cn, Ns = 20, 4, 6000
    Wt = rng.normal(size=(Cn, D)); X = rng.normal(size=(Ns, D)); y = (X @ Wt.T).argmax(1)
    for name, lo, hi in (("train", 0, 5000), ("test", 5000, 6000)):
        with open(tmp_path / f"lr_{name}.txt", "w") as f:
            for xi, yi in zip(X[lo:hi], y[lo:hi]):
                f.write(str(int(yi)) + " " + " ".join("%.4f" % v for v in xi) + "\n")
    cfg = tmp_path / "lr.config"
    cfg.write_text(f"""input_size=20
output_size=4
objective_type=softmax
regular_type=L2
updater_type=sgd
learning_rate=0.5
train_epoch=3
minibatch_size=20
use_ps={use_ps}
pipeline=true
sync_frequency=2
train_file={tmp_path}/lr_train.txt
test_file={tmp_path}/lr_test.txt
output_file={tmp_path}/lr.out
output_model_file={tmp_path}/lr.model
""")
    r = subprocess.run(_launcher(n) + [os.path.join(BIN, "logreg_gpu"), str(cfg)], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = _json_lines(r.stdout)
    assert len(res) == n
    for x in res:
        assert x["kernel_launches"] > 0 and x["test_error"] < 0.1 and x["epoch_loss"][-1] < x["epoch_loss"][0]


@pytest.mark.parametrize("n", [1, 2])
def test_gpu_served_c_api(n):
    if n not in _ranks():
        pytest.skip("needs 2 GPUs")
    r = subprocess.run(_launcher(n) + [sys.executable, os.path.join(ROOT, "tools", "check_gpu_c_api.py")],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.count("gpu c api ok") == n, r.stdout[-2000:] + r.stderr[-2000:]


def test_ps_linear_matches_nn_linear(mv_device):
    """Forward = fused Get+GEMM (tf32 tensor cores), backward pushes dW into the table whose sgd updater
    subtracts it: one step must equal nn.Linear + SGD within tf32 accuracy."""
    import multiverso_b200 as mv
    from multiverso_b200.ops import PSLinear
    torch.manual_seed(0)
    out_f, in_f, B, lr = 384, 256, 192, 0.1
    ref = torch.nn.Linear(in_f, out_f, bias=False).cuda()
    table = mv.MatrixTable(out_f, in_f, "float32", updater="sgd")
    table.add(-ref.weight.detach().reshape(-1))                 # sgd subtracts: W_table = W_ref
    assert torch.allclose(table.get().view(out_f, in_f), ref.weight.detach(), atol=1e-6)
    layer = PSLinear(table, push_scale=lr)
    x = torch.randn(B, in_f, device="cuda", requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    tgt = torch.randn(B, out_f, device="cuda")
    y = layer(x)
    y2 = ref(x2)
    assert (y - y2).abs().max().item() < 2e-2                   # tf32 products, fp32 accumulation
    ((y - tgt) ** 2).mean().backward()
    ((y2 - tgt) ** 2).mean().backward()
    torch.cuda.synchronize()
    assert (x.grad - x2.grad).abs().max().item() < 1e-3
    with torch.no_grad():
        w_exp = ref.weight - lr * ref.weight.grad
    w_got = table.get().view(out_f, in_f)
    assert (w_got - w_exp).abs().max().item() < 1e-3


@pytest.mark.parametrize("updater", ["sgd", "momentum_sgd", "adagrad"])
def test_device_checkpoint_roundtrip_with_state(mv_device, tmp_path, updater):
    """mv.save_table / mv.load_table on the device backend: shard + updater state survive, the next Add
    after a restore behaves exactly like the next Add of the original table."""
    import multiverso_b200 as mv
    n = 100003
    g = torch.Generator(device="cuda").manual_seed(3)
    d1 = torch.randn(n, device="cuda", generator=g) * 0.01
    d2 = torch.randn(n, device="cuda", generator=g) * 0.01
    opt = mv.AddOption(learning_rate=0.01, rho=0.1, momentum=0.5)
    a = mv.ArrayTable(n, "float32", updater=updater)
    a.add(d1, opt)
    path = str(tmp_path / "ckpt")
    assert mv.save_table(a, path)
    a.add(d2, opt)
    expect = a.get().clone()
    b = mv.ArrayTable(n, "float32", updater=updater)
    assert mv.load_table(b, path)
    b.add(d2, opt)
    assert torch.allclose(b.get(), expect, rtol=0, atol=1e-7)


@pytest.mark.parametrize("n,use_ps", [(1, "false"), (1, "true"), (2, "true")])
def test_logreg_gpu_binary_ftrl(tmp_path, n, use_ps):
    """FTRL in the native GPU application, locally and THROUGH the parameter server (ps_model.cpp:38-67 FTRLTable:
    the servers hold z and n, workers push (delta z, delta n))."""
    if n not in _ranks():
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(2)
    D, Ns = 2000, 6000
    wtrue = rng.normal(size=D)
    X = np.zeros((Ns, D), np.float32)
    for i in range(Ns):
        X[i, rng.choice(D, size=20, replace=False)] = 1.0
    y = (X @ wtrue > 0).astype(np.int64)
    for name, lo, hi in (("tr", 0, 5000), ("te", 5000, 6000)):
        with open(tmp_path / f"{name}.svm", "w") as f:
            for xi, yi in zip(X[lo:hi], y[lo:hi]):
                nz = np.nonzero(xi)[0]
                f.write(f"{int(yi)} " + " ".join(f"{k}:{xi[k]:.0f}" for k in nz) + "\n")
    cfg = tmp_path / "ftrl.config"
    cfg.write_text(f"input_size={D}\noutput_size=1\nsparse=true\nobjective_type=ftrl\nupdater_type=ftrl\n"
                   f"train_epoch=4\nminibatch_size=20\nlearning_rate=0.5\ntrain_file={tmp_path}/tr.svm\n"
                   f"test_file={tmp_path}/te.svm\noutput_file=\noutput_model_file=\nuse_ps={use_ps}\n"
                   f"alpha=0.1\nbeta=1\nlambda1=0.01\nlambda2=0\nregular_type=default\nsync_frequency=1\n")
    r = subprocess.run(_launcher(n) + [os.path.join(BIN, "logreg_gpu"), str(cfg)], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = _json_lines(r.stdout)
    assert len(res) == n
    for x in res:
        assert x["test_error"] < 0.4 and x["kernel_launches"] > 0, x
