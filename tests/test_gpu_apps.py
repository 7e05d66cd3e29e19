"""End-to-end application tests on one GPU: the wordembedding CLI on a generated corpus and
LogisticRegression (dense softmax / sparse sigmoid / FTRL, local and parameter-server)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _corpus(path, n_sent=3000, vocab=500, seed=0):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for _ in range(n_sent):
            base = rng.integers(0, vocab // 2, size=12) * 2
            words = np.stack([base, base + 1], 1).reshape(-1)      # w(2i) is followed by w(2i+1)
            f.write(" ".join(f"w{int(x)}" for x in words) + "\n")


@pytest.mark.parametrize("flags", ["-cbow 0 -negative 5", "-cbow 1 -negative 5", "-cbow 0 -hs 1 -negative 0"])
def test_wordembedding_app_cli(tmp_path, flags):
    _need_gpu()
    from multiverso_b200.apps import wordembedding as app
    import multiverso_b200 as mv
    mv.FLAGS.reset()
    corpus, vocab, out = str(tmp_path / "c.txt"), str(tmp_path / "v.txt"), str(tmp_path / "vec.txt")
    _corpus(corpus)
    assert app.word_count(corpus, vocab, 1) == 500
    argv = (f"-train_file {corpus} -read_vocab {vocab} -output {out} -size 64 -window 2 -epoch 4 "
            f"-alpha 0.05 -min_count 1 -sample 0 -data_block_size 60000 {flags}").split()
    stats = app.run(argv)
    assert stats["words"] == 4 * 3000 * 24 and stats["vocab"] == 500
    lines = open(out).read().splitlines()
    assert lines[0] == "500 64" and len(lines) == 501
    emb = {l.split()[0]: np.array(l.split()[1:], dtype=np.float32) for l in lines[1:]}
    assert all(np.isfinite(v).all() for v in emb.values())
    # paired words end up closer than random pairs (input-embedding cosine)
    def cos(a, b):
        return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-9))
    paired = np.mean([cos(emb[f"w{2 * i}"], emb[f"w{2 * i + 1}"]) for i in range(100)])
    rand = np.mean([cos(emb[f"w{2 * i}"], emb[f"w{(2 * i + 101) % 500}"]) for i in range(100)])
    if "-cbow 0 -negative" in flags:
        assert paired > rand + 0.05, (paired, rand)


def _write_libsvm(path, X, y):
    with open(path, "w") as f:
        for xi, yi in zip(X, y):
            nz = np.nonzero(xi)[0]
            f.write(f"{int(yi)} " + " ".join(f"{k}:{xi[k]:.4f}" for k in nz) + "\n")


def _write_dense(path, X, y):
    with open(path, "w") as f:
        for xi, yi in zip(X, y):
            f.write(f"{int(yi)} " + " ".join(f"{v:.4f}" for v in xi) + "\n")


@pytest.mark.parametrize("use_ps", [False, True])
def test_logreg_dense_softmax(tmp_path, use_ps):
    _need_gpu()
    from multiverso_b200.apps.logreg import run
    import multiverso_b200 as mv
    mv.FLAGS.reset()
    rng = np.random.default_rng(1)
    C, D, n = 5, 40, 4000
    centers = rng.normal(size=(C, D)) * 2
    y = rng.integers(0, C, size=n)
    X = centers[y] + rng.normal(size=(n, D)) * 0.5
    tr, te = str(tmp_path / "tr.txt"), str(tmp_path / "te.txt")
    _write_dense(tr, X[:3000], y[:3000])
    _write_dense(te, X[3000:], y[3000:])
    cfg = tmp_path / "lr.config"
    cfg.write_text(f"input_size={D}\noutput_size={C}\nsparse=false\nobjective_type=softmax\nupdater_type=sgd\n"
                   f"train_epoch=3\nminibatch_size=50\nlearning_rate=0.1\ntrain_file={tr}\ntest_file={te}\n"
                   f"output_file={tmp_path}/out.txt\noutput_model_file={tmp_path}/model.bin\n"
                   f"use_ps={'true' if use_ps else 'false'}\nsync_frequency=2\npipeline=true\nregular_type=L2\nregular_coef=0.0001\n")
    stats = run(str(cfg))
    assert stats["test_error"] < 0.05, stats
    assert os.path.getsize(tmp_path / "model.bin") == (D + 1) * C * 4
    assert len(open(tmp_path / "out.txt").read().splitlines()) == 1000


@pytest.mark.parametrize("objective,server_updater", [("sigmoid", "sgd"), ("sigmoid", "adagrad"), ("ftrl", "sgd")])
def test_logreg_sparse(tmp_path, objective, server_updater):
    _need_gpu()
    from multiverso_b200.apps.logreg import run
    import multiverso_b200 as mv
    mv.FLAGS.reset()
    rng = np.random.default_rng(2)
    D, n = 2000, 6000
    wtrue = rng.normal(size=D)
    X = np.zeros((n, D), np.float32)
    for i in range(n):
        idx = rng.choice(D, size=20, replace=False)
        X[i, idx] = 1.0
    y = (X @ wtrue > 0).astype(np.int64)
    tr, te = str(tmp_path / "tr.svm"), str(tmp_path / "te.svm")
    _write_libsvm(tr, X[:5000], y[:5000])
    _write_libsvm(te, X[5000:], y[5000:])
    cfg = tmp_path / "lr.config"
    lr = 0.05 if server_updater == "adagrad" else 0.5
    cfg.write_text(f"input_size={D}\noutput_size=1\nsparse=true\nobjective_type={objective}\n"
                   f"updater_type={'ftrl' if objective == 'ftrl' else 'sgd'}\ntrain_epoch=4\nminibatch_size=20\n"
                   f"learning_rate={lr}\ntrain_file={tr}\ntest_file={te}\noutput_file=\noutput_model_file=\n"
                   f"use_ps=true\nserver_updater={server_updater}\nalpha=0.1\nbeta=1\nlambda1=0.01\nlambda2=0\n"
                   f"regular_type=default\n")
    stats = run(str(cfg))
    assert stats["test_error"] < (0.4 if objective == "ftrl" else 0.2), stats
