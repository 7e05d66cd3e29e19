"""CPU tests of the native applications on the host runtime: build/bin/wordembedding,
build/bin/word_count and build/bin/logreg (reference: Applications/WordEmbedding and
Applications/LogisticRegression, which ship without tests; the scenarios here train small
synthetic problems with a known structure on 1-3 ranks and check that the structure is
learnt, that the output files have the reference's formats and that the parameter-server
paths agree with the local model)."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    sys.path.insert(0, ROOT)
    from multiverso_b200 import _build
    _build.build_host()
    for exe in ("wordembedding", "word_count", "logreg"):
        assert os.path.exists(os.path.join(BIN, exe))


def run(nproc, *cmd, timeout=120):
    """Run `cmd` on nproc ranks; returns the per-rank JSON result lines sorted by rank."""
    if nproc == 1:
        r = subprocess.run(list(cmd), capture_output=True, text=True, timeout=timeout)
    else:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", str(nproc),
                            "--timeout", str(timeout), "--", *cmd], capture_output=True, text=True,
                           timeout=timeout + 30)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    assert len(res) == nproc, r.stdout[-2000:]
    return sorted(res, key=lambda d: d["rank"])


# ------------------------------------------------------------------------------- wordembedding
TOPICS, PER_TOPIC = 12, 30


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    """Sentences draw all their words from one of 12 topics: topic-mates must embed together."""
    d = tmp_path_factory.mktemp("we")
    rng = np.random.default_rng(0)
    path = d / "corpus.txt"
    with open(path, "w") as f:
        for _ in range(6000):
            t = rng.integers(TOPICS)
            ws = rng.integers(PER_TOPIC, size=rng.integers(5, 16))
            f.write(" ".join(f"t{t}w{w}" for w in ws) + "\n")
    return d, str(path)


def load_embeddings(path, binary=False):
    with open(path, "rb") as f:
        V, D = map(int, f.readline().split())
        words, vecs = [], []
        for _ in range(V):
            if binary:
                w = b""
                while True:
                    c = f.read(1)
                    if c == b" ":
                        break
                    w += c
                vecs.append(np.frombuffer(f.read(4 * D), dtype=np.float32))
                assert f.read(1) == b"\n"
            else:
                parts = f.readline().split()
                w = parts[0]
                vecs.append(np.array(parts[1:], dtype=np.float32))
            words.append(w.decode())
        assert f.read() == b""
    v = np.stack(vecs)
    assert v.shape == (V, D)
    return words, v


def topic_separation(words, v):
    v = v / np.linalg.norm(v, axis=1, keepdims=True)
    topic = np.array([int(w[1:w.index("w")]) for w in words])
    S = v @ v.T
    same = topic[:, None] == topic[None, :]
    np.fill_diagonal(same, False)
    return float(S[same].mean()), float(S[topic[:, None] != topic[None, :]].mean())


def test_word_count_tool(corpus):
    d, path = corpus
    vocab = str(d / "vocab.txt")
    r = subprocess.run([os.path.join(BIN, "word_count"), "-train_file", path, "-save_vocab", vocab, "-min_count", "1"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    rows = [line.split() for line in open(vocab)]
    assert len(rows) == TOPICS * PER_TOPIC
    freq = [int(f) for _, f in rows]
    assert freq == sorted(freq, reverse=True)
    assert sum(freq) == sum(len(line.split()) for line in open(path))


@pytest.mark.parametrize("nproc,flags,binary", [
    (1, ["-cbow", "0", "-negative", "5"], False),                       # skip-gram, negative sampling
    (1, ["-cbow", "1", "-negative", "5", "-is_pipeline", "0", "-binary", "1"], True),
    (1, ["-cbow", "0", "-hs", "1"], False),                             # hierarchical softmax
    (1, ["-cbow", "1", "-use_adagrad", "1", "-alpha", "0.05"], False),  # AdaGrad tables
    (2, ["-cbow", "0", "-negative", "5"], False),                       # 2 ranks, pipelined block protocol
    (3, ["-cbow", "1", "-negative", "3", "-sample", "0.01", "-alpha", "0.1"], False),   # deltas are averaged over 3 workers
])
def test_wordembedding_learns_topics(corpus, nproc, flags, binary):
    d, path = corpus
    out = str(d / f"vec_{nproc}_{'_'.join(flags).replace('-', '')}.out")
    res = run(nproc, os.path.join(BIN, "wordembedding"), "-train_file", path, "-output", out, "-size", "24",
              "-epoch", "4", "-threads", "2", "-min_count", "1", "-data_block_size", "60000", *flags)
    assert sum(r["words"] for r in res) > 0
    for r in res:
        assert r["vocab"] == TOPICS * PER_TOPIC
        losses = r["epoch_loss"]
        assert losses[-1] < losses[0], losses            # the objective goes down on every rank
    words, v = load_embeddings(out, binary)
    assert len(words) == TOPICS * PER_TOPIC and v.shape[1] == 24
    same, other = topic_separation(words, v)
    assert same > other + 0.2, (same, other)


def test_wordembedding_vocab_file_and_min_count(corpus):
    d, path = corpus
    vocab = str(d / "vocab_mc.txt")
    subprocess.run([os.path.join(BIN, "word_count"), "-train_file", path, "-save_vocab", vocab], check=True,
                   capture_output=True, timeout=60)
    freq = sorted((int(line.split()[1]) for line in open(vocab)), reverse=True)
    cut = freq[len(freq) // 2]
    res = run(1, os.path.join(BIN, "wordembedding"), "-train_file", path, "-read_vocab", vocab, "-size", "8",
              "-min_count", str(cut), "-epoch", "1", "-data_block_size", "200000")
    assert res[0]["vocab"] == sum(1 for f in freq if f >= cut)


def test_wordembedding_usage():
    r = subprocess.run([os.path.join(BIN, "wordembedding")], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "-train_file" in r.stdout


# -------------------------------------------------------------------------------------- logreg
@pytest.fixture(scope="module")
def lr_data(tmp_path_factory):
    d = tmp_path_factory.mktemp("lr")
    rng = np.random.default_rng(1)
    D, C, N = 20, 4, 3600
    Wt = rng.normal(size=(C, D))
    X = rng.normal(size=(N, D))
    y = (X @ Wt.T + 0.1 * rng.normal(size=(N, C))).argmax(1)

    def dense(p, X, y):
        with open(p, "w") as f:
            for xi, yi in zip(X, y):
                f.write(str(int(yi)) + " " + " ".join("%.4f" % v for v in xi) + "\n")
    dense(d / "dense_train.txt", X[:3000], y[:3000])
    dense(d / "dense_test.txt", X[3000:], y[3000:])
    SD = 2000
    w = rng.normal(size=SD) * (rng.random(SD) < 0.3)

    def sparse(p, n, weighted=False):
        with open(p, "w") as f:
            for _ in range(n):
                ks = np.sort(rng.choice(SD, 10, replace=False))
                vs = rng.random(10) + 0.5
                lab = int((w[ks] * vs).sum() > 0)
                head = f"{lab}:1.0" if weighted else str(lab)
                f.write(head + " " + " ".join("%d:%.3f" % (k, v) for k, v in zip(ks, vs)) + "\n")
    sparse(d / "sp_train.txt", 12000)
    sparse(d / "sp_test.txt", 1500)
    sparse(d / "spw_train.txt", 3000, weighted=True)
    # bsparse: u64 count | i32 label | f64 weight | count x u64 keys
    with open(d / "bs_train.bin", "wb") as f:
        for _ in range(3000):
            ks = np.sort(rng.choice(SD, 10, replace=False)).astype(np.uint64)
            lab = int(w[ks.astype(np.int64)].sum() > 0)
            f.write(struct.pack("<Qid", len(ks), lab, 1.0) + ks.tobytes())
    return d


def write_config(d, name, **kv):
    base = dict(train_epoch=3, minibatch_size=10, show_time_per_sample=100000,
                output_file=str(d / f"{name}.out"), output_model_file=str(d / f"{name}.model"))
    base.update(kv)
    path = d / f"{name}.config"
    with open(path, "w") as f:
        f.write("# generated by tests/test_native_apps.py\n")
        for k, v in base.items():
            f.write(f"{k} = {v}\n")
    return str(path)


def dense_cfg(d, name, **kv):
    return write_config(d, name, input_size=20, output_size=4, objective_type="softmax", regular_type="L2",
                        regular_coef=0.0007, updater_type="sgd", learning_rate=0.5, sparse="false",
                        train_file=str(d / "dense_train.txt"), test_file=str(d / "dense_test.txt"), **kv)


def sparse_cfg(d, name, **kv):
    kv.setdefault("train_file", str(d / "sp_train.txt"))
    kv.setdefault("test_file", str(d / "sp_test.txt"))
    kv.setdefault("objective_type", "sigmoid")
    return write_config(d, name, input_size=2000, output_size=1, updater_type="sgd", learning_rate=0.5,
                        sparse="true", **kv)


def test_logreg_dense_softmax_local(lr_data):
    d = lr_data
    res = run(1, os.path.join(BIN, "logreg"), dense_cfg(d, "dense_local"))[0]
    assert res["epoch_loss"][-1] < res["epoch_loss"][0] and res["test_error"] < 0.1, res
    # predictions: one line per test sample, a probability per class
    rows = [list(map(float, line.split())) for line in open(d / "dense_local.out")]
    assert len(rows) == 600 and all(len(r) == 4 and abs(sum(r) - 1) < 1e-3 for r in rows)
    # dense model file = raw dump of out x (in + 1) floats
    assert os.path.getsize(d / "dense_local.model") == 4 * 4 * 21


def test_logreg_ps_equals_local_on_one_rank(lr_data):
    d = lr_data
    local = run(1, os.path.join(BIN, "logreg"), dense_cfg(d, "eq_local"))[0]
    ps = run(1, os.path.join(BIN, "logreg"), dense_cfg(d, "eq_ps", use_ps="true", pipeline="false"))[0]
    assert ps["epoch_loss"] == local["epoch_loss"] and ps["test_error"] == local["test_error"]
    assert open(d / "eq_local.model", "rb").read() == open(d / "eq_ps.model", "rb").read()


@pytest.mark.parametrize("nproc", [2, 3])
@pytest.mark.parametrize("pipeline", ["true", "false"])
def test_logreg_dense_ps(lr_data, nproc, pipeline):
    d = lr_data
    res = run(nproc, os.path.join(BIN, "logreg"),
              dense_cfg(d, f"dense_ps{nproc}{pipeline}", use_ps="true", pipeline=pipeline, sync_frequency=2))
    assert sum(r["samples"] for r in res) == 3 * 3000
    for r in res:
        assert r["test_error"] < 0.12, r
    # every worker writes its own prediction file
    for w in range(nproc):
        assert os.path.exists(d / f"dense_ps{nproc}{pipeline}.out-{w}")


@pytest.mark.parametrize("nproc,extra", [
    (1, {}),
    (2, dict(use_ps="true", pipeline="true", sync_frequency=3)),       # SparseTable, key-set pulls one window ahead
    (3, dict(use_ps="true", pipeline="false")),
    (2, dict(use_ps="true", objective_type="ftrl", alpha=0.1, beta=1, lambda1=0.1, lambda2=0)),   # FTRLTable
    (1, dict(objective_type="ftrl", alpha=0.1, beta=1, lambda1=0.1, lambda2=0)),
])
def test_logreg_sparse(lr_data, nproc, extra):
    d = lr_data
    name = f"sp{nproc}" + "".join(str(v)[:2] for v in extra.values())
    res = run(nproc, os.path.join(BIN, "logreg"), sparse_cfg(d, name, **extra))
    for r in res:
        assert r["epoch_loss"][-1] < r["epoch_loss"][0], r
        assert r["test_error"] < (0.3 if "objective_type" in extra else 0.25), r


def test_logreg_readers_weight_and_bsparse(lr_data):
    d = lr_data
    r = run(1, os.path.join(BIN, "logreg"),
            sparse_cfg(d, "spw", train_file=str(d / "spw_train.txt"), reader_type="weight", test_file=""))[0]
    assert r["samples"] == 3 * 3000 and r["epoch_loss"][-1] < r["epoch_loss"][0]
    r = run(1, os.path.join(BIN, "logreg"),
            sparse_cfg(d, "bs", train_file=str(d / "bs_train.bin"), reader_type="bsparse", test_file=""))[0]
    assert r["samples"] == 3 * 3000 and r["epoch_loss"][-1] < r["epoch_loss"][0]


@pytest.mark.parametrize("use_ps", ["false", "true"])
def test_logreg_init_model_file(lr_data, use_ps):
    """A model saved by one run seeds the next one (Model::Load / PSModel::Load through the servers)."""
    d = lr_data
    first = run(1, os.path.join(BIN, "logreg"), sparse_cfg(d, f"seed{use_ps}"))[0]
    nproc = 2 if use_ps == "true" else 1
    res = run(nproc, os.path.join(BIN, "logreg"),
              sparse_cfg(d, f"resume{use_ps}", init_model_file=str(d / f"seed{use_ps}.model"), train_epoch=1,
                         use_ps=use_ps))
    for r in res:
        assert r["epoch_loss"][0] < 0.8 * first["epoch_loss"][0], (r, first)


def test_logreg_usage_and_bad_config(tmp_path):
    r = subprocess.run([os.path.join(BIN, "logreg")], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "usage" in r.stdout
    bad = tmp_path / "bad.config"
    bad.write_text("output_size=2\n")
    r = subprocess.run([os.path.join(BIN, "logreg"), str(bad)], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "input_size" in r.stderr


# ------------------------------------------------------------------------------------ BSP mode
def test_logreg_dense_ps_bsp_equal_step_counts(lr_data):
    """-sync=true: every worker must issue the same number of Adds / Gets. 3000 samples / 10 per
    minibatch / 2 per window = 150 windows do not divide by 4 ranks: the incomplete last round is
    dropped, every rank trains on the same number of samples and nothing deadlocks."""
    d = lr_data
    res = run(4, os.path.join(BIN, "logreg"),
              dense_cfg(d, "dense_bsp", use_ps="true", pipeline="false", sync_frequency=2), "-sync=true")
    assert len({r["samples"] for r in res}) == 1 and res[0]["samples"] == 3 * (150 // 4) * 20
    for r in res:
        assert r["test_error"] < 0.12, r


def test_logreg_sparse_ps_rejects_bsp(lr_data):
    d = lr_data
    cfg = sparse_cfg(d, "sp_bsp", use_ps="true")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mvrun.py"), "-n", "2", "--timeout", "60", "--",
                        os.path.join(BIN, "logreg"), cfg, "-sync=true"], capture_output=True, text=True, timeout=90)
    assert r.returncode != 0 and "async mode" in (r.stdout + r.stderr)


def test_wordembedding_bsp_complete_rounds(corpus):
    d, path = corpus
    res = run(3, os.path.join(BIN, "wordembedding"), "-train_file", path, "-size", "16", "-epoch", "2", "-threads", "2",
              "-min_count", "1", "-data_block_size", "30000", "-cbow", "0", "-is_pipeline", "0", "-sync=true")
    assert len({r["blocks"] for r in res}) == 1 and res[0]["blocks"] > 0      # same number of blocks on every rank
    for r in res:
        assert r["epoch_loss"][-1] < r["epoch_loss"][0]
