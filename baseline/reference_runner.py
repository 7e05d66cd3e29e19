"""Runs the UNMODIFIED reference WordEmbedding application (built by tools/build_reference.sh
from /root/reference against the MPI shim) on the benchmark config and reports words/sec.

Config = BASELINE.json config 3: skip-gram, dim 300, vocab 1M, 5 negatives, window 5,
synthetic Zipf corpus, one corpus shard per rank (weak scaling), all CPU cores split over the
ranks.  Timed region = wall clock between the reference's own "MV Barrier done." log line
(vocabulary loaded, tables created) and "Finish Training" (last block pushed), max over ranks.
The reference is a CPU program: there is no device in its path, its number is end to end.
"""
from __future__ import annotations

import json
import os
import subprocess
import time

import numpy as np

REF_WORDS_PER_RANK = int(os.environ.get("MV_REF_WORDS", 3_000_000))


def _write_corpus(dirname: str, rank: int, vocab: int, n_words: int):
    rng = np.random.default_rng(17 + rank)
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = 1.0 / ranks
    cdf = np.cumsum(p / p.sum())
    ids = np.minimum(np.searchsorted(cdf, rng.random(n_words), side="right"), vocab - 1)
    words = np.char.add("w", np.arange(vocab).astype(str))
    corpus = os.path.join(dirname, f"corpus_{rank}.txt")
    with open(corpus, "w") as f:
        for s in range(0, n_words, 1000):
            f.write(" ".join(words[ids[s:s + 1000]]) + "\n")
    # the same vocabulary file on every rank: expected Zipf counts (>= 1 so nothing is dropped)
    vocab_file = os.path.join(dirname, f"vocab_{rank}.txt")
    counts = np.maximum((n_words * p / p.sum()).astype(np.int64), 1)
    with open(vocab_file, "w") as f:
        f.write("\n".join(f"w{i} {c}" for i, c in enumerate(counts)) + "\n")
    return corpus, vocab_file, os.path.getsize(corpus)


def run_wordembedding(ref_bin: str, args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    work = os.environ.get("MV_REF_WORKDIR", "/tmp/mv_ref_arm")
    os.makedirs(work, exist_ok=True)
    n_words = REF_WORDS_PER_RANK
    corpus, vocab_file, nbytes = _write_corpus(work, rank, args.vocab, n_words)
    threads = max(1, (os.cpu_count() or 8) // world)
    block_bytes = max(1 << 20, nbytes // 3 + 1)           # three data blocks per rank
    cmd = [ref_bin, "-train_file", corpus, "-read_vocab", vocab_file, "-output", os.path.join(work, f"vec_{rank}.bin"),
           "-size", str(args.dim), "-cbow", "0", "-negative", str(args.negative), "-window", str(args.window),
           "-epoch", "1", "-alpha", "0.025", "-threads", str(threads), "-min_count", "1", "-sample", "0",
           "-binary", "1", "-hs", "0", "-data_block_size", str(block_bytes), "-max_preload_data_size",
           str(8 * block_bytes), "-stopwords", "0", "-use_adagrad", "0", "-is_pipeline", "1"]
    env = dict(os.environ, MV_SHIM_RANK=str(rank), MV_SHIM_SIZE=str(world), OMP_NUM_THREADS=str(threads))
    t_start = t_end = None
    t_launch = time.time()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=work)
    tail = []
    for line in p.stdout:
        now = time.time()
        tail.append(line.rstrip())
        tail = tail[-8:]
        if t_start is None and "MV Barrier done" in line:
            t_start = now
        if "Finish Training" in line:
            t_end = now
    rc = p.wait()
    if rc != 0 or t_start is None or t_end is None:
        raise RuntimeError(f"reference binary rc={rc}; tail: {' | '.join(tail)[-400:]}")
    mine = {"rank": rank, "train_s": t_end - t_start, "total_s": time.time() - t_launch, "words": n_words}
    with open(os.path.join(work, f"result_{rank}.json"), "w") as f:
        json.dump(mine, f)
    if rank != 0:
        return {"impl": "reference", "rank": rank, "note": "aggregated by rank 0"}
    results = [mine]
    deadline = time.time() + 1800
    for r in range(1, world):
        path = os.path.join(work, f"result_{r}.json")
        while not os.path.exists(path) and time.time() < deadline:
            time.sleep(0.2)
        time.sleep(0.05)
        results.append(json.load(open(path)))
    for r in range(world):
        try:
            os.remove(os.path.join(work, f"result_{r}.json"))
        except OSError:
            pass
    t = max(x["train_s"] for x in results)
    total_words = sum(x["words"] for x in results)
    value = total_words / t
    return {
        "metric": "wordembedding_words_per_sec", "value": value, "unit": "words/s", "n_gpus": world,
        "steps": 3, "warmup": 0, "ms_per_step": t * 1e3 / 3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "impl": "reference",
        "config": {"model": "WordEmbedding skip-gram dim=300 vocab=1M neg=5 window=5 (synthetic Zipf corpus)",
                   "global_batch": n_words * world // 3, "seq_len": 1000,
                   "parallelism": f"{world} CPU process(es) x {threads} OpenMP threads, reference PS over the MPI shim",
                   "words_per_rank": n_words,
                   "steps_note": "the reference CLI has no step count: a step here is one of its data blocks "
                                 "(-data_block_size = corpus/3), the whole corpus is timed",
                   "note": "unmodified /root/reference sources (core + Applications/WordEmbedding) compiled against "
                           "baseline/mpi_shim; CPU only -- the reference has no GPU code"},
        "e2e": {"value": value, "unit": "words/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "CPU program reading its corpus from disk: the number is end to end by construction"},
        "gpu_launches": 0,
    }


def run_matrix_bw(ref_bin: str, rows: int = 1_000_000, cols: int = 512, iters: int = 2) -> dict:
    """BASELINE.json config 2 on the UNMODIFIED reference: MatrixTable rows x cols fp32, whole-table Add
    (server-side sgd updater) and whole-table Get through the reference's public C++ API
    (baseline/ref_matrix_bw_main.cpp, the calls of Test/test_matrix_perf.cpp), one process per rank over
    the MPI shim.  Table bytes / wall time, max over ranks."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    work = os.environ.get("MV_REF_WORKDIR", "/tmp/mv_ref_arm")
    os.makedirs(work, exist_ok=True)
    threads = max(1, (os.cpu_count() or 8) // world)
    env = dict(os.environ, MV_SHIM_RANK=str(rank), MV_SHIM_SIZE=str(world), OMP_NUM_THREADS=str(threads))
    p = subprocess.run([ref_bin, str(rows), str(cols), str(iters), f"-omp_threads={min(threads, 64)}"],
                       capture_output=True, text=True, env=env, cwd=work, timeout=1700)
    mine = None
    for line in p.stdout.splitlines():
        if line.startswith("{"):
            mine = json.loads(line)
    if p.returncode != 0 or mine is None:
        raise RuntimeError(f"reference matrix_bw rc={p.returncode}: {(p.stdout + p.stderr)[-300:]}")
    with open(os.path.join(work, f"bw_result_{rank}.json"), "w") as f:
        json.dump(mine, f)
    if rank != 0:
        return {"impl": "reference", "rank": rank, "note": "aggregated by rank 0"}
    results = [mine]
    deadline = time.time() + 1800
    for r in range(1, world):
        path = os.path.join(work, f"bw_result_{r}.json")
        while not os.path.exists(path) and time.time() < deadline:
            time.sleep(0.2)
        time.sleep(0.05)
        results.append(json.load(open(path)))
    for r in range(world):
        try:
            os.remove(os.path.join(work, f"bw_result_{r}.json"))
        except OSError:
            pass
    add_ms = max(x["add_ms"] for x in results)
    get_ms = max(x["get_ms"] for x in results)
    nbytes = rows * cols * 4
    return {"metric": "matrix_table_get_plus_add_gbs", "value": 2 * nbytes / (add_ms + get_ms) / 1e6, "unit": "GB/s",
            "higher_is_better": True, "add_ms": add_ms, "get_ms": get_ms, "add_gbs": nbytes / add_ms / 1e6,
            "get_gbs": nbytes / get_ms / 1e6, "iters": iters, "verified": all(x["verified"] for x in results),
            "config": {"table": f"MatrixTable {rows}x{cols} fp32, whole-table Add (sgd updater) + whole-table Get",
                       "parallelism": f"{world} CPU process(es) x {threads} threads, reference PS over the MPI shim",
                       "timing": "reference Timer (wall clock) around blocking Add / Get, max over ranks"}}
