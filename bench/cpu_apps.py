"""CPU-only comparison of the native `wordembedding` application (host runtime, no GPU) with the
UNMODIFIED reference application (baseline/_ref/bin/wordembedding, built by
tools/build_reference.sh against the MPI shim) on the same machine, corpus and flags.

    python bench/cpu_apps.py [--words 3000000] [--vocab 1000000] [--threads N] [--out FILE]

Config = BASELINE config 3 (skip-gram, dim 300, vocab 1M, 5 negatives, window 5, synthetic Zipf
corpus).  Reference timed region: its "MV Barrier done." .. "Finish Training" log lines (excludes
its vocabulary load); ours: the whole run after MV_Init including the vocabulary load -- the
comparison is conservative.  This is the CPU plumbing mode, not the product path (bench.py is).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline.reference_runner import _write_corpus  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--words", type=int, default=3_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=300)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--work", default="/tmp/mv_cpu_apps")
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True)
    from multiverso_b200 import _build
    _build.build_host()
    corpus, vocab, nbytes = _write_corpus(a.work, 0, a.vocab, a.words)
    block = max(1 << 20, nbytes // 3 + 1)
    flags = ["-train_file", corpus, "-read_vocab", vocab, "-size", str(a.dim), "-cbow", "0", "-negative", "5",
             "-window", "5", "-epoch", "1", "-alpha", "0.025", "-threads", str(a.threads), "-min_count", "1",
             "-sample", "0", "-binary", "1", "-hs", "0", "-data_block_size", str(block),
             "-max_preload_data_size", str(8 * block), "-stopwords", "0", "-use_adagrad", "0", "-is_pipeline", "1"]
    res = {"config": {"words": a.words, "vocab": a.vocab, "dim": a.dim, "threads": a.threads, "negative": 5,
                      "window": 5, "cpus": os.cpu_count()}}
    p = subprocess.run([os.path.join(ROOT, "build", "bin", "wordembedding"), *flags, "-output",
                        os.path.join(a.work, "ours.bin"), f"-omp_threads={a.threads}"],
                       capture_output=True, text=True, cwd=a.work)
    ours = [json.loads(line) for line in p.stdout.splitlines() if line.startswith("{")][0]
    res["native"] = {"words_per_sec": ours["words"] / ours["seconds"], "seconds": ours["seconds"],
                     "train_seconds": ours["train_seconds"], "pull_seconds": ours["pull_seconds"],
                     "push_seconds": ours["push_seconds"], "loss": ours["epoch_loss"][-1]}
    ref_bin = os.path.join(ROOT, "baseline", "_ref", "bin", "wordembedding")
    if not a.skip_reference and os.path.exists(ref_bin):
        env = dict(os.environ, MV_SHIM_RANK="0", MV_SHIM_SIZE="1", OMP_NUM_THREADS=str(a.threads))
        t0 = t1 = None
        proc = subprocess.Popen([ref_bin, *flags, "-output", os.path.join(a.work, "ref.bin")], stdout=subprocess.PIPE,
                                stderr=subprocess.STDOUT, text=True, env=env, cwd=a.work)
        for line in proc.stdout:
            if t0 is None and "MV Barrier done" in line:
                t0 = time.time()
            if "Finish Training" in line:
                t1 = time.time()
        proc.wait()
        if t0 and t1:
            res["reference"] = {"words_per_sec": a.words / (t1 - t0), "seconds": t1 - t0}
            res["speedup"] = res["native"]["words_per_sec"] / res["reference"]["words_per_sec"]
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
