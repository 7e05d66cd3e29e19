"""The ``wordembedding`` application (reference: Applications/WordEmbedding).

    python -m multiverso_b200.apps.wordembedding -train_file corpus.txt -read_vocab vocab.txt \
        -output vec.txt -size 300 -cbow 0 -negative 5 -window 5 -epoch 1 -alpha 0.025 ...
    torchrun --nproc-per-node 8 -m multiverso_b200.apps.wordembedding ...      (8 GPUs)

Same 21 flags as the reference (util.cpp:31-56).  Pipeline, per rank:
loader thread (native reader: tokenise, dictionary, stop words, sub-sampling) -> bounded
BlockQueue of pinned token blocks (``-max_preload_data_size``) -> H2D -> K7 training kernel on
the HBM tables (block protocol across GPUs) -> word-count KV table -> lr decay;
rank 0 saves the embeddings (text or binary word2vec format) at the end.
"""
from __future__ import annotations

import ctypes as C
import queue
import sys
import threading
import time
from typing import List, Optional

import numpy as np

from ..utils import Log
from ._applib import lib

# The CPU implementation of this application is the native binary build/bin/wordembedding
# (csrc/host/apps/wordembedding); run() execs it when there is no GPU.
FLAG_HELP = """-size <int> -train_file <file> -endpoints_file <file> -read_vocab <file> -binary <0|1|2>
-cbow <0|1> -alpha <float> -output <file> -window <int> -sample <float> -hs <0|1>
-data_block_size <bytes> -max_preload_data_size <bytes> -negative <int> -threads <int>
-min_count <int> -epoch <int> -stopwords <0|1> -sw_file <file> -use_adagrad <0|1> -is_pipeline <0|1>"""


def parse_args(argv: List[str]):
    """``-flag value`` pairs exactly like Option::ParseArgs (util.cpp:31-56)."""
    from ..models.wordembedding import WordEmbeddingOption
    o = WordEmbeddingOption()
    o.data_block_size = 1 << 20
    binary = 0
    m = {"-size": ("embeding_size", int), "-train_file": ("train_file", str),
         "-endpoints_file": ("endpoints_file", str), "-read_vocab": ("read_vocab_file", str),
         "-cbow": ("cbow", lambda v: bool(int(v))), "-alpha": ("init_learning_rate", float),
         "-output": ("output_file", str), "-window": ("window_size", int), "-sample": ("sample", float),
         "-hs": ("hs", lambda v: bool(int(v))), "-data_block_size": ("data_block_size", int),
         "-max_preload_data_size": ("max_preload_data_size", int), "-negative": ("negative_num", int),
         "-threads": ("thread_cnt", int), "-min_count": ("min_count", int), "-epoch": ("epoch", int),
         "-stopwords": ("stopwords", lambda v: bool(int(v))), "-sw_file": ("sw_file", str),
         "-use_adagrad": ("use_adagrad", lambda v: bool(int(v))),
         "-is_pipeline": ("is_pipeline", lambda v: bool(int(v)))}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-binary" and i + 1 < len(argv):
            binary = int(argv[i + 1])
            i += 2
        elif a in m and i + 1 < len(argv):
            name, conv = m[a]
            setattr(o, name, conv(argv[i + 1]))
            i += 2
        else:
            i += 1
    o.output_binary = binary > 0
    return o


class Dictionary:
    """Native dictionary (dictionary.cpp): word <-> id, frequencies, min_count filter."""

    def __init__(self, vocab_file: Optional[str], train_file: Optional[str], min_count: int):
        L = lib()
        if vocab_file:
            self.h = L.MVA_DictLoad(vocab_file.encode(), min_count)
        else:
            self.h = L.MVA_DictFromCorpus(train_file.encode(), min_count)
        if not self.h:
            Log.fatal("cannot build the dictionary (vocab=%s corpus=%s)", vocab_file, train_file)
        self.size = L.MVA_DictSize(self.h)
        self.total_words = L.MVA_DictTotalWords(self.h)
        self.counts = np.zeros(self.size, dtype=np.int64)
        L.MVA_DictCounts(self.h, self.counts.ctypes.data_as(C.c_void_p))

    def word(self, i: int) -> str:
        return lib().MVA_DictWord(self.h, i).decode(errors="replace")

    def index(self, w: str) -> int:
        return lib().MVA_DictIndex(self.h, w.encode())

    def words(self) -> List[str]:
        return [self.word(i) for i in range(self.size)]


class BlockLoader(threading.Thread):
    """Loader thread + BlockQueue (distributed_wordembedding.cpp:33-56, block_queue.cpp): reads
    token blocks of ``block_tokens`` ids; ranks take blocks round-robin (block i -> rank i % size);
    the queue is bounded by ``max_preload`` bytes."""

    def __init__(self, dictionary: Dictionary, opt, rank: int, size: int, block_tokens: int, epochs: int):
        super().__init__(daemon=True)
        L = lib()
        sw = opt.sw_file if (opt.stopwords and opt.sw_file) else ""
        self.c = L.MVA_CorpusOpen(dictionary.h, opt.train_file.encode(), sw.encode(), float(opt.sample),
                                  C.c_uint64(12345 + rank))
        if not self.c:
            Log.fatal("cannot open %s", opt.train_file)
        self.rank, self.size, self.block_tokens, self.epochs = rank, size, block_tokens, epochs
        depth = max(2, int(opt.max_preload_data_size // max(1, block_tokens * 4)))
        self.q: "queue.Queue" = queue.Queue(maxsize=min(depth, 64))

    def run(self):
        import torch
        L = lib()
        i = 0                                       # blocks are dealt round-robin across epoch boundaries
        for ep in range(self.epochs):
            L.MVA_CorpusReset(self.c)
            while True:
                buf = torch.empty(self.block_tokens, dtype=torch.int32).pin_memory() \
                    if torch.cuda.is_available() else torch.empty(self.block_tokens, dtype=torch.int32)
                words = C.c_int64(0)
                n = L.MVA_CorpusNextBlock(self.c, C.c_void_p(buf.data_ptr()), self.block_tokens, C.byref(words))
                if n <= 0:
                    break
                if i % self.size == self.rank:
                    self.q.put((buf[:n], int(words.value), ep))
                i += 1
        self.q.put(None)
        L.MVA_CorpusClose(self.c)


def run(argv: List[str]) -> dict:
    import torch
    import multiverso_b200 as mv
    from ..models.wordembedding import WordEmbedding

    opt = parse_args(argv)
    if not opt.train_file:
        print("usage: wordembedding " + FLAG_HELP)
        return {}
    if not torch.cuda.is_available():
        # no GPU: the native CPU implementation of the same application (csrc/host/apps/wordembedding)
        from ._applib import run_native
        return run_native("wordembedding", argv)
    mv.init()
    rank, size = mv.rank(), mv.size()
    t0 = time.time()
    dictionary = Dictionary(opt.read_vocab_file, opt.train_file, opt.min_count)
    opt.total_words = dictionary.total_words
    Log.info("vocabulary %d words, %d corpus words, dim %d, %s %s", dictionary.size, dictionary.total_words,
             opt.embeding_size, "cbow" if opt.cbow else "skip-gram",
             "hs" if opt.hs else f"negative={opt.negative_num}")
    we = WordEmbedding(opt, dictionary.size, dictionary.counts.astype(np.float64))
    # data_block_size is in corpus BYTES in the reference; ~6 bytes per English token
    block_tokens = max(1024, int(opt.data_block_size // 6))
    loader = BlockLoader(dictionary, opt, rank, size, block_tokens, opt.epoch)
    loader.start()
    dev = torch.device("cuda", torch.cuda.current_device())
    tok_dev = [torch.empty(block_tokens, dtype=torch.int32, device=dev) for _ in range(2)]
    words_done, blocks, last_log = 0, 0, time.time()
    pipelined = bool(opt.is_pipeline) and size > 1

    def stage(item, slot):
        toks, words, ep = item
        buf = tok_dev[slot][: toks.numel()]
        buf.copy_(toks, non_blocking=True)                  # H2D of block i+1 overlaps training of block i
        return buf, words, ep

    item = loader.q.get()
    cur = stage(item, 0) if item is not None else None
    while cur is not None:
        item = loader.q.get()
        nxt = stage(item, (blocks + 1) & 1) if item is not None else None
        buf, words, ep = cur
        we.update_learning_rate()
        # -is_pipeline (default on, util.cpp:25): the next block's parameters are requested while
        # this one trains (distributed_wordembedding.cpp:199-222)
        we.train_block(buf, compute_loss=(blocks % 50 == 0),
                       next_tokens=nxt[0] if (pipelined and nxt is not None) else None)
        we.add_word_count(words * 1)                        # AddDeltaWordCount -> KV table
        words_done += words
        blocks += 1
        if blocks % 20 == 0:
            we.get_word_count()                             # GetAllWordCount: global progress -> lr decay
            if rank == 0 and time.time() - last_log > 5:
                el = time.time() - t0
                Log.info("Epoch %d  Words/sec %.0fk  lr %.6f  progress %.2f%%", ep, words_done / el / 1e3,
                         we.learning_rate, 100.0 * we.word_count_actual / max(1, opt.total_words * opt.epoch))
                last_log = time.time()
        cur = nxt
    we.flush()
    torch.cuda.synchronize()
    mv.barrier()
    elapsed = time.time() - t0
    if opt.output_file:
        we.save_embedding(opt.output_file, dictionary.words() if rank == 0 else None, binary=opt.output_binary)
        # writing a large vocabulary as text takes minutes on rank 0: the other ranks wait on the control
        # plane (no watchdog) so that shutdown's device barrier finds everybody already there
        mv.runtime.Runtime.get().host_barrier()
    stats = {"words": words_done, "seconds": elapsed, "words_per_sec": words_done / max(elapsed, 1e-9),
             "vocab": dictionary.size, "rank": rank}
    if rank == 0:
        Log.info("trained %d words in %.2fs on this rank (%.0f words/s/rank, %d ranks)", words_done, elapsed,
                 stats["words_per_sec"], size)
    mv.shutdown()
    return stats


def word_count(train_file: str, out_vocab: str, min_count: int = 0) -> int:
    """The ``word_count`` preprocessing tool (preprocess/word_count.cpp:30-46)."""
    return int(lib().MVA_WordCount(train_file.encode(), out_vocab.encode(), int(min_count)))


if __name__ == "__main__":
    run(sys.argv[1:])
