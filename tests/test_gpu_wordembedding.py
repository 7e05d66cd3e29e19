"""K7 numerics vs a plain PyTorch fp32 reference of the same sample schedule, plus
convergence checks for every mode (skip-gram / CBOW x NS / HS, AdaGrad)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sig(x):
    return 1.0 / (1.0 + torch.exp(-x))


@pytest.mark.parametrize("variant", [5, 1, 10])
def test_sgns_fast_kernel_matches_reference(mv_device, variant):
    """window=1 (no random shrink) and a one-word negative pool make the sample schedule
    deterministic: centre p trains (ctx p-1 -> p) then (ctx p+1 -> p), each with K draws of
    the pool word evaluated on pre-update rows. All words distinct, so cross-warp coupling
    is only through red.add accumulation: warps may or may not observe each other's updates
    (Hogwild), an O(lr^2) effect, so the test uses a small lr and compares the UPDATES with a
    relative tolerance (a wrong sign / missing term would be off by >= 100%)."""
    import ctypes as C
    from multiverso_b200 import _native as N
    V, D, K, lr = 64, 300, 5, 1e-3
    torch.manual_seed(0)
    w_in = (torch.rand(V, D, device="cuda") - 0.5) * 0.2
    w_out = (torch.rand(V, D, device="cuda") - 0.5) * 0.2
    toks = torch.tensor([3, 9, 27, 5, -1, 11, 2, -1, 40], dtype=torch.int32, device="cuda")
    pool = torch.tensor([63], dtype=torch.int32, device="cuda")
    ref_in, ref_out = w_in.double().cpu().clone(), w_out.double().cpu().clone()
    i0, o0 = ref_in.clone(), ref_out.clone()
    tl = toks.cpu().tolist()
    for p, c in enumerate(tl):
        if c < 0:
            continue
        crow = o0[c].clone()
        for q in (p - 1, p + 1):
            if q < 0 or q >= len(tl) or tl[q] < 0:
                continue
            h = i0[tl[q]]
            herr = torch.zeros(D, dtype=torch.float64)
            g = (1 - _sig(h @ crow)) * lr
            herr += g * crow
            ref_out[c] += g * h
            crow = crow + g * h
            nrow = o0[63]
            gn = (0 - _sig(h @ nrow)) * lr
            herr += K * gn * nrow
            ref_out[63] += K * gn * h
            ref_in[tl[q]] += herr
    a = N.Sgns()
    loss = torch.zeros(1, device="cuda")
    pairs = torch.zeros(1, dtype=torch.int64, device="cuda")
    a.tokens, a.n_tokens = toks.data_ptr(), toks.numel()
    a.w_in, a.w_out, a.dim, a.ld = w_in.data_ptr(), w_out.data_ptr(), D, D
    a.window, a.negative, a.lr = 1, K, lr
    a.vocab, a.neg_pool, a.neg_pool_size = V, pool.data_ptr(), 1
    a.seed, a.loss_sum, a.pair_count = 1234, loss.data_ptr(), pairs.data_ptr()
    a.variant = variant
    N.check(N.cuda_lib().mvb_sgns_train(C.byref(a), C.c_void_p(N.stream_ptr())))
    torch.cuda.synchronize()
    assert int(pairs.item()) == 8
    d_in, d_out = w_in.double().cpu() - i0, w_out.double().cpu() - o0
    r_in, r_out = ref_in - i0, ref_out - o0
    assert r_in.abs().max() > 1e-5 and r_out.abs().max() > 1e-5
    e_in = ((d_in - r_in).norm() / r_in.norm()).item()
    e_out = ((d_out - r_out).norm() / r_out.norm()).item()
    assert e_in < 0.05 and e_out < 0.05, (e_in, e_out)
    assert float(loss.item()) > 0


@pytest.mark.parametrize("mode", ["sg_ns", "cbow_ns", "sg_hs", "cbow_hs", "sg_ns_adagrad", "sg_ns_d100", "sg_ns_tma"])
def test_wordembedding_loss_decreases(mv_device, mode):
    from multiverso_b200.models.wordembedding import (WordEmbedding, WordEmbeddingOption,
                                                      synthetic_zipf_corpus)
    V = 2000
    opt = WordEmbeddingOption(embeding_size=100 if mode.endswith("d100") else 64, window_size=5, negative_num=5,
                              cbow="cbow" in mode, hs="hs" in mode, use_adagrad="adagrad" in mode,
                              init_learning_rate=0.025 if "hs" in mode else 0.05)
    we = WordEmbedding(opt, V)
    if mode.endswith("tma"):
        we.kernel_variant = 10
    # a learnable corpus: word 2i is always followed by word 2i+1
    rng = np.random.default_rng(0)
    base = rng.integers(0, V // 2, size=40000) * 2
    toks = np.stack([base, base + 1], 1).reshape(-1).astype(np.int32)
    toks[100::101] = -1
    toks = torch.from_numpy(toks).cuda()
    losses = []
    for it in range(6):
        we.loss.zero_(); we.pairs.zero_()
        we.train_block(toks)
        torch.cuda.synchronize()
        losses.append(float(we.loss.item()) / max(int(we.pairs.item()), 1))
    assert losses[-1] < losses[0] * 0.9, losses
    emb = we.embeddings()
    assert torch.isfinite(emb).all()


def test_synthetic_corpus_shape():
    from multiverso_b200.models.wordembedding import synthetic_zipf_corpus
    c = synthetic_zipf_corpus(10000, 1000, sentence_len=100)
    assert c.shape == (10000,) and c.max() < 1000 and (c == -1).sum() == 99
