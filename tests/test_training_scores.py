"""Row f3: the tokenization loop of the training worker (training/trainvocab.go:932-1170) --
per-token "characters covered" histogram, tokensInText and the bytes without a token --
against the oracle's restatement (oracle/tm_oracle.c tmo_score)."""
import os

import numpy as np
import pytest

from oracle import oracle
from workloads import configs, corpus, vocabgen


def flat(texts):
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    return np.frombuffer(b"".join(texts), np.uint8), off


def test_oracle_score_invariants():
    """Every normalized character is covered exactly once, a retry branch adds one for its
    delete token (which covers nothing), and tokensInText equals the id count of Vocab.Tokenize
    when nothing is missing."""
    vb = configs.vocab_bytes("c2")
    o = oracle.OracleVocab(vb, "port")
    t = bytes(corpus.base_text(3, 50000, 0.3))
    scores, hist, tit = o.score([t])
    ids, miss = o.tokenize(t)
    assert miss == 0 and hist.sum() == 0 and tit == len(ids)
    n_del = int((ids == o.delete_token).sum())          # literal "D" markers + retry branches
    extra = int(scores.sum()) - len(o.normalize(t))      # = retry branches
    assert 0 < extra <= n_del and int(scores[o.delete_token]) == n_del


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c1"])
def test_training_scores_against_oracle(built, name):
    from tokenmonster_b200 import Vocab
    vb = configs.vocab_bytes(name)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    base = corpus.base_text(77, 1 << 19, 0.4)
    d, o = corpus.make_batch(base, 3000, 0, 700, 5, "cpu")
    d, o = d.numpy(), o.numpy().astype(np.uint64)
    texts = [d[int(o[s]):int(o[s + 1])].tobytes() for s in range(len(o) - 1)] + [bytes(corpus.base_text(5, 100_000, 0.3))]
    data, off = flat(texts)
    scores, hist, tit, status = v.score_batch(data, off)
    want_s, want_h, want_t = chk.score(texts)
    assert not status.any()
    assert tit == want_t
    assert np.array_equal(scores, want_s[:len(scores)])
    assert np.array_equal(hist, want_h)
    v.close()


@pytest.mark.gpu
def test_training_scores_missing_bytes_and_non_ascii(built):
    """A vocabulary that does not cover every byte (the worker's missingList) and strings with
    non-ASCII text (set aside before the first pass so that nothing is scored twice)."""
    from oracle import unicode_oracle as uo
    from tokenmonster_b200 import Vocab
    toks = [b" the", b"the", b" a", b"an", b" and", b"nd", b" qu", b"ick", b" brown", b"fox", b"o", b"x", b"D", b"C", b"W",
            b" ", b"e", b"t", b"h", b"\xcc", b"\x81", b"c", b"a", b"f"]
    vb = vocabgen.generate_vocab(toks, capcode=2, charset=1, normflag=1, level=2, reserve=32)
    v, chk = Vocab(vb, device=0), oracle.OracleVocab(vb, "port")
    rng = np.random.default_rng(3)
    words = [b"the", b" quick", b" brown", b" fox", b" and", b" an", b"?", b" Z", b"\n", b".", b" The", b" FOX"]
    texts = [b"".join(words[k] for k in rng.integers(0, len(words), int(m))) for m in rng.integers(0, 80, 500)]
    texts += ["the café and thé".encode(), b"\xff bad", "É fox".encode()]
    data, off = flat(texts)
    scores, hist, tit, status = v.score_batch(data, off)
    want_s = np.zeros_like(scores)
    want_h = np.zeros(256, np.uint64)
    want_t = 0
    for t in texts:
        try:
            norm = uo.normalize(t, 1, 2) if any(b >= 0x80 for b in t) else chk.normalize(t)
        except uo.Refused:
            continue
        s1, h1, t1 = chk.score([norm], normalized=True)
        want_s += s1[:len(want_s)]
        want_h += h1
        want_t += t1
    assert status.tolist() == [0] * (len(texts) - 2) + [13, 0]
    assert int(hist.sum()) > 100 and tit == want_t
    assert np.array_equal(scores, want_s) and np.array_equal(hist, want_h)
    v.close()
