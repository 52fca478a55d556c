"""Host logic of NnetCtcUpdater::ComputeTotAccuracy (ctc-nnet-update.cc:261-314): the product's
collapse + edit distance against the oracle's restatement, and known answers."""
import numpy as np

from kaldi_ctc_b200 import nnet
from oracle import pyoracle


def test_collapse_keeps_frame_zero_even_if_blank():
    assert nnet.collapse_best_path(np.array([0, 0, 3, 3, 0, 3, 4])) == [0, 3, 3, 4]
    assert nnet.collapse_best_path(np.array([2, 2, 2])) == [2]
    assert nnet.collapse_best_path(np.array([0])) == [0]


def test_levenshtein_known_answers():
    assert nnet.levenshtein([1, 2, 3], [1, 2, 3]) == 0
    assert nnet.levenshtein([], [4, 5]) == 2
    assert nnet.levenshtein([1, 2, 3, 4], [2, 3, 5]) == 2
    assert nnet.levenshtein(list("kitten"), list("sitting")) == 3


def test_tot_accuracy_matches_oracle_on_random_outputs():
    rng = np.random.default_rng(11)
    for trial in range(20):
        B, A = int(rng.integers(1, 6)), int(rng.integers(3, 9))
        il = rng.integers(1, 30, size=B)
        T = int(il.max())
        il[rng.integers(0, B)] = T
        ll = np.array([rng.integers(1, max(2, t // 2 + 1)) for t in il])
        fl = rng.integers(1, A, size=int(ll.sum()))
        out = rng.standard_normal((T * B, A)).astype(np.float32)
        out[rng.integers(0, T * B, size=T), 0] += 3.0  # some blanks win
        want = pyoracle.tot_accuracy(out, fl, ll, il, B)
        got = nnet.tot_accuracy(out.argmax(1).reshape(T, B), fl, ll, il)
        assert got == want
