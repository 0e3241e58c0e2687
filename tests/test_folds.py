"""CPU: K-fold index assignment is bit-exact with the restatement of romcomma/data/storage.py:180-203 for equal `random` state."""
import random

import pytest

from oracle import folds as oracle_folds
from romcomma.data.storage import k_fold_indices


@pytest.mark.parametrize('N,K,shuffle', [(2048, 10, False), (2048, 10, True), (300, 2, False), (17, -5, True), (10, 10, False), (7, 1, True),
                                         (1, 1, False), (64, -64, False)])
def test_bit_exact_with_oracle(N, K, shuffle):
    random.seed(2)
    ours = k_fold_indices(N, K, shuffle)
    random.seed(2)
    ref = oracle_folds.into_K_folds(N, K, shuffle)
    assert ours == ref
    assert set(ours.keys()) == set(range(abs(K))) | ({K} if K > 0 else set())
    # partition property: every row is tested exactly once over the proper folds
    tested = sorted(i for k in range(abs(K)) for i in ours[k][1])
    assert tested == list(range(N))
    for k in range(abs(K)):
        train, test = ours[k]
        if abs(K) > 1:
            assert sorted(train + test) == list(range(N)) and not (set(train) & set(test))
        else:
            assert train == test          # a single fold trains on its test rows (storage.py:200)
    if K > 0:
        assert sorted(ours[K][0]) == list(range(N)) and ours[K][0] == ours[K][1]


def test_random_stream_consumption_is_identical():
    """After folding, both implementations leave Python's `random` in the same state (same number and order of shuffles)."""
    random.seed(11)
    k_fold_indices(101, 7, True)
    a = random.random()
    random.seed(11)
    oracle_folds.into_K_folds(101, 7, True)
    assert a == random.random()


@pytest.mark.parametrize('N,K', [(10, 0), (10, 11), (10, -11)])
def test_bad_K_raises(N, K):
    with pytest.raises(IndexError):
        k_fold_indices(N, K)
