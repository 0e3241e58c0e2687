"""Oracle for the K-fold index assignment (integers, bit-exact).  TEST INFRASTRUCTURE ONLY.

Restates romcomma/data/storage.py:180-203: an optional ``random.shuffle`` of range(N); floor(N/K) shuffled copies of
range(K) plus a shuffled range(N % K) form the indicator; fold k tests on {i : indicator_i == k} and trains on the rest
(or on the test rows if there are no others).  K > 0 adds the improper fold K (all rows train and test).
Uses Python's global ``random`` state exactly as the reference does, so equal seeds give equal folds.
"""
from __future__ import annotations

import itertools
import random
from typing import Dict, List, Tuple


def into_K_folds(N: int, K: int, shuffle_before_folding: bool = False) -> Dict[int, Tuple[List[int], List[int]]]:
    if not (1 <= abs(K) <= N):
        raise IndexError(f'K={K:d} does not lie between 1 and N={N:d} inclusive.')
    index = list(range(N))
    if shuffle_before_folding:
        random.shuffle(index)
    folds = {}
    if K > 0:
        folds[K] = (list(index), list(index))
    K = abs(K)
    blocks = [list(range(K)) for _ in range(int(N / K))]
    blocks.append(list(range(N % K)))
    for b in blocks:
        random.shuffle(b)
    indicator = list(itertools.chain(*blocks))
    for k in range(K):
        pairs = tuple(zip(index, indicator))
        train = [i for i, ind in pairs if k != ind]
        test = [i for i, ind in pairs if k == ind]
        folds[k] = (test if train == [] else train, test)
    return folds
