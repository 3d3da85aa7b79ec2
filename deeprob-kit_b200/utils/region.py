"""Host-side region graph (interface of deeprob/utils/region.py:11-99).

The region graph *is* the model structure (it becomes the leaf gather table), so it has to be
reproduced bit-exactly: the same `RandomState.permutation` calls are issued in the same order
(one per region, breadth-first inside a repetition, repetitions in sequence) and each half is
sorted -- see region.py:64-74 and :91-97.  tests/test_region_graph.py pins this against tables
dumped from the reference.
"""
from typing import List, Optional

import numpy as np

from .random import RandomState, check_random_state


class RegionGraph:
    def __init__(self, n_features: int, depth: int, random_state: Optional[RandomState] = None):
        if n_features <= 0:
            raise ValueError("The number of features must be positive")
        if depth <= 0:
            raise ValueError("The region graph depth must be positive")
        if depth > int(np.log2(n_features)):
            raise ValueError("Invalid region graph depth based on the number of features")
        self.items = tuple(range(n_features))
        self.depth = depth
        self.random_state = check_random_state(random_state)

    def _split(self, region: tuple):
        shuffled = self.random_state.permutation(region).tolist()
        cut = len(region) // 2
        return tuple(sorted(shuffled[:cut])), tuple(sorted(shuffled[cut:]))

    def random_layers(self) -> List[List[tuple]]:
        """One repetition: [root regions, partitions, regions, ..., leaf regions]."""
        out = [[self.items]]
        frontier = [self.items]
        for _ in range(self.depth):
            halves = [self._split(r) for r in frontier]
            frontier = [h for pair in halves for h in pair]
            out.append(halves)
            out.append(frontier)
        return out

    def make_layers(self, n_repetitions: int = 1) -> List[List[tuple]]:
        """Level-wise concatenation of `n_repetitions` independent repetitions."""
        if n_repetitions <= 0:
            raise ValueError("The number of repetitions must be positve")
        merged: List[list] = [[self.items]] + [[] for _ in range(2 * self.depth)]
        for _ in range(n_repetitions):
            rep = self.random_layers()
            for level in range(1, len(rep)):
                merged[level] = merged[level] + rep[level]
        return merged
