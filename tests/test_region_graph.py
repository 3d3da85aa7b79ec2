"""CPU: host-side region graph of the product package (bit-exact structure + the reference's own checks,
tests/test_ratspn.py:24-43 of deeprob-kit)."""
import numpy as np
import pytest

from conftest import load_golden
from deeprob_kit_b200.utils.region import RegionGraph


def test_matches_reference_tables():
    gold = load_golden("region_graph")
    for key, tab in gold.items():
        _, d, depth, reps, seed = key.split("_")
        leaf = RegionGraph(int(d), int(depth), int(seed)).make_layers(int(reps))[-1]
        assert len(leaf) == tab.shape[0]
        for reg, row in zip(leaf, tab):
            assert tuple(int(v) for v in row[row >= 0]) == tuple(reg)


def test_structure_properties():
    n_features, depth, reps = 15, 2, 2
    layers = RegionGraph(n_features, depth=depth, random_state=42).make_layers(n_repetitions=reps)
    assert layers[0] == [tuple(range(n_features))]
    assert len(layers) == 2 * depth + 1
    leaf = layers[-1]
    assert all(len(r) in (3, 4) for r in leaf)
    counts = np.bincount(sum(leaf, tuple()), minlength=n_features)
    assert np.all(counts == reps)
    for part_level, region_level in zip(layers[1::2], layers[2::2]):
        assert [r for pair in part_level for r in pair] == region_level


def test_value_errors():
    with pytest.raises(ValueError):
        RegionGraph(0, depth=1)
    with pytest.raises(ValueError):
        RegionGraph(8, depth=0)
    with pytest.raises(ValueError):
        RegionGraph(8, depth=4)
    with pytest.raises(ValueError):
        RegionGraph(8, depth=2).make_layers(n_repetitions=-1)
    with pytest.raises(ValueError):
        RegionGraph(8, depth=2, random_state="seed")
