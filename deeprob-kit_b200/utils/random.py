"""Random-state coercion (interface of deeprob/utils/random.py:11-27)."""
from typing import Optional, Union

import numpy as np

RandomState = Union[int, np.random.RandomState]


def check_random_state(random_state: Optional[RandomState] = None) -> np.random.RandomState:
    """None -> fresh RandomState, int -> seeded RandomState, RandomState -> itself; else ValueError."""
    if isinstance(random_state, np.random.RandomState):
        return random_state
    if random_state is None or isinstance(random_state, int):
        return np.random.RandomState(random_state)
    raise ValueError("The random state must be either None, a seed integer or a Numpy RandomState object")
