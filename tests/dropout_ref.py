"""NumPy mirror of the counter-based dropout generator of csrc/ratspn_dropout.cu (dpk_dropout_draw): the tests use it
to inject into the oracle exactly the Bernoulli draws the CUDA path derives from (seed, stream, element index)."""
import numpy as np

_M64 = (1 << 64) - 1


def draw(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """Uniform 24-bit integers, element-wise over `index` (uint64 array)."""
    with np.errstate(over="ignore"):
        idx = np.asarray(index, dtype=np.uint64)
        z = np.uint64((seed ^ (stream << 56)) & _M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(40)).astype(np.uint32)


def threshold(rate: float) -> int:
    return 0 if not rate > 0 else int(min(float(np.float32(rate)) * 16777216.0, 16777215.0))


def dropped(seed: int, stream: int, shape, rate: float) -> np.ndarray:
    """Boolean mask of the given element shape (C-order index = element index of the stream)."""
    n = int(np.prod(shape))
    return (draw(seed, stream, np.arange(n, dtype=np.uint64)) < threshold(rate)).reshape(shape)
