"""Host-side helpers with the reference's names (`gym_socks/utils/__init__.py`, `utils/batch.py`).

On the product path `normalize` and `indicator_fn` are fused into the CUDA kernels (sr.cu); these
numpy versions exist for API parity with user code that imports them and are not used by the
algorithms.
"""
import numpy as np

__all__ = ["normalize", "indicator_fn", "generate_batches"]


def normalize(v):
    """Column L1 normalisation (`gym_socks/utils/__init__.py:8-22`)."""
    return v / np.sum(v, axis=0)


def indicator_fn(points, space):
    """Closed-box membership (`gym_socks/utils/__init__.py:25-47`)."""
    return np.array(np.all(points >= space.low, axis=1) & np.all(points <= space.high, axis=1), dtype=bool)


def generate_batches(num_elements, batch_size):
    """Slices of at most `batch_size` (`gym_socks/utils/batch.py:1-28`)."""
    start = 0
    for _ in range(num_elements // batch_size):
        yield slice(start, start + batch_size)
        start += batch_size
    if start < num_elements:
        yield slice(start, num_elements)
