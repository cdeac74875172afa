"""Sample transposition, as `gym_socks/sampling/transform.py:4-26` (host glue)."""


def transpose_sample(sample):
    """[(x_1, u_1, y_1), ...] -> ([x_1, ...], [u_1, ...], [y_1, ...])."""
    return tuple(map(list, zip(*sample)))
