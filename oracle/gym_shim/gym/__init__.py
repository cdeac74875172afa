"""Test-only stand-in for the `gym` package (gym is not installed in this image).

ORACLE INFRASTRUCTURE — not product code.  It provides exactly the symbols the
reference's hot-path modules touch at import time, so that
`/root/reference/gym_socks` can be imported unmodified by `oracle/make_golden.py`.
Nothing under `socks_b200/` may import this.
"""
from . import spaces, utils, envs  # noqa: F401


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]


def make(id, **kwargs):
    return envs.registration.make(id, **kwargs)
