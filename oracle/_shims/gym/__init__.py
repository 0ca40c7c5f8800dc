"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference (Abmarl 0.2.7) imports `gym.spaces` for space *metadata*; `gym` is
not installed in this image.  This shim supplies just enough of the API for the
unmodified reference to import and run as the live oracle inside the build
container.  It is never imported by the product package.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, **kwargs):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, **kwargs):
        pass

    def close(self):
        pass
