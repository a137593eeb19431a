"""reference src/envs/core.py:3-11 -- the construction entry point every script uses."""
from .spinsystem import SpinSystemFactory


def make(id, *args, **kwargs):
    if id == "SpinSystem":
        return SpinSystemFactory.get(*args, **kwargs)
    raise NotImplementedError()
