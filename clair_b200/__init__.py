"""clair_b200: B200-native (sm_100a) implementation of Clair's batched variant-calling forward path."""
from . import param, weights, synth  # noqa: F401

__all__ = ["Clair", "param", "weights", "synth"]


def __getattr__(name):
    if name == "Clair":
        from .model import Clair
        return Clair
    raise AttributeError(name)
