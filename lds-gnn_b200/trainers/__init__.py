from typing import NamedTuple


class Metrics(NamedTuple):
    """(loss, accuracy) of one step, host floats (src/trainers/__init__.py:4-6)."""
    loss: float
    acc: float
