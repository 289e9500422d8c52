"""Synthetic stand-ins for the reference's datasets. The reference's loaders (src/data/, Planetoid/UCI through
torch-geometric and sklearn) are one-off host preprocessing and out of scope; benchmarks
and parity runs use synthetic graphs of the same shapes."""
from .synthetic import SHAPES, make_dataset      # noqa: F401
from .dataloader import DataFactory               # noqa: F401
