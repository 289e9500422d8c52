"""Stand-in for `higher` (git master, scripts/install.sh:3-4). TEST INFRASTRUCTURE ONLY."""
from . import optim  # noqa: F401
